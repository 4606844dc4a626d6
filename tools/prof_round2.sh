#!/bin/bash
# Round-2 profile capture (run under gpurun): plain runs first (each must exit 0), then the ncu passes of the same commands.
# usage: bash tools/prof_round2.sh r02
TAG=${1:-r02}
O=/tmp/prof_${TAG}          # ncu reports stay on the box; only the summaries (tools/make_profile_md.py) come back via gpurun_out/
mkdir -p $O gpurun_out/profiles_${TAG}
NCU="ncu --set full --clock-control none --import-source on"
set -x
python bench.py --steps 20 --warmup 5 > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err || exit 1
python tools/profile_target.py 512 > $O/plain_target_${TAG}.log 2>&1 || exit 1
python tools/profile_vit.py 512 > $O/plain_vit_${TAG}.log 2>&1 || exit 1
python tools/backbone_bench.py 512 > $O/plain_bb_${TAG}.log 2>&1 || exit 1
# 1. launch list of the bench command (shares of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launch_${TAG}.log 2>&1
# 2. full-set capture of the MWT branch at 512 frames: second forward only (DWT + 3 x (upsample, head conv, fusion conv) + multiscale + ...)
$NCU -k regex:"gemm_tc_kernel|dwt3_haar|mwt_upsample" --launch-skip 13 --launch-count 13 \
    -o $O/prof_${TAG}_mwt512 -f python tools/profile_target.py 512 > $O/ncu_full_${TAG}.log 2>&1
# 3. standalone DWT, BASELINE configs[1] (256 frames, all six outputs)
$NCU -k regex:dwt3_haar --launch-skip 5 --launch-count 1 -o $O/prof_${TAG}_dwt256 -f python tools/dwt_bench.py 256 > $O/ncu_dwt_${TAG}.log 2>&1
# 4. SFE head at 512 frames, second pass: split-K patch embedding + the ViT linears (EPI_PARTIAL / EPI_LINEAR flavours)
$NCU -k regex:"gemm_tc_kernel|splitk_reduce|vit_|layernorm" --launch-skip 19 --launch-count 19 \
    -o $O/prof_${TAG}_vit512 -f python tools/profile_vit.py 512 > $O/ncu_vit_${TAG}.log 2>&1
# 5. backbone: glue kernels and a window of the tensor-core convs in stages 4-6
$NCU -k regex:"dwconv|se_gate|conv3x3_c24|stem_conv" --launch-skip 189 --launch-count 14 -o $O/prof_${TAG}_bbglue -f \
    python tools/backbone_bench.py 512 > $O/ncu_bbglue_${TAG}.log 2>&1
$NCU -k regex:gemm_tc --launch-skip 325 --launch-count 10 -o $O/prof_${TAG}_bbgemm_a -f python tools/backbone_bench.py 512 > $O/ncu_bbgemm_a_${TAG}.log 2>&1
python tools/make_profile_md.py ${TAG} $O gpurun_out/profiles_${TAG}
cp $O/bench_${TAG}.json $O/bench_${TAG}.err $O/plain_*_${TAG}.log $O/launches_${TAG}.csv gpurun_out/profiles_${TAG}/ 2>/dev/null
ls -la $O gpurun_out/profiles_${TAG} | tail -30
