#!/bin/bash
# Round profile capture (run under gpurun): plain runs first (must exit 0), then the ncu passes of the same commands.
# usage: bash tools/prof_round.sh r01c
TAG=${1:-r01c}
O=gpurun_out
set -x
python bench.py --steps 10 --warmup 3 > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err || exit 1
python tools/profile_target.py 512 > $O/plain_target_${TAG}.log 2>&1 || exit 1
# 1. launch list of the bench command (shares of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launch_${TAG}.log 2>&1
# 2. full-set capture of the MWT branch at 512 frames: second forward only (9 tensor-core launches + glue per forward)
ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|dwt3_haar|mwt_upsample" --launch-skip 13 --launch-count 13 \
    -o $O/prof_${TAG}_mwt512 -f python tools/profile_target.py 512 > $O/ncu_full_${TAG}.log 2>&1
# 3. standalone DWT, BASELINE configs[1] (256 frames, all six outputs)
ncu --set full --clock-control none --import-source on -k regex:dwt3_haar --launch-skip 5 --launch-count 1 \
    -o $O/prof_${TAG}_dwt256 -f python tools/dwt_bench.py 256 > $O/ncu_dwt_${TAG}.log 2>&1
ls -la $O | tail -12
