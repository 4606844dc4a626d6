#!/bin/bash
# ncu --set full captures of the backbone kernels at 512 frames (run under gpurun): bash tools/prof_bb.sh r01e
TAG=${1:-r01e}
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
python tools/backbone_bench.py 512 > $O/plain_bb_${TAG}.log 2>&1 || exit 1
# glue kernels: 63 launches per forward (stem 1, direct 24->24 convs 2, depthwise 30, SE gates 30); skip the 3 warm-up forwards
$NCU -k regex:"dwconv|se_gate|conv3x3_c24|stem_conv" --launch-skip 189 --launch-count 14 -o $O/prof_${TAG}_bbglue -f \
    python tools/backbone_bench.py 512 > $O/ncu_bbglue_${TAG}.log 2>&1
# tensor-core convs: 107 launches per forward; a window in stages 4-5 (expand / SE-gated project convs) and the stage-2 window convs
$NCU -k regex:gemm_tc --launch-skip 325 --launch-count 10 -o $O/prof_${TAG}_bbgemm_a -f python tools/backbone_bench.py 512 > $O/ncu_bbgemm_a_${TAG}.log 2>&1
$NCU -k regex:gemm_tc --launch-skip 358 --launch-count 8 -o $O/prof_${TAG}_bbgemm_b -f python tools/backbone_bench.py 512 > $O/ncu_bbgemm_b_${TAG}.log 2>&1
ls -la $O | tail -6
