set -x
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:se_gate --launch-skip 97 --launch-count 1 -o gpurun_out/prof_r01c_segate -f python tools/backbone_bench.py 512 > gpurun_out/ncu_c1.log 2>&1
$NCU -k regex:dwconv --launch-skip 97 --launch-count 1 -o gpurun_out/prof_r01c_dw -f python tools/backbone_bench.py 512 > gpurun_out/ncu_c2.log 2>&1
$NCU -k regex:gemm_tc --launch-skip 269 --launch-count 4 -o gpurun_out/prof_r01c_conv1 -f python tools/backbone_bench.py 512 > gpurun_out/ncu_c3.log 2>&1
ls -la gpurun_out/
