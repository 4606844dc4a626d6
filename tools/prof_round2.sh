#!/bin/bash
# Round-2 profile capture (run under gpurun), one STEP per call so that every capture is summarised and copied to gpurun_out/
# before the next one starts (a whole-script run once hit its time limit in the last step and lost everything):
#   bash tools/prof_round2.sh r02 launches|mwt|dwt|vit|bbglue|bbgemm [...]
# Plain runs come first (each must exit 0 without ncu), then the ncu pass of the same command.  The .ncu-rep files stay on the box
# (tens of MB each); tools/make_profile_md.py writes the per-launch summaries that are committed under profiles/.
TAG=${1:-r02}; shift
O=/tmp/prof_${TAG}
D=gpurun_out/profiles_${TAG}
mkdir -p $O $D
NCU="ncu --set full --clock-control none --import-source on"
for STEP in "$@"; do
  case $STEP in
    launches)   # launch list of the bench command (shares of the step)
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/plain_bench.log 2>&1 || exit 1
      # (the first 450 launches = the warm-up forwards: every forward of the command launches the same kernels, so the shares are
      # those of a step; profiling all ~3000 launches of the command costs minutes of box time for the same table)
      timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 450 --csv --log-file $O/launches_${TAG}.csv \
          python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_launch.log 2>&1 ;;
    mwt)        # MWT branch at 512 frames, second forward only: DWT, upsample3, three-level head conv, 3 x fusion conv, multiscale, freq_conv, pool conv
      # MWTN frames (default 128): ncu saves and restores every buffer a kernel writes for each of its ~40 replay passes, and at 512
      # frames those are GBs per kernel (the step then takes > 10 min of box time); 128 frames = 12996 tiles, 88 per SM
      python tools/profile_target.py ${MWTN:-128} > $O/plain_target.log 2>&1 || exit 1
      timeout 300 $NCU -k regex:"gemm_tc_kernel|dwt3_haar|mwt_upsample" --launch-skip 9 --launch-count 9 -o $O/prof_${TAG}_mwt${MWTN:-128} -f \
          python tools/profile_target.py ${MWTN:-128} > $O/ncu_mwt.log 2>&1 ;;
    dwt)        # standalone DWT, BASELINE configs[1] (256 frames, all six outputs)
      python tools/dwt_bench.py 256 > $O/plain_dwt.log 2>&1 || exit 1
      timeout 200 $NCU -k regex:dwt3_haar --launch-skip 5 --launch-count 1 -o $O/prof_${TAG}_dwt256 -f python tools/dwt_bench.py 256 > $O/ncu_dwt.log 2>&1 ;;
    vit)        # SFE head at 512 frames, second pass: split-K patch embedding + the ViT linears (EPI_PARTIAL / EPI_LINEAR flavours) + glue
      python tools/profile_vit.py 512 > $O/plain_vit.log 2>&1 || exit 1
      timeout 240 $NCU -k regex:"gemm_tc_kernel|splitk_reduce|vit_|layernorm" --launch-skip 19 --launch-count 19 -o $O/prof_${TAG}_vit512 -f \
          python tools/profile_vit.py 512 > $O/ncu_vit.log 2>&1 ;;
    bbglue)     # backbone glue kernels of one forward (stem, direct 24->24 convs, first depthwise / SE pairs), 128 frames keep the replays short
      python tools/backbone_bench.py 128 > $O/plain_bb.log 2>&1 || exit 1
      timeout 300 $NCU -k regex:"dwconv|se_gate|conv3x3_c24|stem_conv" --launch-skip 63 --launch-count 14 -o $O/prof_${TAG}_bbglue128 -f \
          python tools/backbone_bench.py 128 > $O/ncu_bbglue.log 2>&1 ;;
    bbgemm)     # a window of the backbone's tensor-core convs in stages 4-6
      timeout 300 $NCU -k regex:gemm_tc --launch-skip 94 --launch-count 10 -o $O/prof_${TAG}_bbgemm128 -f python tools/backbone_bench.py 128 > $O/ncu_bbgemm.log 2>&1 ;;
  esac
  python tools/make_profile_md.py ${TAG} $O $D > $O/md.log 2>&1
  cp $O/launches_${TAG}.csv $O/*.log $D/ 2>/dev/null
  echo "step $STEP done: $(ls $D | wc -l) files"
done
