#!/bin/bash
# ncu captures of the hot kernels at the benchmark size (run under gpurun; outputs land in gpurun_out/).
# usage: tools/ncu_capture.sh <tag> [kernels...]   e.g. tools/ncu_capture.sh r01 onesweep_pass lbvh_fused
set -u
TAG=$1; shift
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
for K in "$@"; do
  case $K in
    ploc_iter|ploc_setup|ploc_tail|ploc_merge) ALGO=ploc;;
    hploc|hploc_setup|hploc_kernel|hploc_tile) ALGO=hploc;;
    lbvh_karras_emit|lbvh_refit) ALGO="twopass --two-kernel";;
    lbvh_fused_karras) ALGO=twopass;;
    split_level|split_remap) ALGO=split;;
    batched_lbvh) ALGO=batched;;
    *) ALGO=singlepass;;
  esac
  case $K in
    collapse_number) CNT="-c 1";;
    radix_scatter|radix_count) CNT="-c 4";;
    ploc_iter) CNT="-c 5";;
    onesweep_pass) CNT="-c 4";;
    *) CNT="-c 1";;
  esac
  timeout 240 $NCU -k regex:$K $CNT -f -o gpurun_out/${TAG}_$K python tools/kernel_bench.py --algo $ALGO --reps 1 --warmup 0 > gpurun_out/${TAG}_$K.log 2>&1
  tail -2 gpurun_out/${TAG}_$K.log
done
