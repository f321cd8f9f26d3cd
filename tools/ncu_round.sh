#!/bin/bash
# One ncu session per round (run under gpurun, one GPU): `--set full` captures of every hot kernel at the benchmark size + the traversal
# kernels on sponza, and the launch list of the bench command (kernel shares of a step).  Reports land in gpurun_out/<tag>_*.ncu-rep;
# read them here with tools/ncu_summary.py / tools/ncu_source.py and commit the summaries under profiles/.
# usage: tools/ncu_round.sh <tag>
set -u
TAG=$1
mkdir -p gpurun_out
tools/ncu_capture.sh $TAG primref_extents morton30 radix_count radix_scan radix_scatter lbvh_tile lbvh_group lbvh_climb collapse_expand collapse_number collapse_emit ploc_merge hploc_kernel
NCU="ncu --set full --clock-control none --import-source on"
for K in traverse_kernel traverse_step_kernel traverse_wide4_kernel; do
  timeout 300 $NCU -k regex:$K -c 1 -f -o gpurun_out/${TAG}_$K python tools/trace_bench.py --mesh sponza > gpurun_out/${TAG}_$K.log 2>&1
  tail -1 gpurun_out/${TAG}_$K.log
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/${TAG}_launches_bench.log 2>&1
ls -la gpurun_out/${TAG}_*.ncu-rep | wc -l
