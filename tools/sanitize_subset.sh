#!/bin/bash
# development: all four sanitizer tools over one subset of tools/sanitize_driver.py —  tools/sanitize_subset.sh <subset> [out_dir]
set -u
SUB=${1:-lbvh}; OUT=${2:-gpurun_out/sanitize_$SUB}; mkdir -p $OUT; : > $OUT/summary.txt
CS=/usr/local/cuda/bin/compute-sanitizer
run() { local tool=$1; shift; local log="$OUT/${tool}.log"; local t0=$(date +%s)
  timeout 420 $CS --tool $tool "$@" --error-exitcode 66 python tools/sanitize_driver.py $SUB > $log 2>&1; local rc=$?
  echo "$tool [$SUB] rc=$rc ok=$(grep -c 'sanitize_driver ok' $log) $(( $(date +%s) - t0 ))s :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)" | tee -a $OUT/summary.txt; }
export B2_SANITIZE_N=${B2_SANITIZE_N:-40000}
run memcheck --leak-check full
run synccheck
run racecheck --racecheck-report all
run initcheck
