#!/bin/bash
# compute-sanitizer over every kernel of libb2bvh.so (SURVEY §5 "race detection"; the reference's known hazards: HplocKernel.h:183-185,
# spin-wait persistent kernels, Appendix C).  Run on a GPU box:  tools/sanitize.sh [out_dir]   (default gpurun_out/sanitize)
#   memcheck   out-of-bounds / misaligned global, shared and local accesses, leaks of device memory at context teardown
#   racecheck  shared-memory hazards between the threads of a CTA (the tile / merge / ranking kernels keep their state in shared memory)
#   synccheck  divergent or mismatched barriers (named barriers of shrinking width, mbarrier waits)
#   initcheck  reads of device memory that was never written (the scratch buffers are reused across builds and stages)
# One summary line per tool lands in $OUT/summary.txt; the full logs sit next to it.
set -u
OUT=${1:-gpurun_out/sanitize}
mkdir -p "$OUT"
CS=${COMPUTE_SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
PY=${PYTHON:-python}
: > "$OUT/summary.txt"
run() {  # tool, subset, extra flags...
  local tool=$1 subset=$2; shift 2
  local log="$OUT/${tool}_${subset//,/+}.log"
  local t0=$(date +%s)
  timeout ${B2_SANITIZE_TIMEOUT:-900} "$CS" --tool "$tool" "$@" --error-exitcode 66 "$PY" tools/sanitize_driver.py "$subset" > "$log" 2>&1
  local rc=$?
  local errs=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
  local ok=$(grep -c "sanitize_driver ok" "$log")
  echo "$tool [$subset] rc=$rc driver_ok=$ok $(( $(date +%s) - t0 ))s :: ${errs:-no summary line}" | tee -a "$OUT/summary.txt"
}
export B2_SANITIZE_N=${B2_SANITIZE_N:-40000}
run memcheck  lbvh,ploc,hploc,sizes,split,batched,m60,sort,trace --leak-check full
run synccheck lbvh,ploc,hploc,sizes,split,batched,m60,sort,trace
run racecheck lbvh,sort,split,batched,m60 --racecheck-report all
run racecheck ploc,hploc,trace --racecheck-report all
run memcheck  large --leak-check full
run racecheck large --racecheck-report all
export B2_SANITIZE_N=20000
run initcheck lbvh,ploc,hploc,split,batched,m60,sort,trace
cat "$OUT/summary.txt"
