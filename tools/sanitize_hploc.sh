set -u
OUT=gpurun_out/sanitize_hploc; mkdir -p $OUT; : > $OUT/summary.txt
CS=/usr/local/cuda/bin/compute-sanitizer
run() { local tool=$1 subset=$2; shift 2; local log="$OUT/${tool}_${subset}.log"; local t0=$(date +%s)
  timeout 420 $CS --tool $tool "$@" --error-exitcode 66 python tools/sanitize_driver.py $subset > $log 2>&1; local rc=$?
  echo "$tool [$subset] rc=$rc ok=$(grep -c 'sanitize_driver ok' $log) $(( $(date +%s) - t0 ))s :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)" | tee -a $OUT/summary.txt; }
export B2_SANITIZE_N=40000
run memcheck hploc --leak-check full
run synccheck hploc
run racecheck hploc --racecheck-report all
run initcheck hploc
