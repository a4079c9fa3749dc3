#!/bin/bash
# compute-sanitizer racecheck + memcheck over every kernel variant (run on the GPU box via gpurun).
# Logs: gpurun_out/sanitizer/<tool>__<variant>.log ; summary: gpurun_out/sanitizer/summary.txt
out=gpurun_out/sanitizer; mkdir -p $out; : > $out/summary.txt
run() {   # name, env assignments, args...
  name=$1; envs=$2; shift 2
  for tool in racecheck memcheck; do
    log=$out/${tool}__${name}.log
    ( env $envs timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py "$@" ) > $log 2>&1
    rc=$?
    echo "$tool $name rc=$rc :: $(grep -E 'sanitize_case ok' $log | head -1 | cut -c1-160) :: $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' $log | tail -1)" >> $out/summary.txt
  done
}
if [ -n "$SANITIZE_ALL" ]; then
run t192x4            "TSC_B200_THREADS=192" --capacity 600
run t192x5            "TSC_B200_THREADS=192 TSC_B200_MIN_BLOCKS=5" --capacity 560
run t160x5            "TSC_B200_THREADS=160" --capacity 560
run t256x2            "TSC_B200_THREADS=256" --capacity 1560
run stage_cp_async    "TSC_B200_ASYNC_STAGE=1" --capacity 600
run stage_plain       "TSC_B200_ASYNC_STAGE=0" --capacity 600
run ctl_greedy        "X=1" --capacity 600 --controller greedy
run ctl_sotl          "X=1" --capacity 600 --controller sotl
fi
run t256x4_fixed672   "X=1" --capacity 640
run gmem1024_hybrid   "TSC_B200_GMEM=1" --capacity 600
run gmem1024_global   "TSC_B200_GMEM=1 TSC_B200_GMEM_META_SHARED=0" --capacity 600
run registered_host   "X=1" --capacity 640 --registered
if [ -z "$SANITIZE_SHORT" ]; then
run t256x4_generic    "X=1" --capacity 600
run t256x3_fixed1184  "X=1" --capacity 1150
run t384x2_fixed1568  "X=1" --capacity 1530
run t512              "X=1" --capacity 2000
run ctl_max_pressure  "X=1" --capacity 600 --controller max_pressure --obs position_matrix
fi
cat $out/summary.txt
