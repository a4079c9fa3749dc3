#!/bin/bash
# variant sweep on the bench workload: "name|env assignments|extra bench args"
mkdir -p gpurun_out; : > gpurun_out/sweep.txt
run() {
  name=$1; envs=$2; shift 2
  ( env $envs timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline "$@" 2> gpurun_out/sweep_$name.err ) | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); k=d['config']['kernel']
    print('$name: dev ms %.4f  e2e ms %.4f  V %.1f  frac %.4f  match %s  nt %d x %d regs %d smem %d' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['mean_running_vehicles'], d['roofline']['frac'], d['e2e']['matches_device_leg'], k['threads'], k['blocks_per_sm'], k['regs'], k['smem_bytes']))
except Exception as e:
    print('$name failed', e)
" | tee -a gpurun_out/sweep.txt
}
run base_192x4_cap640 "X=1"
run 192x4_cap600 "X=1" --vehicle-capacity 600
run 160x5_cap600 "TSC_B200_THREADS=160" --vehicle-capacity 600
run 192x5_cap600 "TSC_B200_THREADS=192 TSC_B200_MIN_BLOCKS=5" --vehicle-capacity 600
run 256x4_cap600 "TSC_B200_THREADS=256 TSC_B200_MIN_BLOCKS=4" --vehicle-capacity 600
run 256x3_cap600 "TSC_B200_THREADS=256 TSC_B200_MIN_BLOCKS=3" --vehicle-capacity 600
run noprefetch "TSC_B200_PREFETCH=0"
run jinan "X=1" --config jinan
run manhattan_1520 "X=1" --config manhattan --vehicle-capacity 1520
run manhattan_1560 "X=1" --config manhattan
