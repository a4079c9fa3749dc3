#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_episode_gpu.py -m gpu -x -q > gpurun_out/pytest_engine.log 2>&1; tail -8 gpurun_out/pytest_engine.log
python tools/phase_timing.py 640 > gpurun_out/phase_timing.txt 2>&1; cat gpurun_out/phase_timing.txt
: > gpurun_out/sweep.txt
run() {
  name=$1; envs=$2; shift 2
  ( env $envs timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline "$@" 2> gpurun_out/sweep_$name.err ) | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); k=d['config']['kernel']
    print('$name: dev ms %.4f  e2e ms %.4f  V %.1f  frac %.4f  match %s  nt %d x %d regs %d smem %d clocks %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['mean_running_vehicles'], d['roofline']['frac'], d['e2e']['matches_device_leg'], k['threads'], k['blocks_per_sm'], k['regs'], k['smem_bytes'], d['clocks']))
except Exception as e:
    print('$name failed', e)
" | tee -a gpurun_out/sweep.txt
}
run default "X=1"
run 192x4 "TSC_B200_THREADS=192"
run jinan "X=1" --config jinan
run manhattan "X=1" --config manhattan
run grid16 "X=1" --config grid16
