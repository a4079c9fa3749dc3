#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_episode_gpu.py tests/test_state_gpu.py tests/test_edge_gpu.py -m gpu -x -q > gpurun_out/pytest_engine.log 2>&1; tail -8 gpurun_out/pytest_engine.log
: > gpurun_out/sweep.txt
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
run bulk_default "X=1"
run cp_async "TSC_B200_ASYNC_STAGE=1"
run plain "TSC_B200_ASYNC_STAGE=0"
run manhattan "X=1" --config manhattan
run manhattan_256 "TSC_B200_THREADS=256" --config manhattan
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_case.py --capacity 600 > gpurun_out/memcheck_bulk.log 2>&1; tail -3 gpurun_out/memcheck_bulk.log
timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_case.py --capacity 600 > gpurun_out/racecheck_bulk.log 2>&1; tail -3 gpurun_out/racecheck_bulk.log
