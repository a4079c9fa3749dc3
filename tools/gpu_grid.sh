#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_grid_gpu.py tests/test_engine_gpu.py -m gpu -x -q > gpurun_out/pytest_grid.log 2>&1; tail -5 gpurun_out/pytest_grid.log
run() {
  name=$1; envs=$2; shift 2
  ( env $envs timeout 600 python bench.py --config grid16 --steps 40 --warmup 5 --no-cpu-baseline "$@" 2> gpurun_out/grid_$name.err ) | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); k=d['config']['kernel']
    print('$name: dev ms %.4f  e2e ms %.4f  V %.1f match %s  nt %d x %d regs %d smem %d gw %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['mean_running_vehicles'], d['e2e']['matches_device_leg'], k['threads'], k['blocks_per_sm'], k['regs'], k['smem_bytes'], k['global_workspace']))
except Exception as e:
    print('$name failed', e)
" | tee -a gpurun_out/grid.txt
}
: > gpurun_out/grid.txt
run hybrid "X=1"
run allglobal "TSC_B200_GMEM_META_SHARED=0"
run hybrid_generic_tmpl "TSC_B200_ONE_TEMPLATE=0"
