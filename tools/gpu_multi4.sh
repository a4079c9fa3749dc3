#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_registered_host_gpu.py -m gpu -x -q > gpurun_out/pytest_reg.log 2>&1; tail -3 gpurun_out/pytest_reg.log
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_1.json 2> gpurun_out/bench_1.err; python -c "
import json
d=json.load(open('gpurun_out/bench_1.json')); print('N=1: dev ms %.4f e2e ms %.4f policy %.3f %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['host_policy_ms_per_step'], d['e2e']['environments']))"
LONG_STEPS=100 bash tools/gpu_multi.sh 4
