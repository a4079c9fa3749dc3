#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_registered_host_gpu.py -m gpu -x -q > gpurun_out/pytest_reg.log 2>&1; tail -3 gpurun_out/pytest_reg.log
run() {
name=$1; shift
timeout 300 "$@" python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; python -c "
import json
d=json.load(open('gpurun_out/bench_$name.json')); print('$name: dev ms %.4f e2e ms %.4f policy %.3f match %s %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['host_policy_ms_per_step'], d['e2e']['matches_device_leg'], d['e2e']['environments']))"
}
run all env X=1
run cpus4 taskset -c 0-3
run cpus2 taskset -c 0-1
