#!/bin/bash
# GPU visit 2: whole GPU test suite on the new ABI (flow sets, registered host path), bench in the loaded regime, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; tail -c 2500 gpurun_out/bench_short.json; tail -5 gpurun_out/bench_short.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
for t in 1 2 4 8 12; do TSC_B200_HOST_THREADS=$t timeout 300 python bench.py --steps 60 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('threads $t', 'dev ms %.3f'%d['ms_per_step'], 'e2e ms %.3f'%d['e2e']['ms_per_step'], 'policy ms %.3f'%d['e2e']['host_policy_ms_per_step'], 'match', d['e2e']['matches_device_leg'], 'V', d['config']['mean_running_vehicles'])"; done
python tools/phase_timing.py 640 > gpurun_out/phase_timing.txt 2>&1; cat gpurun_out/phase_timing.txt
