#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_episode_gpu.py -m gpu -x -q > gpurun_out/pytest_engine.log 2>&1; tail -8 gpurun_out/pytest_engine.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_short.json"))
    print("bench short: dev ms %.4f  e2e ms %.4f  policy %.3f  V %.1f  frac %.4f  match %s  kernel %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["host_policy_ms_per_step"], d["config"]["mean_running_vehicles"], d["roofline"]["frac"], d["e2e"]["matches_device_leg"], d["config"]["kernel"]))
except Exception as e:
    print("bench short failed", e); print(open("gpurun_out/bench_short.err").read()[-3000:])
PY
python tools/phase_timing.py 640 > gpurun_out/phase_timing.txt 2>&1; cat gpurun_out/phase_timing.txt
timeout 600 python -m pytest tests/test_plugin_gpu.py -m gpu -q > gpurun_out/pytest_plugin.log 2>&1; tail -8 gpurun_out/pytest_plugin.log
