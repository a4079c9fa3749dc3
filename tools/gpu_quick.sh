#!/bin/bash
# parity tests, phase timing, then bench (no CPU baseline) at the given capacities
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
for cap in ${CAPS:-1024 600}; do
  python tools/phase_timing.py $cap
  timeout 300 python bench.py --no-cpu-baseline --vehicle-capacity $cap > gpurun_out/bench_cap$cap.json 2> gpurun_out/bench_cap$cap.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_cap$cap.json"))
    print("cap $cap", "value %.3e"%d["value"], "ms %.3f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], "frac %.4f"%d["roofline"]["frac"], d["config"]["kernel"])
except Exception as e:
    print("cap $cap failed", e); print(open("gpurun_out/bench_cap$cap.err").read()[-2000:])
PY
done
