#!/bin/bash
# parity tests, then bench (no CPU baseline) under a list of env settings: tools/gpu_ab.sh "VAR=1" "VAR=2 OTHER=3" ...
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log; fi
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 300 python bench.py --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/ab$i.json 2> gpurun_out/ab$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab$i.json"))
    print("[$cfg]", "ms %.4f"%d["ms_per_step"], "e2e ms %.3f"%d["e2e"]["ms_per_step"], "policy ms %.3f"%d["e2e"].get("host_policy_ms_per_step",-1), "frac %.4f"%d["roofline"]["frac"], d["config"]["kernel"])
except Exception as e:
    print("[$cfg] failed", e); print(open("gpurun_out/ab$i.err").read()[-1500:])
PY
done
