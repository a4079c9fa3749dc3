#!/bin/bash
mkdir -p gpurun_out
for cap in 4 0; do
BENCH_CORES_PER_RANK=$cap timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2951$cap bench.py --gpus 4 --steps 60 --warmup 5 > gpurun_out/scale4_cap$cap.json 2> gpurun_out/scale4_cap$cap.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale4_cap$cap.json") if l.startswith("{")][-1])
    print("cap $cap: dev ms %.4f value %.1fM  e2e ms %.4f e2e %.1fM policy %.3f | %s | %s" % (d["ms_per_step"], d["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["host_policy_ms_per_step"], d["e2e"]["environments"], d["config"]["host_cpu_affinity"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/scale4_cap$cap.err").read()[-1500:])
PY
done
