#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nproc
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_$N.json") if l.startswith("{")][-1])
    print("n_gpus", d["n_gpus"], "dev ms %.4f value %.1fM  e2e ms %.4f e2e %.1fM policy %.3f | %s | %s | clocks %s" % (d["ms_per_step"], d["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["host_policy_ms_per_step"], d["e2e"]["environments"], d["config"]["host_cpu_affinity"], d["clocks"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/scale_$N.err").read()[-2000:])
PY
