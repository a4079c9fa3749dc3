#!/bin/bash
# usage: bash tools/gpu_multi.sh N   -- both bench arms on N GPUs of one box, the way the driver launches them
N=${1:-2}
mkdir -p gpurun_out
nproc; nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_ref_$N.json 2> gpurun_out/scale_ref_$N.err; tail -c 400 gpurun_out/scale_ref_$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps ${LONG_STEPS:-200} --warmup 20 > gpurun_out/scale_long_$N.json 2> gpurun_out/scale_long_$N.err
python - <<PY
import json
for f in ("scale_$N", "scale_long_$N"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.json" % f) if l.startswith("{")][-1])
        print(f, "n_gpus", d["n_gpus"], "dev ms %.4f value %.1fM  e2e ms %.4f e2e %.1fM  host_threads %s affinity %s clocks %s" % (d["ms_per_step"], d["value"]/1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["host_threads"], d["config"]["host_cpu_affinity"], d["clocks"]))
    except Exception as e:
        print(f, "failed", e); print(open("gpurun_out/%s.err" % f).read()[-2000:])
try:
    r=json.loads([l for l in open("gpurun_out/scale_ref_$N.json") if l.startswith("{")][-1]); print("reference arm", r["value"], r["cpu_baseline"]["cores"])
except Exception as e:
    print("ref failed", e)
PY
