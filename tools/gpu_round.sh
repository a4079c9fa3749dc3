#!/bin/bash
# One full GPU-box visit: whole GPU suite, compute-sanitizer over the kernel variants, bench (both arms), ncu launch list and one
# full capture of the step kernel at the loaded state, the other BASELINE configs.
mkdir -p gpurun_out
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 1500 bash tools/gpu_sanitize.sh > gpurun_out/sanitize.log 2>&1; cat gpurun_out/sanitizer/summary.txt | cut -c1-60,200-330
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1200 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 480 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tsc_step -s 372 -c 1 -f -o gpurun_out/prof \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
for cfgn in jinan manhattan grid16; do
  timeout 900 python bench.py --config $cfgn --steps 100 --warmup 10 --cpu-steps 20 > gpurun_out/bench_$cfgn.json 2> gpurun_out/bench_$cfgn.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$cfgn.json")); k=d["config"]["kernel"]
    print("$cfgn: dev ms %.4f  e2e ms %.4f  V %.1f  frac %.4f  match %s  cpu %.0f  nt %d x %d" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["mean_running_vehicles"], d["roofline"]["frac"], d["e2e"]["matches_device_leg"], d["cpu_baseline"]["value"], k["threads"], k["blocks_per_sm"]))
except Exception as e:
    print("$cfgn failed", e); print(open("gpurun_out/bench_$cfgn.err").read()[-1500:])
PY
done
timeout 300 python tools/phase_timing.py 640 > gpurun_out/phase_timing.txt 2>&1; cat gpurun_out/phase_timing.txt
ls -la gpurun_out | head -40
