#!/usr/bin/env python
"""GPU box: per-phase cycle breakdown of tsc_step_kernel on the bench workload.
usage: python tools/phase_timing.py [vehicle_capacity] [replicas]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pytsc_b200 import _build
TIMING_LIB = os.path.join(ROOT, "tools", "ab", "lib_timing.so")      # built here (CPU container) so that it travels to the GPU box
if not os.path.exists(TIMING_LIB) or os.path.getmtime(TIMING_LIB) < os.path.getmtime(_build.SRC):
    os.makedirs(os.path.dirname(TIMING_LIB), exist_ok=True)
    _build.build_variant(TIMING_LIB, ["-DTSC_PHASE_TIMING"])
os.environ["TSC_B200_LIB"] = TIMING_LIB
if "--build-only" in sys.argv:
    sys.exit(0)
import torch
import bench
from pytsc_b200.backend.config import Config
from pytsc_b200.backend.network_parser import NetworkParser
from pytsc_b200.binding import Engine
from pytsc_b200.scenario import compile_scenario

cap = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
W = bench.workload(os.environ.get("BENCH_CONFIG", "hangzhou"))
cfg = Config(W["scenario"], **W["kw"])
cs = compile_scenario(cfg, NetworkParser(cfg))
eng = Engine(cs, B, 0, vehicle_capacity=cap)
bufs = eng.alloc_outputs(["obs", "reward", "reward_global", "mask", "lane_count", "lane_queued", "lane_occupancy", "lane_mean_speed", "sim"])
eng.init_program(0)
for _ in range(400):
    eng.env_step(None, bufs, n_ticks=5, controller=1, controller_arg=25)
torch.cuda.synchronize()
eng.debug_timing(True)
N = 100
t0 = time.perf_counter()
for _ in range(N):
    eng.env_step(None, bufs, n_ticks=5, controller=1, controller_arg=25)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
cyc = eng.debug_timing(False)
extra = {k: cyc.pop(k) for k in list(cyc) if k.startswith("n_")}
tot = sum(cyc.values())
info = eng.kernel_info()
print(f"cap {cap} B {B} kernel {info}  {1e3 * wall / N:.3f} ms/step (with timing on)  running {int(bufs['sim'][0,0])}")
per = B * N
TICK = ("spawn", "decisions", "cross", "leave", "enter")
for k, v in cyc.items():
    print(f"  {k:11s} {100 * v / tot:5.1f}%  {v / per:9.0f} cycles / replica-step" + (f"  ({v / per / 5:7.0f} / tick)" if k in TICK else ""))
print(f"  total       {tot / per:9.0f} cycles / replica-step")
print("  per tick: " + ", ".join(f"{k[2:]} {v / (per * 5):.1f}" for k, v in extra.items()))
eng.check()
