#!/usr/bin/env python
"""GPU box: device-resident throughput of the other BASELINE.json configs on ONE
GPU, each with its per-GPU share of the batch, in-kernel fixed-time control (green 25 s), L2 flushed
between launches, CUDA events.  bench.py stays the contract line (config 2); this is the side table.
usage: python tools/bench_configs.py [jinan manhattan grid16 hangzhou]  -> gpurun_out/configs.json"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pytsc_b200.backend.config import Config  # noqa: E402
from pytsc_b200.backend.network_parser import NetworkParser  # noqa: E402
from pytsc_b200.binding import Engine  # noqa: E402
from pytsc_b200.scenario import compile_scenario  # noqa: E402

LF = dict(observation_space="lane_features", action_space="phase_selection", round_robin=False)
CONFIGS = {
    # name: (scenario, kwargs, replicas on this GPU, vehicle capacity, warm-up steps, timed steps, note)
    "hangzhou": ("hangzhou_4_4", dict(cityflow=dict(flow_file="anon_4_4_hangzhou_real.json"), signal=dict(LF, reward_function="max_pressure")),
                 4096, 640, 72, 648, "config 2: B = 4096 on one GPU"),
    "jinan": ("jinan_3_4", dict(signal=dict(LF, reward_function="queue_length")), 2048, 1150, 72, 648,
              "config 3: B = 16384 over 8 GPUs = 2048 per GPU"),
    "manhattan": ("manhattan_16_3", dict(signal=dict(LF, reward_function="queue_length")), 4096, 1560, 72, 648,
                  "config 5: B = 4096, 3600 s horizon"),
    "grid16": (None, dict(signal=dict(LF, reward_function="max_pressure")), 128, 24000, 24, 96,
               "config 4: 16 x 16 grid, 900 veh/h/road, B = 1024 over 8 GPUs = 128 per GPU (global-memory working set)"),
}


def run(name):
    scen, kw, B, cap, W, K, note = CONFIGS[name]
    if scen is None:
        from pytsc_b200.generators import write_grid_scenario
        scen = write_grid_scenario(tempfile.mkdtemp(prefix="grid16_"), 16, 16, vehicles_per_hour_per_road=900, horizon=3600, seed=0)
    cfg = Config(scen, **kw)
    cs = compile_scenario(cfg, NetworkParser(cfg))
    eng = Engine(cs, B, 0, vehicle_capacity=cap)
    bufs = eng.alloc_outputs(["obs", "reward", "reward_global", "mask", "sim"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    n_ticks = int(cfg.simulator["delta_time"])
    eng.init_program(0)
    for _ in range(W):
        eng.env_step(None, bufs, n_ticks=n_ticks, controller=1, controller_arg=25)
        flush.zero_()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    vsum = torch.zeros((), dtype=torch.float64, device="cuda")
    t0 = time.perf_counter()
    for k in range(K):
        flush.zero_()
        ev[k][0].record()
        eng.env_step(None, bufs, n_ticks=n_ticks, controller=1, controller_arg=25)
        ev[k][1].record()
        vsum += bufs["sim"][:, 0].sum()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    eng.check()
    ms = sum(a.elapsed_time(b) for a, b in ev) / K
    c = eng.counters()
    line = {"config": name, "note": note, "scenario": os.path.basename(str(scen)), "signals": eng.A, "lanes": cs.n_lanes,
            "lane_links": cs.n_lanelinks, "replicas_on_gpu": B, "vehicle_capacity": cap, "steps": K, "warmup": W,
            "ms_per_env_step": ms, "agent_steps_per_s": B * eng.A * 1e3 / ms, "env_steps_per_s": B * 1e3 / ms,
            "mean_running_vehicles": float(vsum.item()) / (K * B), "peak_slots": int(c["n_slots"].max()),
            "final_tick": int(c["tick"][0]), "kernel": eng.kernel_info(), "wall_s": wall}
    eng.close()
    print(json.dumps(line), flush=True)
    return line


if __name__ == "__main__":
    names = sys.argv[1:] or ["jinan", "manhattan", "grid16"]
    out = [run(n) for n in names]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as f:
        json.dump(out, f, indent=1)
