#!/usr/bin/env python
"""bench.py -- agent-steps/s of the batched traffic-signal-control env-step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one ``TrafficSignalNetwork.step`` for every replica: phase program
(fixed-time controller, green 25 s), delta_time = 5 engine ticks, Retriever
reductions, per-signal stats, pressure reward, action mask and lane-feature
observations -- a single launch of ``tsc_step_kernel`` through the C ABI.
Workload: BASELINE.json configs[1], Hangzhou 4x4 (16 signals, 240 lanes, 576
lane-links, 2983 vehicles/h), max_pressure reward, B = 4096 replicas per GPU.

Prints ONE JSON line (rank 0).  ``value``: device-resident throughput, CUDA
events on the launching stream around every launch, L2 flushed between
launches, max over ranks.  ``e2e``: the same metric through ``tsc_env_step_host``
with pinned HOST buffers (actions in; observations, rewards, masks out), wall
clock with a synchronize on both sides.  ``roofline``: algorithmic bytes per
launch (SURVEY.md 8d formula, with the measured mean vehicle count) over the
mean launch duration, against MEASURED_PEAKS.json.  ``cpu_baseline``: the CPU
port (C++ oracle engine + Python port of pytsc's hot path) on all host cores.

``--impl reference`` times only that CPU port (the reference's engine,
CityFlow, is a third-party module that cannot be installed here).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCENARIO = "hangzhou_4_4"
SCENARIO_KW = dict(
    cityflow=dict(flow_file="anon_4_4_hangzhou_real.json", flow_rate_type="constant"),
    signal=dict(observation_space="lane_features", reward_function="max_pressure",
                action_space="phase_selection", round_robin=False),
)
GREEN_TIME = 25
METRIC = "agent-steps/sec (B=4096 4x4-grid envs)"


# ---------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# ---------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One process: fixed-time control of one Hangzhou replica for `n_steps` env-steps
    through the full port (step + mask + observations + local rewards)."""
    n_steps, engine_only = args
    from oracle.pytsc_port import PortEnv
    env = PortEnv(SCENARIO, **SCENARIO_KW)
    t0 = time.perf_counter()
    if engine_only:
        for _ in range(n_steps):
            env.engine.next_steps(5)
    else:
        for _ in range(n_steps):
            acts = env.fixed_time_actions(GREEN_TIME)
            env.step(acts)
            env.get_action_mask()
            env.get_observations()
            env.get_rewards()
    return time.perf_counter() - t0, env.n_agents, env.step_measurements["sim"]["n_vehicles"] if not engine_only else 0


def cpu_port_throughput(n_steps, procs=None, engine_only=False):
    """Sum of agent-steps/s over `procs` independent processes (one per host core)."""
    import multiprocessing as mp
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(n_steps, engine_only)] * procs)
    agents = res[0][1]
    total = sum(agents * n_steps / r[0] for r in res)
    return total, procs, max(r[0] for r in res)


def cpu_baseline_block(n_steps):
    from oracle import engine as oracle_engine
    oracle_engine.build()
    v, cores, wall = cpu_port_throughput(n_steps)
    ve, _, _ = cpu_port_throughput(n_steps, engine_only=True)
    return {
        "value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port",
        "sample": f"{cores} processes x {n_steps} env-steps ({5 * n_steps} s simulated) of {SCENARIO}, fixed-time, "
                  f"step+mask+obs+rewards through the Python port over the C++ oracle engine; {wall:.1f} s wall",
        "engine_only_value": ve,
        "engine_only_note": "same processes, C++ oracle engine ticks only (no pytsc Python glue)",
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = max(8, min(720, args.steps))
    # warm-up: build the oracle, import, page in
    from oracle import engine as oracle_engine
    oracle_engine.build()
    for _ in range(min(args.warmup, 1)):
        cpu_port_throughput(8)
    t0 = time.perf_counter()
    v, cores, wall = cpu_port_throughput(n)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": n, "warmup": args.warmup, "ms_per_step": 1e3 * wall / n, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{SCENARIO}: 16 signals, max_pressure reward, lane_features obs, fixed-time green {GREEN_TIME} s",
                   "replicas": cores, "parallelism": f"{cores} host processes"},
        "cpu_baseline": {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{cores} processes x {n} env-steps of {SCENARIO} through the Python port over the C++ "
                                   f"oracle engine (CityFlow itself is not installable)"},
        "e2e": {"value": v, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the GPU is under the bench load.
    Started before the warm-up steps so that short runs still get samples; the samples that fall inside
    the timed region are the ones reported whenever there are at least three of them."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.t_timed = index, [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def mark_timed_region(self):
        self.t_timed = time.perf_counter()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()
        timed = [r for t, r in self.rows if self.t_timed is not None and t >= self.t_timed]
        window = "timed region"
        if len(timed) < 3:
            timed, window = [r for _, r in self.rows], "warm-up + timed region (timed region shorter than three samples)"
        sm = sorted(int(r[0]) for r in timed if r and r[0].isdigit())
        mx = [int(r[1]) for r in timed if len(r) > 1 and r[1].isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in timed)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "window": window}


# ---------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(index):
    """Pin this process (and the pinned buffers it is about to allocate) to the CPUs of the GPU's NUMA
    node, as a multi-socket deployment would per rank: host <-> device copies then stay on the local PCIe
    root.  Returns (description, previous affinity); silent no-op where the topology is not exposed."""
    try:
        if os.environ.get("BENCH_NO_AFFINITY"):
            return "unchanged (BENCH_NO_AFFINITY)", None
        import torch
        p = torch.cuda.get_device_properties(index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        prev = os.sched_getaffinity(0)
        cpus &= prev
        if node < 0 or not cpus or cpus == prev:
            return f"unchanged (numa node {node})", None
        os.sched_setaffinity(0, cpus)
        return f"numa node {node} of GPU {bdf}: {len(cpus)} of {len(prev)} cpus", prev
    except Exception as e:      # no sysfs, no permission, old torch ...
        return f"unchanged ({type(e).__name__})", None


class HostFixedTimePolicy:
    """FixedTimeController.get_action (controllers/controllers.py:39-54) for all B x A signals on the
    host, with its own copy of the programs' state: the e2e leg's stand-in for a user's policy.

    The controller is a finite-state machine per signal -- state = (current phase index, time on phase in
    units of delta_time, saturating) -- so the whole B x A batch advances with two table look-ups per
    step (action, next state).  tests/test_port.py checks it against the rule written out."""
    T = 64      # time-on-phase slots per phase (saturating: beyond green_time nothing changes)

    def __init__(self, sig_phase_green, sig_n_phases, B, A, green_time, n_ticks):
        import numpy as np
        self.np = np
        green = np.ascontiguousarray(sig_phase_green).reshape(A, -1).astype(bool)
        P, T = green.shape[1], self.T
        assert green_time < (T - 1) * n_ticks
        nph = np.asarray(sig_n_phases, np.int64).reshape(A)
        nxt_state = np.zeros((A, P * T), np.uint16)
        action = np.zeros((A, P * T), np.int32)
        for a in range(A):
            for c in range(int(nph[a])):
                for t in range(T):
                    stay = bool(green[a, c]) and t * n_ticks < green_time       # on green for less than green_time
                    n = c if stay else (c + 1) % int(nph[a])
                    # BaseTSProgram.update_current_phase (common/traffic_signal.py:94-109): += delta_time or = delta_time
                    nt = min(t + 1, T - 1) if n == c else 1
                    nxt_state[a, c * T + t] = n * T + nt
                    action[a, c * T + t] = n
        self.nxt_state, self.action = nxt_state.ravel(), action.ravel()
        self.base = (np.arange(A, dtype=np.intp) * (P * T))[None, :]
        self.state = np.zeros((B, A), np.uint16)
        self.idx = np.empty((B, A), np.intp)

    def reset(self):
        self.state.fill(0)

    def act(self, out):
        np = self.np
        np.add(self.base, self.state, out=self.idx)
        np.take(self.action, self.idx, out=out, mode="clip")
        np.take(self.nxt_state, self.idx, out=self.state, mode="clip")


def algorithmic_bytes_per_env_step(V, L, K, A, obs_dim, P, n_ticks=5):
    """per tick 40 B per vehicle + 8 B per drivable + 4 B per signal; per env-step
    16 B per lane + per agent (4 obs_dim + reward 4 + mask P + action 4)."""
    return n_ticks * (40.0 * V + 8 * (L + K) + 4 * A) + 16 * L + A * (4 * obs_dim + P + 8)


def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from pytsc_b200 import _build
    from pytsc_b200.backend.config import Config
    from pytsc_b200.backend.network_parser import NetworkParser
    from pytsc_b200.binding import Engine
    from pytsc_b200.scenario import compile_scenario

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the gpu backend has no CPU fallback)")
    torch.cuda.set_device(local)
    affinity, prev_affinity = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        _build.build()
    if world > 1:
        dist.barrier()

    cfg = Config(SCENARIO, **SCENARIO_KW)
    parser = NetworkParser(cfg)
    cs = compile_scenario(cfg, parser)
    B = args.replicas
    eng = Engine(cs, B, local, vehicle_capacity=args.vehicle_capacity)
    A, L, K = eng.A, cs.n_lanes, cs.n_lanelinks
    n_ticks = int(cfg.simulator["delta_time"])
    sim_len_steps = int(cfg.simulator["sim_length"]) // n_ticks
    names = ["obs", "reward", "reward_global", "mask", "lane_count", "lane_queued", "lane_occupancy",
             "lane_mean_speed", "sim"]
    bufs = eng.alloc_outputs(names)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    vsum = torch.zeros((), dtype=torch.float64, device="cuda")
    state = {"step": 0}

    def restart():
        eng.reset()
        eng.init_program(0)
        state["step"] = 0

    def one_step():
        if state["step"] == sim_len_steps:      # simulator.is_terminated -> restart (pytsc/__init__.py:164-176)
            restart()
        eng.env_step(None, bufs, n_ticks=n_ticks, controller=1, controller_arg=GREEN_TIME)
        state["step"] += 1

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident leg ----------------------------------------------------------
    restart()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        one_step()
        flush.zero_()
    sync_all()
    launches0 = eng.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sync_all()
    clocks.mark_timed_region()
    t_wall = time.perf_counter()
    for k in range(args.steps):
        flush.zero_()                      # evict the replica images and last step's outputs from L2
        ev[k][0].record()
        one_step()
        ev[k][1].record()
        vsum += bufs["sim"][:, 0].sum()
    sync_all()
    t_wall = time.perf_counter() - t_wall
    launches = eng.launch_count() - launches0
    clk = clocks.stop() if rank == 0 else None
    eng.check()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    vbar = float(vsum.item()) / (args.steps * B)
    final_tick = int(eng.counters()["tick"][0])

    # ---- end-to-end leg: host actions in, host observations / rewards / masks out ----------
    pin = dict(pin_memory=True)
    h_act = torch.zeros((B, A), dtype=torch.int32, **pin)
    h_obs = torch.empty((B, A, eng.dims["obs_dim"]), dtype=torch.float32, **pin)
    h_rew = torch.empty((B, A), dtype=torch.float32, **pin)
    h_mask = torch.empty((B, A, eng.dims["n_actions"]), dtype=torch.uint8, **pin)
    h_rg = torch.empty((B,), dtype=torch.float32, **pin)
    policy = HostFixedTimePolicy(cs.sig_phase_green, cs.sig_n_phases, B, A, GREEN_TIME, n_ticks)
    act_np = h_act.numpy()

    def host_policy():
        policy.act(act_np)

    def e2e_restart():
        restart()
        policy.reset()
        torch.cuda.synchronize()

    policy_s = [0.0]

    def e2e_step():
        if state["step"] == sim_len_steps:
            e2e_restart()
        t_p = time.perf_counter()
        host_policy()
        policy_s[0] += time.perf_counter() - t_p
        eng.env_step_host(h_act.numpy(), obs=h_obs.numpy(), reward=h_rew.numpy(), mask=h_mask.numpy(),
                          reward_global=h_rg.numpy(), n_ticks=n_ticks)
        state["step"] += 1

    e2e_restart()
    for _ in range(args.warmup):
        e2e_step()
    sync_all()
    l0 = eng.launch_count()
    policy_s[0] = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    sync_all()
    e2e_s = time.perf_counter() - t0
    e2e_launches = eng.launch_count() - l0
    eng.check()
    h2d = h_act.numel() * 4
    d2h = h_obs.numel() * 4 + h_rew.numel() * 4 + h_mask.numel() + h_rg.numel() * 4
    e2e_reward = float(h_rg.mean())

    # ---- max over ranks; episode metrics all-reduced once (the only collective) ----------
    t = torch.tensor([dev_ms, e2e_s, t_wall], dtype=torch.float64, device="cuda")
    sim = bufs["sim"]
    epi = torch.stack([sim[:, 1].sum(), sim[:, 3].sum(), vsum, torch.tensor(float(B), device="cuda", dtype=torch.float64)])
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(epi, op=dist.ReduceOp.SUM)
    dev_ms, e2e_s, t_wall = [float(x) for x in t.tolist()]
    info = eng.kernel_info()
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return

    total_B = B * world
    value = total_B * A * args.steps / (dev_ms / 1e3)
    e2e_value = total_B * A * args.steps / e2e_s
    alg = algorithmic_bytes_per_env_step(vbar, L, K, A, eng.dims["obs_dim"], eng.dims["n_actions"], n_ticks)
    peaks, peak_src = {}, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        peak_src = "MEASURED_PEAKS.json"
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    launch_ms = dev_ms / args.steps
    achieved = alg * B / (launch_ms / 1e3) / 1e9          # per GPU: one launch handles this rank's B replicas
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": launch_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{SCENARIO} (BASELINE.json configs[1]): 16 signals, 240 lanes, 576 lane-links, "
                               f"anon_4_4_hangzhou_real flows, max_pressure reward, lane_features obs, "
                               f"fixed-time controller green {GREEN_TIME} s, delta_time 5",
                   "replicas_per_gpu": B, "replicas_total": total_B, "parallelism": f"replica-sharded x{world}, no step-path collective",
                   "l2": "256 MiB flush write between timed launches", "mean_running_vehicles": vbar,
                   "final_tick": final_tick, "env_steps_per_s": value / A, "engine_ticks_per_s": value / A * n_ticks,
                   "kernel": {"name": "tsc_step_kernel", **info}, "wall_s_timed_region": t_wall,
                   "host_cpu_affinity": affinity},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_env_step": alg, "units_per_launch": B},
        "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s / args.steps, "mean_global_reward_last_step": e2e_reward,
                "host_policy_ms_per_step": 1e3 * policy_s[0] / args.steps,
                "note": "host fixed-time policy (numpy, inside the timed region) -> actions H2D -> chunked launches, "
                        "observation rows D2H while the next chunk is stepped -> rewards / masks D2H"},
        "gpu_launches": int(launches),
        "e2e_gpu_launches": int(e2e_launches),
        "clocks": clk,
        "episode": {"mean_average_travel_time_s": float(epi[0] / epi[3]), "finished_vehicles_per_replica": float(epi[1] / epi[3])},
    }
    if world == 1 and not args.no_cpu_baseline:
        if prev_affinity:
            os.sched_setaffinity(0, prev_affinity)      # the CPU baseline gets every host core back
        line["cpu_baseline"] = cpu_baseline_block(args.cpu_steps)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=648)
    ap.add_argument("--warmup", type=int, default=72)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--replicas", type=int, default=4096, help="replicas per GPU")
    ap.add_argument("--vehicle-capacity", type=int, default=640,
                    help="running vehicles per replica the shared-memory image is sized for (peak on this workload: 561)")
    ap.add_argument("--cpu-steps", type=int, default=360, help="env-steps per process of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
