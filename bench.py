#!/usr/bin/env python
"""bench.py -- agent-steps/s of the batched traffic-signal-control env-step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--config hangzhou|jinan|manhattan|grid16] [--fast-forward F]

One "step" = one ``TrafficSignalNetwork.step`` for every replica: phase program (fixed-time controller, green
25 s), delta_time = 5 engine ticks, Retriever reductions, per-signal stats, reward, action mask and
lane-feature observations -- a single launch of ``tsc_step_kernel`` through the C ABI.

Workloads (``--config``; BASELINE.json ``configs``):
  hangzhou   configs[1]  Hangzhou 4x4, max_pressure reward, B = 4096 replicas per GPU          (default: the metric's config)
  jinan      configs[2]  Jinan 3x4, queue reward, through the batched EPyMARL wrapper, B = 2048 per GPU (16384 over 8)
  grid16     configs[3]  generated 16x16 grid (256 signals), 900 veh/h/road, B = 128 per GPU (1024 over 8)
  manhattan  configs[4]  Manhattan 16x3 (48 signals), 3600 s horizon, B = 4096 per GPU

Both arms are measured in the LOADED regime: before warm-up every replica is fast-forwarded, untimed, by
``--fast-forward`` env-steps (default 360 = tick 1800 of the simulated hour) under the same controller; the
reference arm fast-forwards its engines likewise.  ``config.mean_running_vehicles`` reports the load.

Prints ONE JSON line (rank 0).
``value``     device-resident throughput: CUDA events on the launching stream around every launch, L2 flushed
              (256 MiB write) between launches, max over ranks.
``e2e``       the same metric through the public host API (``BatchedTrafficSignalNetwork.step_host`` ->
              ``tsc_env_step_registered``): host policy (numpy) -> actions from page-locked host memory H2D -> ONE
              launch whose replica blocks store compact packets into page-locked host memory -> host threads finish
              the caller's fp32 observation rows / rewards / masks while the launch runs; wall clock, synchronous.
``roofline``  algorithmic bytes per launch (SURVEY.md 8d formula with the measured mean vehicle count) over the mean
              launch duration, against MEASURED_PEAKS.json; ``issue`` = the SM-side bound from the committed ncu capture.
``cpu_baseline`` / ``--impl reference``: the reference's own Python (``baseline/_ref`` pytsc: TrafficSignalNetwork,
              CityFlow backend plugin, FixedTimeController driven the way Evaluate.run does) over the C++ oracle engine
              standing in for the absent ``cityflow`` module, one process per host core
              (kind "reference-python+oracle-engine"); the Python port is the labelled fallback.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GREEN_TIME = 25
METRIC = "agent-steps/sec (B=4096 4x4-grid envs)"
LF = dict(observation_space="lane_features", action_space="phase_selection", round_robin=False)

# name -> workload.  replicas = per GPU; capacity = running vehicles per replica the image is sized for
# (oracle-measured peak under this controller + margin; exceeding it is reported by eng.check()).
CONFIGS = {
    "hangzhou": dict(scenario="hangzhou_4_4", label="BASELINE.json configs[1]",
                     kw=dict(cityflow=dict(flow_file="anon_4_4_hangzhou_real.json", flow_rate_type="constant"),
                             signal=dict(LF, reward_function="max_pressure")),
                     replicas=4096, capacity=640, api="BatchedTrafficSignalNetwork"),
    "jinan": dict(scenario="jinan_3_4", label="BASELINE.json configs[2]: B = 16384 over 8 GPUs = 2048 per GPU",
                  kw=dict(cityflow=dict(flow_rate_type="constant"), signal=dict(LF, reward_function="queue_length")),
                  replicas=2048, capacity=1150, api="BatchedEPyMARLTrafficSignalNetwork"),
    "manhattan": dict(scenario="manhattan_16_3", label="BASELINE.json configs[4]: 3600 s horizon",
                      kw=dict(cityflow=dict(flow_rate_type="constant", episode_limit=3600), signal=dict(LF, reward_function="queue_length")),
                      replicas=4096, capacity=1530, api="BatchedTrafficSignalNetwork"),
    "grid16": dict(scenario=None, label="BASELINE.json configs[3]: 16x16 grid, 900 veh/h/road, B = 1024 over 8 GPUs = 128 per GPU",
                   kw=dict(cityflow=dict(flow_rate_type="constant"), signal=dict(LF, reward_function="max_pressure")),
                   replicas=128, capacity=24000, api="BatchedTrafficSignalNetwork"),
}


def workload(name):
    w = dict(CONFIGS[name])
    w["name"] = name
    if w["scenario"] is None:      # generated grid (the reference shells out to CityFlow's generator, grid_generator.py:39-73)
        from pytsc_b200.generators import write_grid_scenario
        w["scenario"] = write_grid_scenario(tempfile.mkdtemp(prefix="grid16_"), 16, 16, vehicles_per_hour_per_road=900,
                                            horizon=3600, seed=0)
    return w


def build_hash():
    from pytsc_b200 import _build
    return _build.source_hash()[:16]


# ---------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# ---------------------------------------------------------------------------------------
def materialise_reference_scenario(w, root):
    """Write the workload's roadnet / flow as CityFlow JSON plus a config.yaml the way the reference lays its
    scenarios out (scenarios/cityflow/<name>/), under ``root``: baseline/_ref ships code only."""
    import shutil
    import yaml
    from pytsc_b200 import bundle
    from pytsc_b200.backend.config import Config
    cfg = Config(w["scenario"], **w["kw"])
    name = os.path.basename(os.path.normpath(str(w["scenario"])))
    d = os.path.join(root, "cityflow", name)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "roadnet.json"), "w") as f:
        json.dump(bundle.load_roadnet(cfg.cityflow_roadnet_file), f)
    with open(os.path.join(d, "flow.json"), "w") as f:
        json.dump(bundle.load_flow(cfg.create_and_save_cityflow_cfg()), f)
    y = {"cityflow": {k: v for k, v in cfg.simulator.items() if k not in ("roadnet_file", "flow_file", "flow_files")},
         "signal": dict(cfg.signal)}
    y["cityflow"].update(roadnet_file="roadnet.json", flow_file="flow.json", flow_rate_type="constant",
                         roadnet_log_file="roadnet_log_file.json", replay_log_file="replay_log_file.txt", save_replay=False)
    with open(os.path.join(d, "config.yaml"), "w") as f:
        yaml.safe_dump(y, f)
    import pytsc
    os.makedirs(os.path.join(root, "default"), exist_ok=True)
    shutil.copy(os.path.join(os.path.dirname(pytsc.__file__), "scenarios", "default", "config.yaml"),
                os.path.join(root, "default", "config.yaml"))
    return name


def _reference_python_worker(args):
    """One process: the UNMODIFIED reference stack (baseline/_ref pytsc) on one replica -- TrafficSignalNetwork with the
    CityFlow backend plugin, FixedTimeController objects driven as controllers/evaluate.py:112-137 does -- over the oracle
    engine as ``cityflow.Engine``.  Modes: "step" = network.step + mask + observations + local rewards per env-step (what
    an RL loop pulls); "evaluate" = Evaluate.run itself (its _get_actions recomputes all observations once per agent)."""
    w, n_ff, n_steps, mode = args
    import logging
    from pytsc_b200 import compat
    from oracle.engine import Engine as OracleEngine
    compat.install_stubs(engine_factory=OracleEngine)
    where = compat.find_reference_pytsc()
    if where is None:
        raise RuntimeError("reference pytsc not importable")
    logging.disable(logging.CRITICAL)
    import pytsc
    import pytsc.backends.cityflow.config as cf_config
    import pytsc.common.config as base_config
    root = tempfile.mkdtemp(prefix="tsc_refscn_")
    name = materialise_reference_scenario(w, root)
    base_config.CONFIG_DIR = root                               # where the reference looks scenarios up
    cf_config.CONFIG_DIR = os.path.join(root, "cityflow")
    kw = {k: dict(v) for k, v in w["kw"].items()}
    kw.get("cityflow", {}).pop("flow_file", None)
    from pytsc.controllers.evaluate import Evaluate
    ev = Evaluate(name, "cityflow", "fixed_time", add_env_args=kw, add_controller_args={"green_time": GREEN_TIME})
    net = ev.network

    def cheap_actions():      # the controllers' own get_action, without Evaluate._get_actions' per-agent observation rebuild
        return [ev.controllers[ts_id].get_action(None) for ts_id in net.traffic_signals]

    for _ in range(n_ff):          # untimed fast-forward into the loaded regime
        net.step(cheap_actions())
    for _ in range(3):             # untimed: first calls of the getters (lazy properties, caches)
        net.step(cheap_actions())
        net.get_action_mask(); net.get_observations(); net.get_rewards()
    t0 = time.perf_counter()
    if mode == "evaluate":         # Evaluate.run itself (controllers/evaluate.py:71-95)
        ev.run((n_steps + 0.5) * ev.delta_time / 3600.0, output_folder=tempfile.mkdtemp(prefix="tsc_eval_"))
    else:
        for _ in range(n_steps):
            net.step(cheap_actions())
            net.get_action_mask(); net.get_observations(); net.get_rewards()
    dt = time.perf_counter() - t0
    net = ev.network
    return dt, len(net.traffic_signals), net.simulator.step_measurements["sim"]["n_vehicles"]


def _port_worker(args):
    """Fallback: the same loop through oracle/pytsc_port.py (restatement of the reference's Python half)."""
    w, n_ff, n_steps, mode = args
    from oracle.pytsc_port import PortEnv
    env = PortEnv(w["scenario"], **w["kw"])
    for _ in range(n_ff):
        env.step(env.fixed_time_actions(GREEN_TIME))
    t0 = time.perf_counter()
    if mode == "engine":
        for _ in range(n_steps):
            env.engine.next_steps(5)
    else:
        for _ in range(n_steps):
            env.step(env.fixed_time_actions(GREEN_TIME))
            env.get_action_mask(); env.get_observations(); env.get_rewards()
    return time.perf_counter() - t0, env.n_agents, env.engine.get_vehicle_count()


def cpu_throughput(w, n_ff, n_steps, mode="step", procs=None, kind="reference"):
    """Sum of agent-steps/s over `procs` independent processes (one per host core)."""
    import multiprocessing as mp
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    fn = _reference_python_worker if kind == "reference" else _port_worker
    with ctx.Pool(procs) as pool:
        res = pool.map(fn, [(w, n_ff, n_steps, mode)] * procs)
    agents = res[0][1]
    return dict(value=sum(agents * n_steps / r[0] for r in res), cores=procs, wall=max(r[0] for r in res),
                vehicles=float(sum(r[2] for r in res)) / len(res))


def reference_kind():
    """"reference" when the reference package (baseline/_ref) imports on this box, else "port"."""
    try:
        from pytsc_b200 import compat
        if compat.find_reference_pytsc() is not None:
            return "reference"
    except Exception:
        pass
    return "port"


def probe_cityflow():
    import importlib.util
    try:
        return importlib.util.find_spec("cityflow") is not None
    except Exception:
        return False


def cpu_baseline_block(w, n_ff, n_steps):
    from oracle import engine as oracle_engine
    oracle_engine.build()
    kind = reference_kind()
    r = cpu_throughput(w, n_ff, n_steps, "step", kind=kind)
    label = "reference-python+oracle-engine" if kind == "reference" else "port"
    out = {"value": r["value"], "unit": "agent-steps/s", "cores": r["cores"], "kind": label,
           "sample": f"{r['cores']} processes x {n_steps} env-steps after {n_ff} untimed fast-forward steps of {w['name']}, fixed-time "
                     f"green {GREEN_TIME} s: network.step + action mask + observations + local rewards per step, "
                     + ("the unmodified reference Python (baseline/_ref) over the C++ oracle engine as cityflow.Engine"
                        if kind == "reference" else "the Python port over the C++ oracle engine")
                     + f"; {r['wall']:.1f} s wall; mean running vehicles at the end {r['vehicles']:.0f}",
           "real_cityflow_importable": probe_cityflow()}
    if kind == "reference":
        ev = cpu_throughput(w, n_ff, max(8, n_steps // 4), "evaluate", kind=kind)
        out["evaluate_run_value"] = ev["value"]
        out["evaluate_run_note"] = ("Evaluate.run itself (controllers/evaluate.py:71-95,112-124: observations recomputed once "
                                    "per agent inside _get_actions; no mask / reward pulls), same process count")
    eng = cpu_throughput(w, n_ff, n_steps, "engine", kind="port")
    out["engine_only_value"] = eng["value"]
    out["engine_only_note"] = "C++ oracle engine ticks only (no pytsc Python), same process count"
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(args.config)
    n = max(8, min(360, args.steps))
    from oracle import engine as oracle_engine
    oracle_engine.build()
    kind = reference_kind()
    if args.warmup:
        cpu_throughput(w, 0, 4, "step", kind=kind)       # imports, page-in
    r = cpu_throughput(w, args.fast_forward, n, "step", kind=kind)
    label = "reference-python+oracle-engine" if kind == "reference" else "port"
    sample = (f"{r['cores']} processes x {n} env-steps of {w['name']} after {args.fast_forward} untimed fast-forward steps; "
              + ("unmodified reference Python (baseline/_ref) over the C++ oracle engine as cityflow.Engine"
                 if kind == "reference" else "Python port over the C++ oracle engine")
              + " (CityFlow itself is not installable: real_cityflow_importable below)")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "agent-steps/s", "n_gpus": args.gpus,
        "steps": n, "warmup": args.warmup, "ms_per_step": 1e3 * r["wall"] / n, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{w['name']} ({w['label']}): fixed-time green {GREEN_TIME} s, step + mask + observations + rewards",
                   "replicas": r["cores"], "parallelism": f"{r['cores']} host processes", "fast_forward_steps": args.fast_forward,
                   "mean_running_vehicles": r["vehicles"]},
        "cpu_baseline": {"value": r["value"], "unit": "agent-steps/s", "cores": r["cores"], "kind": label, "sample": sample,
                         "real_cityflow_importable": probe_cityflow()},
        "e2e": {"value": r["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    """Pin this process (and the buffers it is about to allocate) to the CPUs of the GPU's NUMA node, as a multi-socket
    deployment would per rank.  Returns (description, previous affinity); silent no-op where the topology is not exposed."""
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


def bind_to_rank_cpu_slice(local_rank, local_world):
    """Several ranks on one host without a NUMA hint: give every rank its own contiguous share of the allowed CPUs, whole
    physical cores where sysfs shows the hyper-thread siblings, so that one rank's policy thread and row-finishing workers
    do not share cores with another rank's.  Returns a description (None: nothing done)."""
    try:
        if os.environ.get("BENCH_NO_AFFINITY") or local_world < 2:
            return None
        allowed = sorted(os.sched_getaffinity(0))
        cores, seen = [], set()
        for cpu in allowed:
            if cpu in seen:
                continue
            sib = {cpu}
            try:
                txt = open(f"/sys/devices/system/cpu/cpu{cpu}/topology/thread_siblings_list").read().strip()
                for part in txt.split(","):
                    lo, _, hi = part.partition("-")
                    sib.update(range(int(lo), int(hi or lo) + 1))
            except Exception:
                pass
            sib &= set(allowed)
            seen |= sib
            cores.append(sorted(sib))
        per = len(cores) // local_world
        if per < 1:
            return None
        cap = int(os.environ.get("BENCH_CORES_PER_RANK", "0"))      # (experiments: a denser host than this one)
        if cap > 0:
            per = min(per, cap)
        mine = [c for core in cores[local_rank * per:(local_rank + 1) * per] for c in core]
        os.sched_setaffinity(0, mine)
        return f"rank slice: {len(mine)} of {len(allowed)} cpus ({per} cores)"
    except Exception as e:
        return f"unchanged ({type(e).__name__})"


class HostFixedTimePolicy:
    """FixedTimeController.get_action (controllers/controllers.py:39-54) for all B x A signals on the host, with its own
    copy of the programs' state: the e2e leg's stand-in for a user's policy.

    The controller is a finite-state machine per signal -- state = (signal, current phase index, time on phase in units of
    delta_time, saturating) -- so the whole B x A batch advances with two table look-ups per step (action, next state).
    tests/test_port.py checks it against the rule written out."""
    T = 64      # time-on-phase slots per phase (saturating: beyond green_time nothing changes)

    def __init__(self, sig_phase_green, sig_n_phases, B, A, green_time, n_ticks):
        import numpy as np
        self.np = np
        green = np.ascontiguousarray(sig_phase_green).reshape(A, -1).astype(bool)
        P, T = green.shape[1], self.T
        assert green_time < (T - 1) * n_ticks
        nph = np.asarray(sig_n_phases, np.int64).reshape(A)
        nxt_state = np.zeros((A, P * T), np.intp)
        action = np.zeros((A, P * T), np.int32)
        for a in range(A):
            for c in range(int(nph[a])):
                for t in range(T):
                    stay = bool(green[a, c]) and t * n_ticks < green_time       # on green for less than green_time
                    n = c if stay else (c + 1) % int(nph[a])
                    # BaseTSProgram.update_current_phase (common/traffic_signal.py:94-109): += delta_time or = delta_time
                    nt = min(t + 1, T - 1) if n == c else 1
                    nxt_state[a, c * T + t] = a * P * T + n * T + nt        # global table index: one take per step
                    action[a, c * T + t] = n
        self.nxt_state, self.action = nxt_state.ravel(), action.ravel()
        self.state0 = np.ascontiguousarray(np.broadcast_to((np.arange(A, dtype=np.intp) * (P * T))[None, :], (B, A)))
        self.state = self.state0.copy()
        self.tmp = np.empty_like(self.state)

    def reset(self):
        self.state[...] = self.state0

    def snapshot(self):
        return self.state.copy()

    def restore(self, s):
        self.state[...] = s

    def act(self, out):
        np = self.np
        # (mode="wrap": numpy buffers `out` under the default mode="raise"; the indices are table entries, always in range)
        np.take(self.action, self.state, out=out, mode="wrap")
        np.take(self.nxt_state, self.state, out=self.tmp, mode="wrap")
        self.state, self.tmp = self.tmp, self.state


def algorithmic_bytes_per_env_step(V, L, K, A, obs_dim, P, n_ticks=5):
    """per tick 40 B per vehicle + 8 B per drivable + 4 B per signal; per env-step
    16 B per lane + per agent (4 obs_dim + reward 4 + mask P + action 4)."""
    return n_ticks * (40.0 * V + 8 * (L + K) + 4 * A) + 16 * L + A * (4 * obs_dim + P + 8)


def ncu_side_table(name, vbar):
    """DRAM bytes and executed warp instructions per launch from the committed ncu capture of this workload
    (profiles/traffic.json, written by tools/make_profile_summary.py) -- only when that capture was taken at a
    comparable load (mean running vehicles within 15 %); otherwise null: a number measured at another load is not
    printed beside this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        t = t.get(name, t if name == "hangzhou" and "dram_bytes_per_launch" in t else None)
        if not t or not t.get("mean_running_vehicles"):
            return None
        if abs(t["mean_running_vehicles"] - vbar) > 0.15 * vbar:
            return None
        return t
    except Exception:
        return None


def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from pytsc_b200 import _build

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the gpu backend has no CPU fallback)")
    torch.cuda.set_device(local)
    affinity, prev_affinity = bind_to_gpu_numa_node(local)
    if prev_affinity is None:
        sliced = bind_to_rank_cpu_slice(local, int(os.environ.get("LOCAL_WORLD_SIZE", str(world))))
        if sliced:
            affinity = sliced
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        _build.build()
    if world > 1:
        dist.barrier()
    from pytsc_b200 import BatchedEPyMARLTrafficSignalNetwork, BatchedTrafficSignalNetwork

    w = workload(args.config)
    B = args.replicas or w["replicas"]
    cap = args.vehicle_capacity or w["capacity"]
    kw = {k: dict(v) for k, v in w["kw"].items()}
    kw["gpu"] = dict(vehicle_capacity=cap)
    if w["api"] == "BatchedEPyMARLTrafficSignalNetwork":      # config 3: the MARL wrapper's API shape (epymarl.py:96-111)
        wrapper = BatchedEPyMARLTrafficSignalNetwork(map_name=w["scenario"], simulator_backend="gpu", n_replicas=B, device=local, **kw)
        env = wrapper.tsc_env
    else:
        wrapper = None
        env = BatchedTrafficSignalNetwork(w["scenario"], n_replicas=B, device=local, **kw)
    eng, cs, cfg = env.engine, env.scenario, env.config
    A, L, K = eng.A, cs.n_lanes, cs.n_lanelinks
    n_ticks = int(cfg.simulator["delta_time"])
    sim_len_steps = int(cfg.simulator["sim_length"]) // n_ticks
    ff = min(args.fast_forward, sim_len_steps - 1)
    names = ["obs", "reward", "reward_global", "mask", "lane_count", "lane_queued", "lane_occupancy", "lane_mean_speed", "sim"]
    bufs = eng.alloc_outputs(names)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    vsum = torch.zeros((), dtype=torch.float64, device="cuda")
    policy = HostFixedTimePolicy(cs.sig_phase_green, cs.sig_n_phases, B, A, GREEN_TIME, n_ticks)
    act_pinned = torch.zeros((B, A), dtype=torch.int32, pin_memory=True)
    act_np = act_pinned.numpy()
    state = {"step": 0}

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()      # started before the fast-forward: nvidia-smi needs a few hundred ms before its first sample
    # ---- untimed fast-forward into the loaded regime; the state (and the host policy's) is kept for both legs ----------
    eng.reset()
    eng.init_program(0)
    policy.reset()
    for _ in range(ff):
        eng.env_step(None, None, n_ticks=n_ticks, controller=1, controller_arg=GREEN_TIME)
        policy.act(act_np)
    torch.cuda.synchronize()
    eng.check()
    loaded_state = eng.save_state(device=True)
    loaded_policy = policy.snapshot()

    def restart(from_loaded):
        if from_loaded:
            eng.load_state(loaded_state)
            policy.restore(loaded_policy)
            state["step"] = ff
        else:                                   # simulator.is_terminated -> a new engine at tick 0 (pytsc/__init__.py:164-176)
            eng.reset()
            eng.init_program(0)
            policy.reset()
            state["step"] = 0

    def one_step():
        if state["step"] == sim_len_steps:
            restart(False)
        eng.env_step(None, bufs, n_ticks=n_ticks, controller=1, controller_arg=GREEN_TIME)
        state["step"] += 1

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident leg ----------------------------------------------------------
    restart(True)
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
    sim_end = bufs["sim"].clone()

    # ---- end-to-end leg: the public host API ------------------------------------------------------------------
    # host policy -> actions (page-locked host memory) -> step_host -> observation rows, rewards, masks, global
    # reward in HOST numpy arrays, every step.  Double-buffered sampling: the rank's B replicas are two environments
    # of B / 2, stepped alternately through step_host_begin / step_host_wait, so that the policy and the row finishing
    # of one half overlap the launch of the other (every half's actions still follow its own previous observations).
    n_half = 1 if (args.e2e_halves < 2 or B < 2) else 2
    sizes = [B] if n_half == 1 else [B // 2, B - B // 2]
    # row-finishing workers: one per CPU of this process (its rank slice, if it has one), split between the halves; they sleep
    # between steps and hand groups of replicas out dynamically, so the policy thread joins in whenever it waits
    my_cpus = len(os.sched_getaffinity(0))
    if "rank slice" not in str(affinity):
        my_cpus //= max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    threads = int(os.environ.get("TSC_B200_HOST_THREADS", "0")) or max(1, min(8, my_cpus))
    threads_per_half = max(1, threads // n_half)
    halves = []
    for hb in sizes:
        if n_half == 1:
            h_env, h_wrap = env, wrapper
        elif w["api"] == "BatchedEPyMARLTrafficSignalNetwork":
            h_wrap = BatchedEPyMARLTrafficSignalNetwork(map_name=w["scenario"], simulator_backend="gpu", n_replicas=hb, device=local, **kw)
            h_env = h_wrap.tsc_env
        else:
            h_wrap, h_env = None, BatchedTrafficSignalNetwork(w["scenario"], n_replicas=hb, device=local, **kw)
        h_pol = HostFixedTimePolicy(cs.sig_phase_green, cs.sig_n_phases, hb, A, GREEN_TIME, n_ticks)
        h_act = torch.zeros((hb, A), dtype=torch.int32, pin_memory=True)
        h = {"env": h_env, "wrap": h_wrap, "eng": h_env.engine, "policy": h_pol, "act": h_act, "act_np": h_act.numpy(), "B": hb,
             "step": 0, "out": h_env.register_host_buffers(threads=threads_per_half)}
        if n_half > 1:      # its own untimed fast-forward into the loaded regime
            h["eng"].reset(); h["eng"].init_program(0); h_pol.reset()
            for _ in range(ff):
                h["eng"].env_step(None, None, n_ticks=n_ticks, controller=1, controller_arg=GREEN_TIME)
                h_pol.act(h["act_np"])
            torch.cuda.synchronize()
            h["eng"].check()
            h["loaded"] = (h["eng"].save_state(device=True), h_pol.snapshot())
        else:
            h["loaded"] = (loaded_state, loaded_policy)
        halves.append(h)
    policy_s = [0.0]

    def half_restart(h, from_loaded):
        if from_loaded:
            h["eng"].load_state(h["loaded"][0]); h["policy"].restore(h["loaded"][1]); h["step"] = ff
        else:
            h["eng"].reset(); h["eng"].init_program(0); h["policy"].reset(); h["step"] = 0

    def half_begin(h):
        if h["step"] == sim_len_steps:
            half_restart(h, False)
        t_p = time.perf_counter()
        h["policy"].act(h["act_np"])
        policy_s[0] += time.perf_counter() - t_p
        h["env"].step_host_begin(h["act_np"], controller="phase_index")

    def half_wait(h):
        h["env"].step_host_wait()
        if h["wrap"] is not None:
            _ = h["out"]["reward_global"] / A          # epymarl.py:106-108: common reward = global / n_agents
        h["step"] += 1

    def e2e_run(n):      # n env-steps of every half, software-pipelined across the halves
        if n <= 0:
            return
        half_begin(halves[0])
        for k in range(n):
            for j in range(1, n_half):
                half_begin(halves[j])
            half_wait(halves[0])
            if k + 1 < n:
                half_begin(halves[0])
            for j in range(1, n_half):
                half_wait(halves[j])

    for h in halves:
        half_restart(h, True)
    torch.cuda.synchronize()
    e2e_run(args.warmup)
    sync_all()
    l0 = sum(h["eng"].launch_count() for h in halves)
    policy_s[0] = 0.0
    t0 = time.perf_counter()
    e2e_run(args.steps)
    sync_all()
    e2e_s = time.perf_counter() - t0
    e2e_launches = sum(h["eng"].launch_count() for h in halves) - l0
    for h in halves:
        h["eng"].check()
    h2d = sum(h["act_np"].nbytes for h in halves)
    d2h = sum(h["eng"].host_packet_bytes() for h in halves)
    host_bytes_finished = sum(v.nbytes for h in halves for v in h["out"].values())
    e2e_reward = float(np.concatenate([h["out"]["reward_global"] for h in halves]).mean())
    # the host arrays must hold what the device leg computed for the same state and actions (both legs ran the same
    # number of steps from the same state under the same rule)
    dev_rows = {k: bufs[k].cpu().numpy() for k in ("obs", "reward", "mask")}
    e2e_matches_device, lo = True, 0
    for h in halves:
        for k in ("obs", "reward", "mask"):
            e2e_matches_device = e2e_matches_device and bool(np.array_equal(h["out"][k], dev_rows[k][lo:lo + h["B"]]))
        lo += h["B"]
    for h in halves:
        if h["env"] is not env:
            h["env"].close()

    # ---- max over ranks; episode metrics all-reduced once (the only collective) ----------
    t = torch.tensor([dev_ms, e2e_s, t_wall], dtype=torch.float64, device="cuda")
    epi = torch.stack([sim_end[:, 1].sum(), sim_end[:, 3].sum(), vsum, torch.tensor(float(B), device="cuda", dtype=torch.float64)])
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(epi, op=dist.ReduceOp.SUM)
    dev_ms, e2e_s, t_wall = [float(x) for x in t.tolist()]
    info = eng.kernel_info()
    host_threads = os.environ.get("TSC_B200_HOST_THREADS", "auto")
    env.close()
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
    side = ncu_side_table(w["name"], vbar)
    issue = None
    if side and side.get("inst_executed_per_launch") and clk and clk.get("sm_mhz"):
        slots = 148 * 4 * clk["sm_mhz"] * 1e6 * (launch_ms / 1e3)       # SM sub-partitions x clock x launch time: one warp instruction each
        issue = {"warp_instructions_per_launch": side["inst_executed_per_launch"], "issue_slots_per_launch": slots,
                 "frac": side["inst_executed_per_launch"] / slots, "source": side.get("source")}
    line = {
        "metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": launch_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{w['name']} ({w['label']}): {A} signals, {L} lanes, {K} lane-links, "
                               f"{cfg.signal['reward_function']} reward, lane_features obs, fixed-time controller green {GREEN_TIME} s, "
                               f"delta_time {n_ticks}, through {w['api']}",
                   "replicas_per_gpu": B, "replicas_total": total_B, "parallelism": f"replica-sharded x{world}, no step-path collective",
                   "l2": "256 MiB flush write between timed launches", "fast_forward_steps": ff,
                   "mean_running_vehicles": vbar, "final_tick": final_tick, "env_steps_per_s": value / A,
                   "engine_ticks_per_s": value / A * n_ticks, "vehicle_capacity": cap,
                   "kernel": {"name": "tsc_step_kernel", **info}, "wall_s_timed_region": t_wall,
                   "host_cpu_affinity": affinity, "build_hash": build_hash()},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": side.get("dram_bytes_per_launch") if side else None,
                     "traffic_source": side.get("source") if side else "no ncu capture at this load committed",
                     "peak_source": peak_src, "algorithmic_bytes_per_env_step": alg, "units_per_launch": B, "issue": issue},
        "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s / args.steps, "mean_global_reward_last_step": e2e_reward,
                "host_policy_ms_per_step": 1e3 * policy_s[0] / args.steps,
                "host_result_bytes_per_step": host_bytes_finished, "host_threads": host_threads,
                "environments": f"{n_half} x {sizes[0]} replicas, {threads_per_half} host workers each" + (", stepped alternately (double-buffered sampling)" if n_half > 1 else ""),
                "matches_device_leg": e2e_matches_device,
                "note": "host fixed-time policy (numpy, inside the timed region) -> actions H2D from page-locked memory -> one launch per environment; "
                        "each replica block stores a compact packet (per-lane queue / occupancy / speed as the row shows them, phase, "
                        "rewards, action bits) into page-locked host memory and raises a flag; host threads finish the fp32 "
                        "observation rows, rewards and masks in the caller's numpy arrays while the launch runs (d2h = packet bytes; "
                        "host_result_bytes = the arrays the caller reads)"},
        "gpu_launches": int(launches),
        "e2e_gpu_launches": int(e2e_launches),
        "clocks": clk,
        "episode": {"mean_average_travel_time_s": float(epi[0] / epi[3]), "finished_vehicles_per_replica": float(epi[1] / epi[3])},
    }
    if world == 1 and not args.no_cpu_baseline:
        if prev_affinity:
            os.sched_setaffinity(0, prev_affinity)      # the CPU baseline gets every host core back
        line["cpu_baseline"] = cpu_baseline_block(w, ff, args.cpu_steps)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="hangzhou", choices=sorted(CONFIGS))
    ap.add_argument("--fast-forward", type=int, default=360,
                    help="untimed env-steps before warm-up (both arms): 360 = tick 1800, the loaded regime")
    ap.add_argument("--replicas", type=int, default=0, help="replicas per GPU (0 = the config's)")
    ap.add_argument("--e2e-halves", type=int, default=2,
                    help="e2e leg: 2 = double-buffered sampling over two environments of B/2 replicas (default), 1 = one synchronous step_host")
    ap.add_argument("--vehicle-capacity", type=int, default=0,
                    help="running vehicles per replica the image is sized for (0 = the config's)")
    ap.add_argument("--cpu-steps", type=int, default=60, help="env-steps per process of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
