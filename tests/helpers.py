"""Shared helpers for the test-suite (test infrastructure only)."""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pytsc_b200 import bundle  # noqa: E402
from pytsc_b200.backend.config import Config  # noqa: E402
from pytsc_b200.backend.network_parser import NetworkParser  # noqa: E402
from pytsc_b200.scenario import compile_scenario  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
_tmp_dirs = {}


def scenario_json_dir(config) -> str:
    """Write roadnet/flow of ``config`` as plain CityFlow JSON (the oracle reads
    JSON like the real engine) and return the directory."""
    key = (config.dir, config.simulator["roadnet_file"], config.simulator["flow_file"])
    if key not in _tmp_dirs:
        d = tempfile.mkdtemp(prefix="tsc_oracle_")
        with open(os.path.join(d, "roadnet.json"), "w") as f:
            json.dump(bundle.load_roadnet(config.cityflow_roadnet_file), f)
        with open(os.path.join(d, "flow.json"), "w") as f:
            json.dump(bundle.load_flow(config.create_and_save_cityflow_cfg()), f)
        _tmp_dirs[key] = d
    return _tmp_dirs[key]


def oracle_engine(config):
    from oracle.engine import Engine
    d = scenario_json_dir(config)
    cfg = dict(dir=d + os.sep, roadnetFile="roadnet.json", flowFile="flow.json",
               interval=config.simulator["interval"], rlTrafficLight=True, laneChange=False,
               seed=config.simulator["seed"], saveReplay=False)
    fn = os.path.join(d, "engine_cfg.json")
    with open(fn, "w") as f:
        json.dump(cfg, f)
    return Engine(fn)


def build_scenario(name, **kwargs):
    # the fixed flow file of the scenario unless the test asks for the reference's random / sequential draw
    # (several shipped config.yaml files say flow_rate_type: random)
    kwargs = {k: dict(v) if isinstance(v, dict) else v for k, v in kwargs.items()}
    kwargs.setdefault("cityflow", {}).setdefault("flow_rate_type", "constant")
    cfg = Config(name, **kwargs)
    parser = NetworkParser(cfg)
    return cfg, parser, compile_scenario(cfg, parser)


def signal_inter_indices(parser):
    """all-intersection index of every signal (agent order)."""
    ids = [it["id"] for it in parser.intersections]
    return [ids.index(t) for t in parser.traffic_signals]


def compare_snapshots(so, sg, tol=0.0):
    """Oracle snapshot vs GPU snapshot (both drivable-major, front to back)."""
    if len(so["uid"]) != len(sg["uid"]):
        return f"vehicle count {len(so['uid'])} != {len(sg['uid'])}"
    if not np.array_equal(so["uid"], sg["uid"]):
        i = int(np.argmax(so["uid"] != sg["uid"]))
        return f"vehicle order differs at {i}: {so['uid'][i]} vs {sg['uid'][i]}"
    if not np.array_equal(so["drivable"], sg["drivable"]):
        i = int(np.argmax(so["drivable"] != sg["drivable"]))
        return f"drivable differs for uid {so['uid'][i]}: {so['drivable'][i]} vs {sg['drivable'][i]}"
    for k in ("distance", "speed"):
        d = np.abs(so[k] - sg[k])
        if d.size and d.max() > tol:
            i = int(np.argmax(d))
            return f"{k} differs for uid {so['uid'][i]}: {so[k][i]!r} vs {sg[k][i]!r}"
    return None


# ---- golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py) ----
def golden_cases():
    """Step fixtures (random-over-mask actions); the closed-loop controller fixtures are ``controller_cases()``."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith(("ctl_", "trips_")))


def controller_cases():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith("ctl_"))


def load_golden(name):
    import ast
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        g = {k: z[k] for k in z.files}
    g["scenario"] = str(g["scenario"])
    g["kwargs"] = ast.literal_eval(str(g["kwargs"]))
    return g


def golden_scenario(g, **gpu):
    """Compile the scenario a golden case was recorded on."""
    kw = {k: dict(v) for k, v in g["kwargs"].items()}
    if gpu:
        kw["gpu"] = gpu
    cfg, parser, cs = build_scenario(g["scenario"], **kw)
    assert cs.lane_ids == [str(x) for x in g["lane_ids"]]
    assert cs.signal_ids == [str(x) for x in g["signal_ids"]]
    return cfg, parser, cs


# ---- a CPU stand-in for binding.Engine (host-logic tests only) ---------------------------
class FakeDeviceEngine:
    """Test double with the ``binding.Engine`` surface the plugin classes use
    (``pytsc_b200/backend/simulator.py``), computing its outputs with the CPU
    oracle.  It lets the host-side plugin logic -- Simulator, Retriever,
    TrafficSignal, MetricsParser, ``register()`` -- be exercised on a machine
    without a GPU.  Never used by the product."""

    def __init__(self, scenario, n_replicas, device=0, vehicle_capacity=0, port_kwargs=None, scenario_name=None):
        import torch
        from oracle.pytsc_port import PortEnv
        self.torch = torch
        self.device = torch.device("cpu")
        self.scenario = scenario
        self.port = PortEnv(scenario_name, **(port_kwargs or {}))
        cs = scenario
        self.B, self.L, self.A = n_replicas, cs.n_lanes, cs.n_signals
        self.dims = dict(B=n_replicas, L=cs.n_lanes, A=cs.n_signals, obs_dim=cs.obs_dim, state_dim=cs.state_dim,
                         n_actions=cs.n_actions, n_in=cs.n_in_total, n_out=cs.n_out_total, vis=cs.visibility)
        self._launches = 0

    def alloc_outputs(self, names=None):
        from pytsc_b200.binding import OUTPUT_SPECS
        t = self.torch
        return {n: t.zeros(OUTPUT_SPECS[n][0](self.dims), dtype=getattr(t, OUTPUT_SPECS[n][1])) for n in (names or OUTPUT_SPECS)}

    def init_program(self, phase_index=0):
        pass          # PortEnv starts every signal on phase 0 already

    def reset(self):
        raise NotImplementedError

    def step(self, n_ticks=1):
        self.port.engine.next_steps(n_ticks)
        self._launches += 1

    def retrieve(self, bufs):
        p = self.port
        p.retrieve_step_measurements()
        for s in p.signals.values():
            p._update_stats(s)
        self._fill(bufs)
        self._launches += 1

    def env_step(self, actions, bufs, n_ticks=5, controller=0, controller_arg=0):
        assert n_ticks == self.port.config.simulator["delta_time"]
        acts = [int(a) for a in actions[0].tolist()]
        self.port.step(acts, phase_indices=(controller == 2))
        self._fill(bufs)
        self._launches += 1

    def _fill(self, bufs):
        t, p, cs = self.torch, self.port, self.scenario
        lm = p.step_measurements["lane"]
        ids = cs.lane_ids
        sig = list(p.signals.values())
        vals = {
            "lane_count": [lm[l]["n_vehicles"] for l in ids],
            "lane_queued": [lm[l]["n_queued"] for l in ids],
            "lane_occupancy": [float(lm[l]["occupancy"]) for l in ids],
            "lane_mean_speed": [float(lm[l]["mean_speed"]) for l in ids],
            "lane_meas64": [[float(lm[l]["occupancy"]), float(lm[l]["mean_speed"])] for l in ids],
            "pos_in": [s.inc_position_matrices[l] for s in sig for l in s.incoming_lanes],
            "pos_out": [s.out_position_matrices[l] for s in sig for l in s.outgoing_lanes],
            "sig_stats64": [[s.n_queued, s.occupancy, s.mean_speed, s.mean_delay, s.outgoing_occupancy, s.pressure,
                             s.norm_time_on_phase, s.current_phase_index] for s in sig],
        }
        ids = list(p.signals.keys())      # MetricsParser.density_map (backends/cityflow/metrics.py:170-199), written out
        nl = p.parser.neighbors_lanes if hasattr(p, "parser") else p.parsed_network.neighbors_lanes
        dm = np.zeros((len(ids), len(ids)))
        for i, ti in enumerate(ids):
            for j, tj in enumerate(ids):
                ls = (nl.get(ti) or {}).get(tj)
                if ls:
                    dm[i, j] = np.clip(sum(lm[l]["occupancy"] for l in ls) / len(ls), 0, 1)
        adj = np.asarray((p.parser if hasattr(p, "parser") else p.parsed_network).adjacency_matrix, np.float64)
        vals["density_map"] = (dm + dm.T) / 2 + 1e-6 * adj
        sm = p.step_measurements["sim"]
        vals["sim"] = [sm["n_vehicles"], sm["average_travel_time"], sm["time_step"], p.engine.get_finished_vehicle_count()]
        st = p.step_stats()
        flick = float(np.mean([bool(s.phase_changed) for s in sig]))
        vals["metrics"] = [st["n_queued"], st["mean_speed"], st["mean_delay"], st["density"], st["pressure"],
                           st["network_flow"], flick, float(p.norm_mean_speed)]
        for k, buf in bufs.items():
            if k in vals:
                buf[:] = t.as_tensor(np.asarray(vals[k]), dtype=buf.dtype)

    def check(self):
        pass

    def close(self):
        pass

    def launch_count(self):
        return self._launches


def reference_pytsc():
    """The reference package with the gpu backend registered, or None when it is
    not importable on this machine (installed, baseline/_ref, or /root/reference)."""
    try:
        import logging
        import pytsc_b200
        p = pytsc_b200.register()
        logging.disable(logging.CRITICAL)
        return p
    except ImportError:
        return None
