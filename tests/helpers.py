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
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


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
