"""CityFlow-format replay writer (``cityflow.save_replay``, backends/cityflow/config.py:88-99).  CPU: the line format
over oracle snapshots; GPU: the plugin Simulator writes one line per engine tick."""
import json
import os

import numpy as np
import pytest

from helpers import build_scenario, oracle_engine, signal_inter_indices


def test_replay_lines_from_oracle_snapshots(tmp_path):
    from pytsc_b200.replay import ReplayWriter, roadnet_log
    cfg, parser, cs = build_scenario("syn_1x1")
    orc = oracle_engine(cfg)
    inter = signal_inter_indices(parser)
    w = ReplayWriter(cs, parser.net, str(tmp_path / "replay.txt"), str(tmp_path / "roadnet.json"))
    raw = np.ones(cs.n_signals, np.int32)
    for a in range(cs.n_signals):
        orc.set_tl_phase_idx(inter[a], 1)
    for t in range(120):
        if t == 60:
            raw[:] = 0
            orc.set_tl_phase_idx(inter[0], 0)
        orc.next_step()
        w.log_step(orc.snapshot(), raw)
    w.close()
    lines = open(tmp_path / "replay.txt").read().splitlines()
    assert len(lines) == 120
    snap = orc.snapshot()
    veh, lights = lines[-1].split(";")
    cars = [c for c in veh.split(",") if c]
    assert len(cars) == len(snap["uid"]) > 10
    for c in cars:
        x, y, ang, name, lc, length, width = c.split(" ")
        assert name.startswith("flow_") and lc == "0" and float(length) == 5.0 and float(width) == 2.0
        assert -400 <= float(x) <= 400 and -400 <= float(y) <= 400 and -3.2 <= float(ang) <= 3.2
    roads = [r for r in lights.split(",") if r]
    assert len(roads) == 4                                   # the four roads that end at the one real intersection
    assert all(len(r.split(" ")) == 4 for r in roads)        # road id + three lanes
    # raw phase 0 = the all-red phase that still lets right turns (lane 2) through; phase 1 opens more
    assert {r.split(" ")[3] for r in roads} == {"g"} and any(r.split(" ")[1] == "r" for r in roads)
    first = [r for r in lines[10].split(";")[1].split(",") if r]
    assert sum(x == "g" for r in first for x in r.split(" ")[1:]) > sum(x == "g" for r in roads for x in r.split(" ")[1:])
    net = json.load(open(tmp_path / "roadnet.json"))
    assert net == roadnet_log(parser.net)
    assert len(net["static"]["nodes"]) == 5 and len(net["static"]["edges"]) == 8 and net["static"]["edges"][0]["nLane"] == 3


@pytest.mark.gpu
def test_plugin_writes_replay(cuda_lib, tmp_path):
    from helpers import reference_pytsc
    pytsc = reference_pytsc()
    if pytsc is None:
        pytest.skip("reference pytsc not importable on this machine")
    kw = dict(cityflow=dict(flow_rate_type="constant", save_replay=True, replay_log_file="replay.txt", roadnet_log_file="roadnet.json"),
              signal=dict(observation_space="lane_features", action_space="phase_selection", round_robin=False),
              gpu=dict(n_replicas=2, view_replica=1, vehicle_capacity=512, replay_dir=str(tmp_path)))
    net = pytsc.TrafficSignalNetwork("syn_1x1", "gpu", **kw)
    ref = pytsc.TrafficSignalNetwork("syn_1x1", "gpu", **dict(kw, cityflow=dict(flow_rate_type="constant"),
                                                              gpu=dict(n_replicas=1, vehicle_capacity=512)))
    for t in range(30):
        acts = [int(np.flatnonzero(m)[0]) for m in net.get_action_mask()]
        net.step(acts)
        ref.step(acts)
        assert net.get_observations() == ref.get_observations()      # tick-by-tick launches == the fused env-step
    net.simulator.close_simulator(); ref.simulator.close_simulator()
    lines = open(tmp_path / "replay.txt").read().splitlines()
    assert len(lines) == 150 and lines[-1].count(",") > 10
    assert os.path.exists(tmp_path / "roadnet.json")
