"""CPU: flow-set compilation and the config logic that picks flow files per replica."""
import os
import random

import numpy as np
import yaml

from pytsc_b200 import bundle
from pytsc_b200.backend.config import Config, DisruptedConfig
from pytsc_b200.backend.network_parser import NetworkParser
from pytsc_b200.scenario import compile_scenario, derive_vehicle_capacity


def test_flow_sets_compile_to_disjoint_vehicle_ranges():
    cfg = Config("syn_1x1", cityflow=dict(flow_rate_type="constant"))
    parser = NetworkParser(cfg)
    files = ["syn_1x1__gaussian_500_flows.json", "syn_1x1__gaussian_700_flows.json"]
    cs = compile_scenario(cfg, parser, flow_sets=[cfg.resolve_flow_file(f) for f in files])
    singles = [compile_scenario(cfg, parser, flow_file=cfg.resolve_flow_file(f)) for f in files]
    off = cs.stats["flow_set_off"]
    assert cs.n_flow_sets == 2 and off == [0, singles[0].n_vehicles, singles[0].n_vehicles + singles[1].n_vehicles]
    L = cs.n_lanes
    so = np.asarray(cs.lane_spawn_off).reshape(2, L + 1)
    assert so[0, 0] == 0 and so[0, L] == so[1, 0] == off[1] and so[1, L] == off[2]
    for f in range(2):
        one = singles[f]
        sl = slice(off[f], off[f + 1])
        assert np.array_equal(np.asarray(cs.veh_tick)[sl], np.asarray(one.veh_tick)[: one.n_vehicles])
        assert np.array_equal(np.asarray(cs.veh_priority)[sl], np.asarray(one.veh_priority)[: one.n_vehicles])
        # same routes behind possibly different indices
        seq, seq1 = np.asarray(cs.route_seq), np.asarray(one.route_seq)
        for v in range(0, one.n_vehicles, 37):
            a, b = int(np.asarray(cs.veh_seq_start)[off[f] + v]), int(np.asarray(one.veh_seq_start)[v])
            ra, rb = [], []
            while seq[a] >= 0:
                ra.append(int(seq[a])); a += 1
            while seq1[b] >= 0:
                rb.append(int(seq1[b])); b += 1
            assert ra == rb
        # per-set spawn lists hold exactly the set's vehicles, lane by lane in creation order
        vids = np.asarray(cs.lane_spawn_vid)[so[f, 0]:so[f, L]]
        assert sorted(vids.tolist()) == list(range(off[f], off[f + 1]))
        for l in range(L):
            seg = np.asarray(cs.lane_spawn_vid)[so[f, l]:so[f, l + 1]]
            assert (np.diff(seg) > 0).all()
    assert 64 <= derive_vehicle_capacity(cs) <= max(s.n_vehicles for s in singles)


def test_flow_file_universe_and_draws():
    c = Config("syn_1x1", cityflow=dict(flow_rate_type="random"))
    files = c.flow_file_universe()
    assert len(files) == 9 and all(os.path.exists(c.resolve_flow_file(f)) for f in files)
    random.seed(0)
    exp = [random.choice(files) for _ in range(5)]
    c = Config("syn_1x1", cityflow=dict(flow_rate_type="random"))
    got = []
    for _ in range(5):
        c._set_flow_file()
        got.append(c.flow_file)
    assert got == exp
    s = Config("syn_3x3", cityflow=dict(flow_rate_type="sequential"))
    seq = []
    for _ in range(12):
        s._set_flow_file()
        seq.append(s.flow_file)
    assert seq[:10] == s.simulator["flow_files"] and seq[10:] == s.simulator["flow_files"][:2]
    assert Config("hangzhou_4_4").flow_file_universe() == ["anon_4_4_hangzhou_real.npz"]


def test_disrupted_config(tmp_path):
    """DisruptedConfig (backends/cityflow/config.py:106-175) over a scenario directory laid out the way the
    reference's disrupted scenarios are: <mode>/<domain>/<value>/<flow file>."""
    src = Config("syn_1x1", cityflow=dict(flow_rate_type="constant"))
    d = tmp_path / "syn_1x1_disrupted"
    tree = {"train": {"flow_disrupted": {"500": ["a.npz"], "700": ["b.npz", "c.npz"]}, "link_disrupted": {"0_1": ["d.npz"]}}}
    pick = {"a.npz": "syn_1x1__gaussian_500_flows.json", "b.npz": "syn_1x1__gaussian_700_flows.json",
            "c.npz": "syn_1x1__gaussian_675_flows.json", "d.npz": "syn_1x1__gaussian_600_flows.json"}
    os.makedirs(d)
    import shutil
    shutil.copy(src.cityflow_roadnet_file, d / os.path.basename(src.cityflow_roadnet_file))
    for dom, vals in tree["train"].items():
        for v, fl in vals.items():
            os.makedirs(d / "train" / dom / v)
            for f in fl:
                shutil.copy(src.resolve_flow_file(pick[f]), d / "train" / dom / v / f)
    with open(d / "config.yaml", "w") as f:
        yaml.safe_dump({"cityflow": {"roadnet_file": os.path.basename(src.cityflow_roadnet_file), "flow_rate_type": "random", **tree},
                        "signal": {"action_space": "phase_selection", "round_robin": False}}, f)
    c = DisruptedConfig(str(d), mode="train", disrupted=True)
    assert c.domain_classes == [("flow_disrupted", "500"), ("flow_disrupted", "700"), ("link_disrupted", "0_1")]
    uni = c.flow_file_universe()
    assert len(uni) == 4 and all(os.path.exists(c.resolve_flow_file(f)) for f in uni)
    random.seed(c.simulator["seed"])
    dom = random.choice(c.domains); val = random.choice(c.disrup_values[dom]); ff = random.choice(tree["train"][dom][val])
    c._set_flow_file()
    assert c.flow_file == os.path.join("train", dom, val, ff) and c.current_domain_class == c.domain_classes.index((dom, val))
    c.set_domain_class(("link_disrupted", "0_1"))
    c._set_flow_file()
    assert c.flow_file == os.path.join("train", "link_disrupted", "0_1", "d.npz")
    parser = NetworkParser(c)
    cs = compile_scenario(c, parser, flow_sets=[c.resolve_flow_file(f) for f in uni])
    assert cs.n_flow_sets == 4
