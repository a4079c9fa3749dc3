"""GPU: several flow files compiled into one scenario, one per replica (``tsc_reset_flows``) -- the batched
form of ``flow_rate_type: random | sequential`` (``pytsc/backends/cityflow/config.py:63-76``) and of
``DisruptedConfig`` (``:106-175``).  Each replica must reproduce the oracle engine run on its own flow file."""
import numpy as np
import pytest

from helpers import build_scenario, compare_snapshots, oracle_engine, signal_inter_indices
from pytsc_b200 import bundle
from pytsc_b200.backend.config import Config
from pytsc_b200.backend.network_parser import NetworkParser
from pytsc_b200.scenario import compile_scenario

pytestmark = pytest.mark.gpu

FILES = ["syn_1x1__gaussian_500_flows.json", "syn_1x1__gaussian_700_flows.json", "syn_1x1__gaussian_600_flows.json"]


def _shift(snap, off):
    s = dict(snap)
    s["uid"] = snap["uid"] - off
    return s


def test_two_flow_sets_in_one_batch(cuda_lib):
    import torch
    from pytsc_b200.binding import Engine
    cfg = Config("syn_1x1", cityflow=dict(flow_rate_type="constant"))
    parser = NetworkParser(cfg)
    cs = compile_scenario(cfg, parser, flow_sets=[cfg.resolve_flow_file(f) for f in FILES])
    assert cs.n_flow_sets == 3
    off = cs.stats["flow_set_off"]
    orcs = []
    for f in FILES:
        c1, _, _ = build_scenario("syn_1x1", cityflow=dict(flow_file=f))
        orcs.append(oracle_engine(c1))
    B = 7
    assign = np.array([0, 1, 2, 1, 0, 2, 1], np.int32)
    eng = Engine(cs, B, 0, vehicle_capacity=512)
    eng.reset(flow_sets=assign)
    inter = signal_inter_indices(parser)
    bufs = eng.alloc_outputs(["sim"])
    for t in range(600):
        if t % 5 == 0:
            k = (t // 30) % 8
            r = (k + 1) if (t % 30) < 25 else 0
            eng.set_phase(torch.full((B, eng.A), r, dtype=torch.int32, device="cuda"))
            for o in orcs:
                o.set_tl_phase_idx(inter[0], r)
        for o in orcs:
            o.next_step()
        eng.step(1)
        if t % 20 == 19:
            snaps = [o.snapshot() for o in orcs]
            for b in range(B):
                msg = compare_snapshots(snaps[assign[b]], _shift(eng.snapshot(b), off[assign[b]]))
                assert msg is None, f"tick {t} replica {b} (flow set {assign[b]}): {msg}"
    eng.retrieve(bufs)
    sim = bufs["sim"].cpu().numpy()
    for b in range(B):
        o = orcs[assign[b]]
        assert sim[b, 0] == o.get_vehicle_count() and sim[b, 3] == o.get_finished_vehicle_count()
        assert sim[b, 1] == pytest.approx(o.get_average_travel_time(), rel=1e-12)
    # selected replicas restart on other flow sets while the rest run on
    eng.reset_replicas([1, 4], flow_sets=[2, 1])
    c = eng.counters()
    assert c["tick"][1] == 0 and c["tick"][4] == 0 and c["tick"][0] == 600
    eng.check()
    with pytest.raises(Exception):
        eng.reset(flow_sets=np.full(B, 3, np.int32))
    eng.close()


def test_batched_env_draws_flow_files_like_the_reference(cuda_lib):
    """flow_rate_type random: one ``random.choice(flow_files)`` per replica per engine restart, in the reference's
    stream (``random.seed(seed)``, backends/cityflow/config.py:41,63-76); every replica then equals the oracle on
    its file."""
    import random
    from pytsc_b200 import BatchedTrafficSignalNetwork
    kw = dict(cityflow=dict(flow_rate_type="random"), signal=dict(observation_space="lane_features"), gpu=dict(vehicle_capacity=512))
    env = BatchedTrafficSignalNetwork("syn_1x1", n_replicas=6, **kw)
    files = env.config.simulator["flow_files"]
    assert len(files) == 9 and env.scenario.n_flow_sets == 9
    random.seed(env.config.simulator["seed"])
    expect = [files.index(random.choice(files)) for _ in range(6)]
    assert list(env.flow_sets) == expect
    for _ in range(40):
        env.step(controller="fixed_time", green_time=25)
    env.check()
    sim = env.out["sim"].cpu().numpy()
    off = env.scenario.stats["flow_set_off"]
    for b in (0, 3, 5):
        f = files[expect[b]]
        c1, p1, cs1 = build_scenario("syn_1x1", cityflow=dict(flow_file=f))
        orc = oracle_engine(c1)
        from test_episode_gpu import HostFixedTime
        host = HostFixedTime(cs1, orc, signal_inter_indices(p1), 5)
        for _ in range(40):
            host.env_step()
        assert compare_snapshots(orc.snapshot(), _shift(env.engine.snapshot(b), off[expect[b]])) is None, b
        assert sim[b, 3] == orc.get_finished_vehicle_count()
    env.close()
