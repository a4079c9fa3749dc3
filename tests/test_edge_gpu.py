"""GPU: edge cases of the batched engine -- capacity overflow is reported and contained, zero-tick
steps leave the state alone, ragged batch sizes (fewer / more replicas than resident blocks, a partly
filled last wave) give every replica the same answer, and replicas are independent of one another."""
import numpy as np
import pytest

from helpers import build_scenario, compare_snapshots

pytestmark = pytest.mark.gpu

KW = dict(signal=dict(observation_space="lane_features", reward_function="max_pressure"))
OUT = ["obs", "reward", "mask", "reward_global", "lane_count", "lane_queued", "sim"]


def test_capacity_overflow_is_reported_and_replica_frozen(cuda_lib):
    import torch
    from pytsc_b200.binding import Engine
    cfg, parser, cs = build_scenario("hangzhou_4_4", **KW)
    eng = Engine(cs, 2, 0, vehicle_capacity=24)       # far below the ~560 vehicles this scenario reaches
    bufs = eng.alloc_outputs(OUT)
    eng.init_program(0)
    for _ in range(40):
        eng.env_step(None, bufs, n_ticks=5, controller=1, controller_arg=25)
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="vehicle_capacity"):
        eng.check()
    c = eng.counters()
    assert (c["tick"] == 200).all()                   # the clock runs on, the replica's vehicles stay put
    frozen = eng.snapshot(0)
    eng.env_step(None, bufs, n_ticks=5, controller=1, controller_arg=25)
    assert compare_snapshots(frozen, eng.snapshot(0)) is None
    assert c["n_slots"].max() <= 24 + 64              # never wrote past the image
    eng.close()
    ok = Engine(cs, 1, 0, vehicle_capacity=640)       # the device is fine afterwards
    ok.init_program(0)
    ok.env_step(None, ok.alloc_outputs(OUT), n_ticks=5, controller=1, controller_arg=25)
    ok.check()
    ok.close()


def test_zero_tick_step_is_a_retrieve(cuda_lib):
    import torch
    from pytsc_b200.binding import Engine
    cfg, parser, cs = build_scenario("hangzhou_4_4", **KW)
    eng = Engine(cs, 2, 0, vehicle_capacity=640)
    b1, b2 = eng.alloc_outputs(OUT), eng.alloc_outputs(OUT)
    eng.init_program(0)
    for _ in range(30):
        eng.env_step(None, b1, n_ticks=5, controller=1, controller_arg=25)
    before = eng.snapshot(1)
    eng.retrieve(b2)
    eng.step(0)
    torch.cuda.synchronize()
    assert compare_snapshots(before, eng.snapshot(1)) is None
    for k in ("obs", "lane_count", "lane_queued", "sim", "reward", "reward_global"):
        assert np.array_equal(b1[k].cpu().numpy(), b2[k].cpu().numpy()), k
    assert eng.counters()["tick"].tolist() == [150, 150]
    eng.check(); eng.close()


@pytest.mark.parametrize("B", [1, 5, 777])
def test_ragged_batch_sizes_agree(cuda_lib, B):
    import torch
    from pytsc_b200.binding import Engine
    cfg, parser, cs = build_scenario("hangzhou_4_4", **KW)
    ref = Engine(cs, 1, 0, vehicle_capacity=640)
    eng = Engine(cs, B, 0, vehicle_capacity=640)
    br, be = ref.alloc_outputs(OUT), eng.alloc_outputs(OUT)
    ref.init_program(0); eng.init_program(0)
    for _ in range(36):
        ref.env_step(None, br, n_ticks=5, controller=1, controller_arg=25)
        eng.env_step(None, be, n_ticks=5, controller=1, controller_arg=25)
    torch.cuda.synchronize()
    for k in OUT:
        a, r = be[k].cpu().numpy(), br[k].cpu().numpy()
        assert (a == r[0:1]).all(), k
    assert compare_snapshots(ref.snapshot(0), eng.snapshot(B - 1)) is None
    ref.check(); eng.check(); ref.close(); eng.close()


def test_replicas_are_independent(cuda_lib):
    """Different action sequences per replica: each replica of the batch equals a single-replica engine
    driven with its own sequence."""
    import torch
    from pytsc_b200.binding import Engine
    cfg, parser, cs = build_scenario("hangzhou_4_4", **KW)
    B, T = 4, 30
    rng = np.random.RandomState(3)
    eng = Engine(cs, B, 0, vehicle_capacity=640)
    be = eng.alloc_outputs(OUT)
    eng.init_program(0)
    eng.retrieve(be)
    singles = []
    for b in range(B):
        e = Engine(cs, 1, 0, vehicle_capacity=640)
        e.init_program(0)
        bb = e.alloc_outputs(OUT)
        e.retrieve(bb)
        singles.append((e, bb))
    torch.cuda.synchronize()
    for t in range(T):
        mask = be["mask"].cpu().numpy()                      # [B, A, P]
        act = np.zeros((B, eng.A), np.int32)
        for b in range(B):
            for a in range(eng.A):
                allowed = np.flatnonzero(mask[b, a])
                act[b, a] = allowed[rng.randint(len(allowed))]
        eng.env_step(torch.from_numpy(act).cuda(), be, n_ticks=5)
        for b, (e, bb) in enumerate(singles):
            e.env_step(torch.from_numpy(act[b:b + 1]).cuda(), bb, n_ticks=5)
        torch.cuda.synchronize()
        for b, (e, bb) in enumerate(singles):
            for k in OUT:
                assert np.array_equal(be[k][b].cpu().numpy(), bb[k][0].cpu().numpy()), (t, b, k)
    assert not np.array_equal(be["obs"][0].cpu().numpy(), be["obs"][1].cpu().numpy())
    for b, (e, bb) in enumerate(singles):
        assert compare_snapshots(e.snapshot(0), eng.snapshot(b)) is None
        e.close()
    eng.check(); eng.close()


def _oracle_for(tmp_path, net, flows, seed=0):
    import json
    from oracle.engine import Engine as OracleEngine
    (tmp_path / "roadnet.json").write_text(json.dumps(net))
    (tmp_path / "flow.json").write_text(json.dumps(flows))
    cfg = dict(dir=str(tmp_path) + "/", roadnetFile="roadnet.json", flowFile="flow.json", interval=1.0, rlTrafficLight=True,
               laneChange=False, seed=seed, saveReplay=False)
    (tmp_path / "cfg.json").write_text(json.dumps(cfg))
    return OracleEngine(str(tmp_path / "cfg.json"))


def _lockstep_custom_flows(tmp_path, flows, ticks, capacity=900, expect_templates=None):
    import torch
    from pytsc_b200 import bundle
    from pytsc_b200.binding import Engine
    from pytsc_b200.scenario import compile_scenario
    from helpers import oracle_engine, signal_inter_indices
    cfg, parser, _ = build_scenario("hangzhou_4_4", **KW)
    cs = compile_scenario(cfg, parser, flows=flows)
    if expect_templates is not None:
        assert cs.n_templates == expect_templates
    orc = _oracle_for(tmp_path, parser.net, flows, seed=cfg.simulator["seed"])
    eng = Engine(cs, 2, 0, vehicle_capacity=capacity)
    inter = signal_inter_indices(parser)
    seen = 0
    for t in range(ticks):
        if t % 5 == 0:
            r = ((t // 30) % 8 + 1) if (t % 30) < 25 else 0
            eng.set_phase(torch.full((2, eng.A), r, dtype=torch.int32, device="cuda"))
            for a in range(eng.A):
                orc.set_tl_phase_idx(inter[a], r)
        orc.next_step()
        eng.step(1)
        if t % 7 == 0 or t == ticks - 1:
            so = orc.snapshot()
            seen = max(seen, len(so["uid"]))
            msg = compare_snapshots(so, eng.snapshot(1))
            assert msg is None, f"tick {t}: {msg}"
    eng.check()
    c = eng.counters()
    assert c["n_running"][0] == orc.get_vehicle_count() and c["n_finished"][0] == orc.get_finished_vehicle_count()
    info = eng.kernel_info()
    eng.close()
    return seen, info


def test_several_vehicle_templates(cuda_lib, tmp_path):
    """Flow files may give every flow its own vehicle parameters: three templates (lengths, accelerations, speeds, gaps,
    headways) interleaved -- the per-vehicle template look-up path of the kernel, in lock-step with the oracle."""
    from pytsc_b200 import bundle
    cfg, parser, _ = build_scenario("hangzhou_4_4", **KW)
    flows = bundle.load_flow(cfg.create_and_save_cityflow_cfg())
    variants = [dict(length=4.0, minGap=2.0, maxSpeed=9.5, maxPosAcc=1.5, usualPosAcc=1.5, headwayTime=1.8),
                dict(length=6.5, minGap=3.0, maxSpeed=12.5, maxNegAcc=5.0, usualNegAcc=3.5, headwayTime=1.2)]
    for k, f in enumerate(flows):
        if k % 3:
            f["vehicle"] = dict(f["vehicle"], **variants[k % 3 - 1])
    seen, info = _lockstep_custom_flows(tmp_path, flows, 450, expect_templates=3)
    assert seen > 150 and info["threads"] in (256, 512)


def test_empty_flow_file(cuda_lib, tmp_path):
    """No vehicles at all: every launch, retrieve and the registered host path still work; all measurements are zero."""
    import torch
    from pytsc_b200.binding import Engine
    from pytsc_b200.scenario import compile_scenario
    cfg, parser, _ = build_scenario("syn_1x1", **KW)
    cs = compile_scenario(cfg, parser, flows=[])
    assert cs.n_vehicles == 0
    eng = Engine(cs, 3, 0, vehicle_capacity=64)
    bufs = eng.alloc_outputs()
    eng.init_program(0)
    for _ in range(20):
        eng.env_step(None, bufs, n_ticks=5, controller=1, controller_arg=25)
    torch.cuda.synchronize()
    eng.check()
    assert int(bufs["lane_count"].sum()) == 0 and float(bufs["sim"][:, 0].sum()) == 0.0 and float(bufs["sim"][:, 1].sum()) == 0.0
    # common/reward.py:110-118: the global reward starts from 1e-6 and subtracts flicker and pressure, both zero here
    assert np.allclose(bufs["reward_global"].cpu().numpy(), 1e-6, rtol=0, atol=1e-9)
    assert eng.snapshot(2)["uid"].size == 0
    eng.close()


def test_continuous_flows(cuda_lib, tmp_path):
    """CityFlow flow entries may emit a vehicle every `interval` seconds between startTime and endTime (the shipped files
    only use single-vehicle entries): Flow::nextStep's counting, fractional intervals included."""
    from pytsc_b200 import bundle
    cfg, parser, _ = build_scenario("hangzhou_4_4", **KW)
    base = bundle.load_flow(cfg.create_and_save_cityflow_cfg())
    routes = []
    for f in base:
        if f["route"] not in routes:
            routes.append(f["route"])
        if len(routes) == 40:
            break
    flows = [dict(vehicle=base[0]["vehicle"], route=r, interval=[3.0, 4.5, 7.0, 2.5][k % 4], startTime=5 * (k % 6), endTime=(-1 if k % 5 == 0 else 300 + 10 * k))
             for k, r in enumerate(routes)]
    seen, _ = _lockstep_custom_flows(tmp_path, flows, 420)
    assert seen > 200
