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
