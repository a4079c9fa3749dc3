"""GPU: engine state snapshot / restore and per-replica reset (the batched counterpart
of CityFlow's ``engine.snapshot()`` / ``engine.load()`` and of ``engine.reset()`` for one replica).

Properties checked (no oracle needed: the engine is deterministic and already pinned to the oracle by
test_engine_gpu.py): restore-then-replay reproduces the first run bit for bit, and a replica reset in
mid-episode walks the same trajectory as a fresh engine while its neighbours run on undisturbed."""
import numpy as np
import pytest

from helpers import build_scenario, compare_snapshots

pytestmark = pytest.mark.gpu

KW = dict(signal=dict(observation_space="lane_features", reward_function="max_pressure"))
OUT = ["obs", "reward", "mask", "reward_global", "lane_count", "lane_queued", "sim"]


def _run(eng, bufs, n, controller=1, arg=25):
    import torch
    rows = []
    for _ in range(n):
        eng.env_step(None, bufs, n_ticks=5, controller=controller, controller_arg=arg)
        torch.cuda.synchronize()
        rows.append({k: v.cpu().numpy().copy() for k, v in bufs.items()})
    return rows


@pytest.mark.parametrize("device_blob", [False, True])
def test_save_load_replays_bit_for_bit(cuda_lib, device_blob):
    from pytsc_b200.binding import Engine
    cfg, parser, cs = build_scenario("hangzhou_4_4", **KW)
    eng = Engine(cs, 3, 0, vehicle_capacity=640)
    bufs = eng.alloc_outputs(OUT)
    eng.init_program(0)
    _run(eng, bufs, 30)
    blob = eng.save_state(device=device_blob)
    snap0 = eng.snapshot(1)
    first = _run(eng, bufs, 25)
    end1 = eng.snapshot(2)
    eng.load_state(blob)
    assert compare_snapshots(snap0, eng.snapshot(1)) is None
    again = _run(eng, bufs, 25)
    for a, b in zip(first, again):
        for k in OUT:
            assert np.array_equal(a[k], b[k]), k
    assert compare_snapshots(end1, eng.snapshot(2)) is None
    eng.check()
    # a blob from another batch size is refused
    other = Engine(cs, 2, 0, vehicle_capacity=640)
    with pytest.raises(RuntimeError):
        other.load_state(blob)
    other.close()
    eng.close()


def test_reset_replicas_restarts_only_those(cuda_lib):
    from pytsc_b200.binding import Engine
    cfg, parser, cs = build_scenario("hangzhou_4_4", **KW)
    a = Engine(cs, 3, 0, vehicle_capacity=640)
    fresh = Engine(cs, 1, 0, vehicle_capacity=640)
    ba, bf = a.alloc_outputs(OUT), fresh.alloc_outputs(OUT)
    a.init_program(0); fresh.init_program(0)
    ra = _run(a, ba, 24)
    a.reset_replicas([1])
    ra += _run(a, ba, 16)
    rf = _run(fresh, bf, 40)
    # replica 1 restarted at step 24: its next 16 steps are the fresh engine's first 16
    for k in range(16):
        for name in OUT:
            assert np.array_equal(ra[24 + k][name][1], rf[k][name][0]), (k, name)
    # replicas 0 and 2 ran on: all 40 steps equal the fresh engine's
    for k in range(40):
        for name in OUT:
            assert np.array_equal(ra[k][name][0], rf[k][name][0]), (k, name)
            assert np.array_equal(ra[k][name][2], rf[k][name][0]), (k, name)
    assert compare_snapshots(a.snapshot(0), fresh.snapshot(0)) is None
    assert a.counters()["tick"].tolist() == [200, 80, 200]
    a.check(); fresh.check()
    a.close(); fresh.close()
