"""GPU parity on generated grids and on the global-memory working-set variant.

BASELINE config 4 is a synthetic 16 x 16 grid (256 signals, 3264 lanes, 9216
lane-links) under heavy demand: one replica needs ~1 MB of working set, more
than an SM's shared memory, so the step kernel runs its GMEM variant (same
code over a global-memory workspace).  Checked here, bit for bit, against the
CPU oracle on the same generated roadnet / flow files:

* the GMEM variant forced onto scenarios that also run in shared memory
  (Hangzhou, a generated 6 x 6 grid): both variants == oracle;
* the 16 x 16 grid itself (GMEM by necessity): lock-step vehicle snapshots, and
  the fused env-step outputs against the Python port of pytsc's hot path.
"""
import os

import numpy as np
import pytest

from helpers import build_scenario, compare_snapshots, oracle_engine, signal_inter_indices

pytestmark = pytest.mark.gpu


def _lockstep(cs, cfg, parser, ticks, capacity, every=5, B=2):
    import torch
    from pytsc_b200.binding import Engine
    orc = oracle_engine(cfg)
    eng = Engine(cs, B, 0, vehicle_capacity=capacity)
    inter = signal_inter_indices(parser)
    A = eng.A
    raw = np.ones((B, A), np.int32)
    for t in range(ticks):
        if t % 5 == 0:
            k = (t // 30) % 8
            r = np.asarray([((k + a) % 8 + 1) if (t % 30) < 25 else 0 for a in range(A)], np.int32)
            raw[:] = r
            eng.set_phase(torch.from_numpy(raw).cuda())
            for a in range(A):
                orc.set_tl_phase_idx(inter[a], int(r[a]))
        orc.next_step()
        eng.step(1)
        if t % every == 0 or t == ticks - 1:
            assert orc.get_vehicle_count() <= capacity, "test capacity too small"
            msg = compare_snapshots(orc.snapshot(), eng.snapshot(B - 1))
            assert msg is None, f"tick {t}: {msg}"
    eng.check()
    info = eng.kernel_info()
    n = orc.get_vehicle_count()
    eng.close()
    return info, n


@pytest.fixture
def gmem_forced():
    os.environ["TSC_B200_GMEM"] = "1"
    yield
    os.environ.pop("TSC_B200_GMEM", None)


def test_gmem_variant_on_hangzhou(cuda_lib, gmem_forced):
    cfg, parser, cs = build_scenario("hangzhou_4_4")
    info, n = _lockstep(cs, cfg, parser, 400, 1280)
    assert info["threads"] == 1024 and info["global_workspace"] and n > 100


@pytest.fixture(scope="module")
def grid_6x6(tmp_path_factory):
    from pytsc_b200.generators import write_grid_scenario
    return write_grid_scenario(tmp_path_factory.mktemp("grids"), 6, 6, vehicles_per_hour_per_road=900, horizon=600, seed=2)


@pytest.mark.parametrize("capacity", [2200, 2600])
def test_generated_6x6_grid_shared_memory(cuda_lib, grid_6x6, capacity):
    """One 512-thread block per SM, everything in shared memory."""
    cfg, parser, cs = build_scenario(grid_6x6)
    assert cs.n_signals == 36
    info, n = _lockstep(cs, cfg, parser, 360, capacity)
    assert info["threads"] == 512 and n > 1500
    assert not info["global_workspace"]


def test_generated_6x6_grid_gmem(cuda_lib, grid_6x6, gmem_forced):
    cfg, parser, cs = build_scenario(grid_6x6)
    info, n = _lockstep(cs, cfg, parser, 360, 2200)
    assert info["threads"] == 1024


@pytest.fixture(scope="module")
def grid_16x16(tmp_path_factory):
    from pytsc_b200.generators import write_grid_scenario
    return write_grid_scenario(tmp_path_factory.mktemp("grids16"), 16, 16, vehicles_per_hour_per_road=900, horizon=300,
                               seed=0, signal=dict(observation_space="lane_features", reward_function="queue_length"))


def test_16x16_grid_lockstep(cuda_lib, grid_16x16):
    cfg, parser, cs = build_scenario(grid_16x16)
    assert (cs.n_signals, cs.n_lanes, cs.n_lanelinks) == (256, 3264, 9216)
    info, n = _lockstep(cs, cfg, parser, 240, 8000, every=20)
    assert info["threads"] == 1024       # does not fit shared memory
    assert n > 2000


def test_16x16_grid_env_step_against_port(cuda_lib, grid_16x16):
    """Fused env-step (in-kernel fixed-time control) on the 16 x 16 grid vs the port of pytsc's
    Python half over the oracle engine: integers equal, float features within 1e-5 relative."""
    import torch
    from oracle.pytsc_port import PortEnv
    from pytsc_b200.binding import Engine
    kw = dict(signal=dict(observation_space="lane_features", reward_function="queue_length",
                          action_space="phase_selection", round_robin=False))
    cfg, parser, cs = build_scenario(grid_16x16, **kw)
    port = PortEnv(grid_16x16, **kw)
    eng = Engine(cs, 2, 0, vehicle_capacity=8000)
    bufs = eng.alloc_outputs(["obs", "reward", "reward_global", "mask", "lane_count", "lane_queued", "sim"])
    eng.init_program(0)
    ids = cs.lane_ids
    for t in range(24):
        acts = port.fixed_time_actions(25)
        r, done, info = port.step(acts, phase_indices=True)
        eng.env_step(None, bufs, n_ticks=5, controller=1, controller_arg=25)
        torch.cuda.synchronize()
        lm = port.step_measurements["lane"]
        assert [lm[l]["n_vehicles"] for l in ids] == bufs["lane_count"][1].tolist(), t
        assert [lm[l]["n_queued"] for l in ids] == bufs["lane_queued"][1].tolist(), t
        assert np.array_equal(np.asarray(port.get_action_mask(), np.uint8), bufs["mask"][0].cpu().numpy()), t
        assert np.array_equal(np.asarray(port.get_observations(), np.float64), bufs["obs"][0].cpu().numpy().astype(np.float64)), t
        np.testing.assert_allclose(bufs["reward"][1].cpu().numpy(), np.asarray(port.get_rewards()), rtol=1e-5)
        assert float(bufs["reward_global"][0]) == pytest.approx(r, rel=1e-5)
        assert int(bufs["sim"][0, 0]) == port.step_measurements["sim"]["n_vehicles"]
    eng.check()
    eng.close()
