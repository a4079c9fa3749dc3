"""GPU parity, tier 1: the fused CUDA env-step against the golden fixtures.

The fixtures hold what the *reference's own Python* (Retriever, TrafficSignal,
MetricsParser, action spaces, observation spaces, reward functions) reported
after every ``TrafficSignalNetwork.step(actions)``.  The same action sequence
is replayed through ``tsc_env_step`` (one launch per env-step: phase program,
five engine ticks, Retriever reductions, per-signal stats, rewards, masks,
observations).  Integers must be equal; floats within REL_TOL.
"""
import numpy as np
import pytest

from helpers import compare_snapshots, golden_cases, golden_scenario, load_golden

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5        # north star: float features within 1e-5 relative
REL_TOL_F64 = 1e-12   # fp64 outputs differ from the reference only by summation order


def _close(a, b, rtol, what, atol=0.0):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert not (err > 0).any(), f"{what}: max abs diff {np.abs(a - b).max()} at {np.unravel_index(np.argmax(err), err.shape)}"


@pytest.mark.parametrize("case", golden_cases())
def test_env_step_replays_reference(cuda_lib, case):
    import torch
    from pytsc_b200.binding import Engine
    g = load_golden(case)
    cfg, parser, cs = golden_scenario(g, reference_exact=True)
    B = 2
    eng = Engine(cs, B, 0, vehicle_capacity=1280)
    bufs = eng.alloc_outputs()
    eng.init_program(0)
    eng.retrieve(bufs)
    torch.cuda.synchronize()
    assert np.array_equal(bufs["mask"][B - 1].cpu().numpy(), g["mask0"])
    T = int(g["n_steps"])
    snap_steps = [int(x) for x in g["snap_steps"]]
    for t in range(T):
        act = torch.from_numpy(np.repeat(g["actions"][t][None], B, 0).astype(np.int32)).cuda()
        eng.env_step(act, bufs, n_ticks=cfg.simulator["delta_time"])
        torch.cuda.synchronize()
        o = {k: v[B - 1].cpu().numpy() for k, v in bufs.items()}
        tag = f"{case} step {t}"
        # integers: bit-exact
        assert np.array_equal(o["lane_count"], g["lane_count"][t]), tag
        assert np.array_equal(o["lane_queued"], g["lane_queued"][t]), tag
        assert np.array_equal(o["mask"], g["mask"][t]), tag
        assert np.array_equal(o["sig_stats64"][:, 0], g["sig_stats"][t][:, 0]), tag + " n_queued"
        assert np.array_equal(o["sig_stats64"][:, 7], g["sig_stats"][t][:, 7]), tag + " phase index"
        assert o["sim"][0] == g["sim"][t][0] and o["sim"][2] == g["sim"][t][2] and o["sim"][3] == g["sim"][t][3], tag
        assert o["metrics"][0] == g["metrics"][t][0], tag
        # observations / states: the reference's vectors are integer-truncated (pad_list), reproduced exactly
        assert np.array_equal(o["obs"].astype(np.float64), g["obs"][t]), tag + " obs"
        assert np.array_equal(o["state"].astype(np.float64), g["state"][t]), tag + " state"
        # floats
        _close(o["lane_occupancy"], g["lane_occupancy"][t], REL_TOL, tag + " occupancy")
        _close(o["lane_mean_speed"], g["lane_mean_speed"][t], REL_TOL, tag + " mean_speed")
        _close(o["lane_meas64"][:, 0], g["lane_occupancy"][t], REL_TOL_F64, tag + " occupancy64")
        _close(o["lane_meas64"][:, 1], g["lane_mean_speed"][t], REL_TOL_F64, tag + " mean_speed64")
        _close(o["sig_stats64"][:, 1:7], g["sig_stats"][t][:, 1:7], REL_TOL_F64, tag + " sig stats", atol=1e-15)
        _close(o["pos_in"], g["pos_in"][t], REL_TOL, tag + " pos_in")
        _close(o["pos_out"], g["pos_out"][t], REL_TOL, tag + " pos_out")
        _close(o["reward"], g["reward"][t], REL_TOL, tag + " reward")
        _close(o["reward_global"], g["reward_global"][t], REL_TOL, tag + " global reward")
        _close(o["sim"][1], g["sim"][t][1], REL_TOL_F64, tag + " average travel time")
        _close(o["metrics"][1:], g["metrics"][t][1:], 1e-9, tag + " metrics", atol=1e-15)
        if t in snap_steps:
            ref = {k: g[f"snap{t}_{k}"] for k in ("uid", "drivable", "distance", "speed")}
            for b in range(B):
                assert compare_snapshots(ref, eng.snapshot(b)) is None, tag
    eng.check()
    eng.close()


@pytest.mark.parametrize("case", ["hangzhou_4_4__lf_pressure_select", "jinan_3_4__lf_queue_select"])
def test_untruncated_observations(cuda_lib, case):
    """reference_exact=False keeps the float features the reference computes before
    pad_list truncates them: rebuilt here from the golden lane measurements."""
    import torch
    from pytsc_b200.binding import Engine
    g = load_golden(case)
    cfg, parser, cs = golden_scenario(g, reference_exact=False)
    eng = Engine(cs, 1, 0, vehicle_capacity=1280)
    bufs = eng.alloc_outputs(["obs"])
    eng.init_program(0)
    lane_feat = cs.lane_feat.reshape(-1, 9)
    for t in range(int(g["n_steps"])):
        eng.env_step(torch.from_numpy(g["actions"][t][None].astype(np.int32)).cuda(), bufs, n_ticks=5)
        if t % 6:
            continue
        obs = bufs["obs"][0].cpu().numpy()
        for a in range(eng.A):
            exp = []
            for e in range(cs.sig_in_off[a], cs.sig_in_off[a + 1]):
                l = cs.sig_in_lane[e]
                exp += list(lane_feat[l]) + [g["lane_queued"][t][l], g["lane_occupancy"][t][l], g["lane_mean_speed"][t][l]]
            exp += [-1.0] * (16 * 12 - len(exp))
            ph = [0.0] * 20
            ph[int(g["sig_stats"][t][a, 7])] = 1.0
            _close(obs[a], np.asarray(exp + ph), REL_TOL, f"{case} step {t} agent {a}")
    eng.close()


@pytest.mark.parametrize("pinned", [False, True])
def test_host_path_equals_device_path(cuda_lib, pinned):
    """tsc_env_step_host (chunked: step chunk k+1 while chunk k is copied out) returns exactly what
    tsc_env_step writes to device buffers -- pageable and page-locked host buffers, B not a multiple
    of the chunk count."""
    import torch
    from pytsc_b200.binding import Engine
    g = load_golden("hangzhou_4_4__lf_pressure_select")
    cfg, parser, cs = golden_scenario(g, reference_exact=True)
    B = 37
    dev, host = Engine(cs, B, 0, vehicle_capacity=640), Engine(cs, B, 0, vehicle_capacity=640)
    bufs = dev.alloc_outputs(["obs", "reward", "mask", "reward_global"])
    dev.init_program(0); host.init_program(0)
    mk = (lambda *s, dt: torch.empty(*s, dtype=dt, pin_memory=True).numpy()) if pinned else (lambda *s, dt: torch.empty(*s, dtype=dt).numpy())
    h_obs = mk(B, dev.A, dev.dims["obs_dim"], dt=torch.float32)
    h_rew = mk(B, dev.A, dt=torch.float32)
    h_mask = mk(B, dev.A, dev.dims["n_actions"], dt=torch.uint8)
    h_rg = mk(B, dt=torch.float32)
    h_act = mk(B, dev.A, dt=torch.int32)
    for t in range(40):
        h_act[:] = g["actions"][t][None]
        dev.env_step(torch.from_numpy(h_act.copy()).cuda(), bufs, n_ticks=5)
        host.env_step_host(h_act, obs=h_obs, reward=h_rew, mask=h_mask, reward_global=h_rg, n_ticks=5)
        torch.cuda.synchronize()
        assert np.array_equal(bufs["obs"].cpu().numpy(), h_obs), t
        assert np.array_equal(bufs["reward"].cpu().numpy(), h_rew), t
        assert np.array_equal(bufs["mask"].cpu().numpy(), h_mask), t
        assert np.array_equal(bufs["reward_global"].cpu().numpy(), h_rg), t
        assert np.array_equal(h_obs[B - 1].astype(np.float64), g["obs"][t])
    dev.check(); host.check()
    dev.close(); host.close()
