"""GPU: the registered end-to-end path (``tsc_host_register`` / ``tsc_env_step_registered``).

One launch per step; every replica block stores a compact packet into page-locked host memory and raises a
flag; host threads finish the caller's rows, writing only what changed.  The caller's arrays must hold, after
every step, exactly what ``tsc_env_step`` writes on the device: the observation contract
(``pytsc/common/observations.py:305-329``) stays bit-equal, and the golden fixtures recorded from the
reference's own classes are checked through this path too.
"""
import numpy as np
import pytest

from helpers import golden_scenario, load_golden

pytestmark = pytest.mark.gpu


def _host_arrays(eng):
    d = eng.dims
    return dict(obs=np.full((d["B"], d["A"], d["obs_dim"]), 7.0, np.float32), reward=np.zeros((d["B"], d["A"]), np.float32),
                mask=np.full((d["B"], d["A"], d["n_actions"]), 9, np.uint8), reward_global=np.zeros((d["B"],), np.float32))


@pytest.mark.parametrize("case,exact,threads", [
    ("hangzhou_4_4__lf_pressure_select", True, "3"),      # one u32 per lane (small integers)
    ("hangzhou_4_4__lf_pressure_select", False, "1"),     # float features: three floats per lane
    ("hangzhou_4_4_5816__lf_queue_switch", True, "8"),    # phase_switch action space (2 actions), round robin
    ("jinan_3_4__lf_queue_select", True, "2"),
    ("manhattan_16_3__lf_queue_select", True, "4"),
])
def test_registered_path_equals_device_path(cuda_lib, case, exact, threads, monkeypatch):
    import torch
    from pytsc_b200.binding import Engine
    monkeypatch.setenv("TSC_B200_HOST_THREADS", threads)
    g = load_golden(case)
    cfg, parser, cs = golden_scenario(g, reference_exact=exact)
    B = 37                                    # not a multiple of the worker group size
    dev, host = Engine(cs, B, 0, vehicle_capacity=1280), Engine(cs, B, 0, vehicle_capacity=1280)
    bufs = dev.alloc_outputs(["obs", "reward", "mask", "reward_global"])
    h = _host_arrays(host)
    host.host_register(**h)
    assert host.host_packet_bytes() < h["obs"].nbytes / (8 if exact else 4)      # u32 per lane / three floats per lane
    dev.init_program(0); host.init_program(0)
    T = int(g["n_steps"])
    rng = np.random.RandomState(3)
    for t in range(min(T, 60)):
        act = np.repeat(g["actions"][t][None], B, 0).astype(np.int32)
        if t >= 20:                           # replicas diverge: every row has its own history of changes
            act[1::2] = np.roll(act[1::2], t % dev.A, axis=1)
        dev.env_step(torch.from_numpy(act).cuda(), bufs, n_ticks=5)
        host.env_step_registered(act, n_ticks=5)
        torch.cuda.synchronize()
        assert np.array_equal(bufs["obs"].cpu().numpy(), h["obs"]), (case, t)
        assert np.array_equal(bufs["reward"].cpu().numpy(), h["reward"]), (case, t)
        assert np.array_equal(bufs["mask"].cpu().numpy(), h["mask"]), (case, t)
        assert np.array_equal(bufs["reward_global"].cpu().numpy(), h["reward_global"]), (case, t)
        if exact and t < 20:
            assert np.array_equal(h["obs"][B - 1].astype(np.float64), g["obs"][t]), (case, t)      # the reference's own rows
        if t == 30:                           # some replicas restart: their rows must follow
            idx = rng.choice(B, 5, replace=False)
            dev.reset_replicas(idx); host.reset_replicas(idx)
    dev.check(); host.check()
    host.host_unregister()
    dev.close(); host.close()


def test_registered_path_in_kernel_controller_full_batch(cuda_lib):
    """The bench configuration: B = 4096, in-kernel fixed-time control, default worker threads, 100 steps."""
    import torch
    from pytsc_b200.binding import Engine
    from helpers import build_scenario
    cfg, parser, cs = build_scenario("hangzhou_4_4", signal=dict(observation_space="lane_features", reward_function="max_pressure",
                                                                 action_space="phase_selection", round_robin=False))
    B = 4096
    dev, host = Engine(cs, B, 0, vehicle_capacity=640), Engine(cs, B, 0, vehicle_capacity=640)
    bufs = dev.alloc_outputs(["obs", "reward", "mask", "reward_global"])
    h = _host_arrays(host)
    host.host_register(**h)
    dev.init_program(0); host.init_program(0)
    for t in range(100):
        dev.env_step(None, bufs, n_ticks=5, controller=1, controller_arg=25)
        host.env_step_registered(None, n_ticks=5, controller=1, controller_arg=25)
        if t % 10 == 9:
            torch.cuda.synchronize()
            assert np.array_equal(bufs["obs"].cpu().numpy(), h["obs"]), t
            assert np.array_equal(bufs["reward"].cpu().numpy(), h["reward"]), t
            assert np.array_equal(bufs["mask"].cpu().numpy(), h["mask"]), t
            assert np.array_equal(bufs["reward_global"].cpu().numpy(), h["reward_global"]), t
    dev.check(); host.check()
    dev.close(); host.close()


def test_env_step_host_api(cuda_lib):
    """``BatchedTrafficSignalNetwork.step_host``: numpy in, numpy out, the same numbers as ``step``."""
    import torch
    from pytsc_b200 import BatchedTrafficSignalNetwork
    kw = dict(signal=dict(observation_space="lane_features", reward_function="queue_length",
                          action_space="phase_selection", round_robin=False), gpu=dict(vehicle_capacity=1200))
    a = BatchedTrafficSignalNetwork("jinan_3_4", n_replicas=9, **kw)
    b = BatchedTrafficSignalNetwork("jinan_3_4", n_replicas=9, **kw)
    rng = np.random.RandomState(0)
    for t in range(40):
        m = a.get_action_mask().cpu().numpy()
        act = np.array([[rng.choice(np.flatnonzero(m[r, s])) for s in range(a.n_agents)] for r in range(9)], np.int32)
        a.step(torch.from_numpy(act).cuda())
        rg, over, out = b.step_host(act)
        assert np.array_equal(a.get_observations().cpu().numpy(), out["obs"]), t
        assert np.array_equal(a.get_rewards().cpu().numpy(), out["reward"]), t
        assert np.array_equal(a.get_action_mask().cpu().numpy(), out["mask"]), t
        assert np.array_equal(a.get_reward().cpu().numpy(), rg), t
    a.check(); b.check()
    a.close(); b.close()


def test_double_buffered_halves_equal_one_batch(cuda_lib):
    """``step_host_begin`` / ``step_host_wait``: two environments of B / 2 replicas stepped alternately (both launches in
    flight while the host prepares the next actions) hold, half by half, the rows of one synchronous batch of B;
    pairing errors are reported."""
    import torch
    from pytsc_b200 import BatchedTrafficSignalNetwork
    from pytsc_b200.binding import TscError
    kw = dict(signal=dict(observation_space="lane_features", reward_function="max_pressure",
                          action_space="phase_selection", round_robin=False), gpu=dict(vehicle_capacity=640))
    B = 300
    whole = BatchedTrafficSignalNetwork("hangzhou_4_4", n_replicas=B, **kw)
    halves = [BatchedTrafficSignalNetwork("hangzhou_4_4", n_replicas=B // 2, **kw) for _ in range(2)]
    outs = [h.register_host_buffers(threads=2) for h in halves]
    with pytest.raises(TscError):
        halves[0].step_host_wait()                       # nothing in flight
    rng = np.random.RandomState(5)
    acts = [np.zeros((B // 2, whole.n_agents), np.int32) for _ in range(2)]

    def draw(k):                                          # every replica its own valid actions, from its own previous mask
        m = outs[k]["mask"] if draw.started else np.ones_like(outs[k]["mask"])
        first_allowed = m.argmax(-1)
        last_allowed = m.shape[-1] - 1 - m[..., ::-1].argmax(-1)
        acts[k][...] = np.where(rng.rand(*first_allowed.shape) < 0.5, first_allowed, last_allowed)
    draw.started = False
    draw(0); halves[0].step_host_begin(acts[0])
    for t in range(60):
        draw(1); halves[1].step_host_begin(acts[1])
        if t == 0:
            with pytest.raises(TscError):
                halves[1].step_host_begin(acts[1])        # one step per handle in flight
        sent = [acts[0].copy(), acts[1].copy()]
        halves[0].step_host_wait()
        halves[1].step_host_wait()
        draw.started = True
        whole.step(torch.from_numpy(np.concatenate(sent)).cuda())
        ref = {"obs": whole.get_observations().cpu().numpy(), "reward": whole.get_rewards().cpu().numpy(),
               "mask": whole.get_action_mask().cpu().numpy(), "reward_global": whole.get_reward().cpu().numpy()}
        for k in range(2):
            for name, arr in ref.items():
                assert np.array_equal(arr[k * (B // 2):(k + 1) * (B // 2)], outs[k][name]), (t, k, name)
        draw(0); halves[0].step_host_begin(acts[0])
    halves[0].step_host_wait()
    for e in [whole] + halves:
        e.check(); e.close()


def test_bad_phase_is_reported(cuda_lib):
    """tsc_set_phase with a light phase the signal does not have sets a sticky flag (TSC_EINVAL), readable
    without a sync through the ``err`` output; tsc_init_program validates on the host."""
    import torch
    from helpers import build_scenario
    from pytsc_b200.binding import Engine, TscError
    cfg, parser, cs = build_scenario("syn_1x1")
    eng = Engine(cs, 3, 0, vehicle_capacity=256)
    with pytest.raises(TscError):
        eng.init_program(99)
    raw = torch.ones((3, eng.A), dtype=torch.int32, device="cuda")
    raw[1, 0] = 40
    eng.set_phase(raw)
    bufs = eng.alloc_outputs(["err", "sim"])
    eng.step(5)
    eng.retrieve(bufs)
    err = bufs["err"].cpu().numpy()
    assert err[0] == 0 and err[2] == 0 and err[1] & 8
    with pytest.raises(TscError) as e:
        eng.check()
    assert e.value.code == -1
    eng.close()
