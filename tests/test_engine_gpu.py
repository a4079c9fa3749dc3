"""GPU parity, tier 2: the CUDA engine in lock-step with the CPU oracle.

Same roadnet, same flows, same light-phase sequence; after every tick the
running vehicles (order, drivable, fp64 distance and speed) must agree.  The
north star asks for <= 1e-3 m over the first 300 ticks; both sides evaluate the
same fp64 expressions without fused multiply-adds, so the test asks for
bit-equality and runs 600 ticks.
"""
import numpy as np
import pytest

from helpers import build_scenario, compare_snapshots, oracle_engine, signal_inter_indices

pytestmark = pytest.mark.gpu

POSITION_TOL_M = 0.0   # bit-exact (north star: 1e-3 m)


def _phases(mode, t, A, nraw, rng):
    if mode == "random":
        return np.array([rng.randint(0, nraw[a]) for a in range(A)], np.int32)
    k = (t // 30) % 8                      # cyclic plan: 25 s green, 5 s yellow
    return np.full(A, (k + 1) if (t % 30) < 25 else 0, np.int32)


@pytest.mark.parametrize("mode", ["cyclic", "random"])
@pytest.mark.parametrize("name,ticks,kw", [
    ("syn_1x1", 900, {}),
    ("hangzhou_4_4", 600, {}),
    ("hangzhou_4_4", 300, {"cityflow": {"flow_file": "anon_4_4_hangzhou_real_5816.json"}}),
    ("jinan_3_4", 400, {}),
    ("manhattan_16_3", 300, {}),
    ("syn_3x3", 300, {}),
])
def test_lockstep_with_oracle(cuda_lib, name, ticks, kw, mode):
    import torch
    from pytsc_b200.binding import Engine
    cfg, parser, cs = build_scenario(name, **kw)
    orc = oracle_engine(cfg)
    B = 3
    eng = Engine(cs, B, 0, vehicle_capacity=1280)
    inter = signal_inter_indices(parser)
    A = eng.A
    rng = np.random.RandomState(1)
    raw = np.ones((B, A), np.int32)
    seen = 0
    for t in range(ticks):
        if t % 5 == 0:
            r = _phases(mode, t, A, cs.sig_n_raw_phases, rng)
            raw[:] = r
            eng.set_phase(torch.from_numpy(raw).cuda())
            for a in range(A):
                orc.set_tl_phase_idx(inter[a], int(r[a]))
        orc.next_step()
        eng.step(1)
        if t % 3 == 0 or t == ticks - 1:
            so = orc.snapshot()
            seen = max(seen, len(so["uid"]))
            for b in (0, B - 1):
                msg = compare_snapshots(so, eng.snapshot(b), POSITION_TOL_M)
                assert msg is None, f"{name}/{mode} tick {t} replica {b}: {msg}"
    eng.check()
    c = eng.counters()
    assert c["n_running"][0] == orc.get_vehicle_count()
    assert c["n_finished"][0] == orc.get_finished_vehicle_count()
    assert seen > 0
    eng.close()


def test_multi_tick_launch_equals_single_ticks(cuda_lib):
    """tsc_step(h, 5) == 5 x tsc_step(h, 1): the fused shared-memory loop keeps no hidden state."""
    import torch
    from pytsc_b200.binding import Engine
    cfg, parser, cs = build_scenario("hangzhou_4_4")
    e1, e5 = Engine(cs, 2, 0, 1280), Engine(cs, 2, 0, 1280)
    raw = torch.full((2, e1.A), 1, dtype=torch.int32, device="cuda")
    e1.set_phase(raw); e5.set_phase(raw)
    for k in range(60):
        for _ in range(5):
            e1.step(1)
        e5.step(5)
        if k % 10 == 9:
            assert compare_snapshots(e1.snapshot(0), e5.snapshot(1)) is None
    e1.check(); e5.check()
    e1.close(); e5.close()


@pytest.mark.parametrize("env,capacity,variant", [
    ({}, 600, (256, 0, 4)),                                 # default at this size: four 256-thread blocks per SM (64 registers)
    ({"TSC_B200_THREADS": "192"}, 600, (192, 0, 4)),        # four 192-thread blocks per SM (80 registers)
    ({"TSC_B200_THREADS": "192", "TSC_B200_MIN_BLOCKS": "5"}, 560, (192, 0, 5)),
    ({"TSC_B200_THREADS": "160"}, 560, (160, 0, 5)),
    ({"TSC_B200_THREADS": "256", "TSC_B200_MIN_BLOCKS": "3"}, 600, (256, 0, 3)),        # three 256-thread blocks per SM (80 registers)
    ({}, 1100, (256, 0, 3)),                                # larger replica: three 256-thread blocks per SM (80 registers)
    ({}, 1560, (384, 0, 2)),                                # two 384-thread blocks per SM (80 registers)
    ({"TSC_B200_THREADS": "256"}, 1560, (256, 0, 2)),       # two 256-thread blocks per SM (128 registers)
    ({"TSC_B200_ONE_TEMPLATE": "0"}, 600, (256, 0, 2)),     # per-vehicle template look-up although the scenario has one template
    ({"TSC_B200_PREFETCH": "0"}, 600, (256, 0, 4)),
    ({"TSC_B200_ASYNC_STAGE": "0"}, 600, (256, 0, 4)),      # plain vector copies instead of bulk asynchronous copies for staging
    ({"TSC_B200_ASYNC_STAGE": "1"}, 600, (256, 0, 4)),      # cp.async (16 bytes per request per thread)
    ({}, 2000, (512, 0, 1)),                                # one 512-thread block per SM
    ({"TSC_B200_THREADS": "1024"}, 2000, (1024, 0, 1)),     # the same with 32 warps at 64 registers
    ({"TSC_B200_GMEM": "1"}, 600, (1024, 2, 1)),            # vehicle columns in a global-memory workspace, the rest in shared memory
    ({"TSC_B200_GMEM": "1", "TSC_B200_GMEM_META_SHARED": "0"}, 600, (1024, 1, 1)),      # the whole working set in the workspace
])
def test_kernel_variants_agree_with_oracle(cuda_lib, env, capacity, variant, monkeypatch):
    """The code paths an environment switch (or an unusual scenario) selects at tsc_create produce the
    same trajectories: hangzhou_4_4 for 300 ticks in lock-step with the oracle, random light phases."""
    import torch
    from pytsc_b200.binding import Engine
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    cfg, parser, cs = build_scenario("hangzhou_4_4", signal=dict(observation_space="lane_features"))
    orc = oracle_engine(cfg)
    eng = Engine(cs, 2, 0, vehicle_capacity=capacity)
    info = eng.kernel_info()
    assert (info["threads"], info["global_workspace"], info["blocks_per_sm"]) == variant, info
    inter = signal_inter_indices(parser)
    rng = np.random.RandomState(7)
    raw = np.ones((2, eng.A), np.int32)
    for t in range(300):
        if t % 5 == 0:
            r = _phases("random", t, eng.A, cs.sig_n_raw_phases, rng)
            raw[:] = r
            eng.set_phase(torch.from_numpy(raw).cuda())
            for a in range(eng.A):
                orc.set_tl_phase_idx(inter[a], int(r[a]))
        orc.next_step()
        eng.step(1)
        if t % 10 == 9:
            msg = compare_snapshots(orc.snapshot(), eng.snapshot(1), POSITION_TOL_M)
            assert msg is None, f"{env} tick {t}: {msg}"
    eng.check()
    eng.close()
