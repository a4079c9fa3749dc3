"""CPU: pin the C++ oracle engine -- against the vehicle snapshots stored in the
golden fixtures (regression), against conservation laws, and against the
independently computed geometry of the scenario compiler."""
import numpy as np
import pytest

from helpers import build_scenario, golden_scenario, load_golden, oracle_engine, signal_inter_indices


@pytest.mark.parametrize("case", ["syn_1x1__lf_pressure_select", "hangzhou_4_4__lf_pressure_select",
                                  "jinan_3_4__lf_queue_select", "manhattan_16_3__lf_queue_select"])
def test_engine_reproduces_golden_snapshots(case):
    g = load_golden(case)
    cfg, parser, cs = golden_scenario(g)
    orc = oracle_engine(cfg)
    inter = signal_inter_indices(parser)
    raw = cs.sig_phase_raw.reshape(len(inter), -1)
    cur = np.zeros(len(inter), int)
    switch = g["kwargs"]["signal"]["action_space"] == "phase_switch"
    for a, i in enumerate(inter):
        orc.set_tl_phase_idx(i, int(raw[a, 0]))
    snaps = [int(x) for x in g["snap_steps"]]
    for t in range(int(g["n_steps"])):
        for a, i in enumerate(inter):
            act = int(g["actions"][t][a])
            idx = ((cur[a] + 1) % cs.sig_n_phases[a] if act == 1 else cur[a]) if switch else act
            cur[a] = idx
            orc.set_tl_phase_idx(i, int(raw[a, idx]))
        orc.next_steps(5)
        assert orc.get_vehicle_count() == int(g["sim"][t][0])
        if t in snaps:
            s = orc.snapshot()
            for k in ("uid", "drivable", "distance", "speed"):
                assert np.array_equal(s[k], g[f"snap{t}_{k}"]), (case, t, k)


@pytest.mark.parametrize("name", ["syn_1x1", "hangzhou_4_4", "syn_1x3_gaussian", "syn_5x5_oneway", "new_york_arterial"])
def test_vehicle_conservation_and_travel_time(name):
    cfg, parser, cs = build_scenario(name)
    orc = oracle_engine(cfg)
    inter = signal_inter_indices(parser)
    tick = np.asarray(cs.veh_tick)
    for t in range(900):
        if t % 30 == 0:
            for a, i in enumerate(inter):
                n_raw = int(cs.sig_n_raw_phases[a])        # one-way grids have fewer light phases than the 9 of a four-way signal
                orc.set_tl_phase_idx(i, 1 + (t // 30) % (n_raw - 1) if n_raw > 1 else 0)
        orc.next_step()
        if t % 50 == 49:
            created = int((tick <= t).sum())
            waiting = int(orc.waiting_buffer_sizes().sum())
            assert orc.get_created_vehicle_count() == created
            assert orc.get_vehicle_count() + orc.get_finished_vehicle_count() + waiting == created
            s = orc.snapshot()
            assert (s["speed"] >= 0).all() and (s["distance"] >= 0).all()
            # no overlap inside a drivable: front-to-back order, bumper gap non-negative
            same = s["drivable"][1:] == s["drivable"][:-1]
            gap = s["distance"][:-1] - 5.0 - s["distance"][1:]
            assert (gap[same] > -1e-9).all()
    assert orc.get_finished_vehicle_count() > 0
    assert orc.get_average_travel_time() > 0


def test_oracle_is_deterministic_and_resets():
    cfg, parser, cs = build_scenario("syn_1x1")
    a, b = oracle_engine(cfg), oracle_engine(cfg)
    a.next_steps(300)
    b.next_steps(120)
    b.reset()
    b.next_steps(300)
    sa, sb = a.snapshot(), b.snapshot()
    for k in ("uid", "drivable", "distance", "speed"):
        assert np.array_equal(sa[k], sb[k])


@pytest.mark.parametrize("name", ["syn_1x1", "hangzhou_4_4", "manhattan_16_3"])
def test_compiler_geometry_matches_oracle(name):
    """Lane / lane-link lengths and crosses are derived twice, in Python
    (pytsc_b200/roadnet.py) and in C++ (the oracle): bit-identical."""
    cfg, parser, cs = build_scenario(name)
    orc = oracle_engine(cfg)
    assert np.array_equal(orc.drivable_lengths(), cs.drv_length)
    assert orc.lane_ids == cs.lane_ids
    L = cs.n_lanes
    inter = signal_inter_indices(parser)
    n = 0
    for a, i in enumerate(inter):
        ll0, ll1, d0, d1 = orc.crosses(i)
        n += len(ll0)
        for k in range(len(ll0)):
            a0, a1 = ll0[k] - L, ll1[k] - L          # the oracle reports drivable indices
            lo, hi = cs.ll_cross_off[a0], cs.ll_cross_off[a0 + 1]
            hit = [x for x in range(lo, hi) if cs.xr_foe_ll[x] == a1]
            assert len(hit) == 1
            assert cs.xr_dist[hit[0]] == d0[k] and cs.xr_foe_dist[hit[0]] == d1[k]
    assert 2 * n == cs.n_cross_entries
    assert L == orc.n_lanes and cs.n_lanelinks == orc.n_lanelinks
