"""CPU: the oracle (C++ engine + Python port of pytsc's hot path) against the
golden fixtures recorded from the reference's own classes.  Everything must be
identical: same engine, same fp64 Python arithmetic in the same order."""
import numpy as np
import pytest

from helpers import golden_cases, load_golden

FAST = ["syn_1x1__pm_queue_switch", "syn_1x1__lf_pressure_select", "hangzhou_4_4__lf_pressure_select",
        "hangzhou_4_4__pm_queue_select_rr", "jinan_3_4__lf_queue_select", "manhattan_16_3__pm_pressure_switch"]


def test_fixture_inventory():
    assert set(FAST) <= set(golden_cases())


@pytest.mark.parametrize("case", FAST)
def test_port_reproduces_reference(case):
    from oracle.pytsc_port import PortEnv
    g = load_golden(case)
    env = PortEnv(g["scenario"], **g["kwargs"])
    ids = [str(x) for x in g["lane_ids"]]
    assert list(env.signals) == [str(x) for x in g["signal_ids"]]
    assert np.array_equal(np.asarray(env.get_action_mask(), np.uint8), g["mask0"])
    T = min(int(g["n_steps"]), 48)
    for t in range(T):
        r, done, info = env.step([int(a) for a in g["actions"][t]])
        lm = env.step_measurements["lane"]
        assert [lm[l]["n_vehicles"] for l in ids] == list(g["lane_count"][t])
        assert [lm[l]["n_queued"] for l in ids] == list(g["lane_queued"][t])
        assert np.array_equal([float(lm[l]["occupancy"]) for l in ids], g["lane_occupancy"][t])
        assert np.array_equal([float(lm[l]["mean_speed"]) for l in ids], g["lane_mean_speed"][t])
        assert r == g["reward_global"][t]
        assert np.array_equal(np.asarray(env.get_rewards(), np.float64), g["reward"][t])
        assert np.array_equal(np.asarray(env.get_action_mask(), np.uint8), g["mask"][t])
        assert np.array_equal(np.asarray(env.get_observations(), np.float64), g["obs"][t])
        assert np.array_equal(np.asarray(env.get_state(), np.float64), g["state"][t])
        sig = list(env.signals.values())
        got = np.asarray([[s.n_queued, s.occupancy, s.mean_speed, s.mean_delay, s.outgoing_occupancy, s.pressure,
                           s.stat_time_on_phase, s.current_phase_index] for s in sig], np.float64)
        assert np.array_equal(got, g["sig_stats"][t])
        pin = np.asarray([s.inc_position_matrices[l] for s in sig for l in s.incoming_lanes])
        pout = np.asarray([s.out_position_matrices[l] for s in sig for l in s.outgoing_lanes])
        assert np.array_equal(pin, g["pos_in"][t]) and np.array_equal(pout, g["pos_out"][t])
        st = env.step_stats()
        got = [st["n_queued"], st["mean_speed"], st["mean_delay"], st["density"], st["pressure"], st["network_flow"],
               float(env.flickering_signal), float(env.norm_mean_speed)]
        assert np.array_equal(np.asarray(got, np.float64), g["metrics"][t])
        sm = env.step_measurements["sim"]
        assert [sm["n_vehicles"], sm["average_travel_time"], sm["time_step"]] == list(g["sim"][t][:3])
    for t in [int(x) for x in g["snap_steps"]]:
        if t < T:
            pass   # snapshots are checked in test_oracle.py over the full horizon


def test_fixed_time_controller_cycle():
    """FixedTimeController (controllers.py:39-54): 25 s green then 5 s yellow, round the 16-phase plan."""
    from oracle.pytsc_port import PortEnv
    env = PortEnv("syn_1x1", signal=dict(action_space="phase_selection", round_robin=False))
    seq = []
    for _ in range(14):
        a = env.fixed_time_actions(25)
        env.step(a)
        seq.append(a[0])
    assert seq == [0, 0, 0, 0, 0, 1, 2, 2, 2, 2, 2, 3, 4, 4]


def test_bench_host_policy_is_the_fixed_time_rule():
    """bench.py's table-driven host policy == FixedTimeController.get_action + update_current_phase
    written out (controllers/controllers.py:39-54, common/traffic_signal.py:94-109), 800 steps."""
    import numpy as np
    import bench
    from helpers import build_scenario
    w = bench.CONFIGS["hangzhou"]
    cfg, parser, cs = build_scenario(w["scenario"], **w["kw"])
    A, B, G, dt = cs.n_signals, 7, bench.GREEN_TIME, 5
    green = np.ascontiguousarray(cs.sig_phase_green).reshape(A, -1).astype(bool)
    nph = np.asarray(cs.sig_n_phases, np.int64)
    pol = bench.HostFixedTimePolicy(cs.sig_phase_green, cs.sig_n_phases, B, A, G, dt)
    cur = np.zeros((B, A), np.int64)
    top = np.zeros((B, A), np.int64)
    out = np.zeros((B, A), np.int32)
    for step in range(800):
        want = np.empty((B, A), np.int64)
        for b in range(B):
            for a in range(A):
                stay = green[a, cur[b, a]] and top[b, a] < G
                want[b, a] = cur[b, a] if stay else (cur[b, a] + 1) % nph[a]
        top = np.where(want == cur, top + dt, dt)
        cur = want
        pol.act(out)
        assert np.array_equal(out, want), step
    pol.reset()
    pol.act(out)
    assert (out == 0).all() or (out[0] == out[-1]).all()
