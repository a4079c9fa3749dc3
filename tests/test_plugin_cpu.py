"""CPU: host logic of the `gpu` backend plugin under the UNMODIFIED reference.

``pytsc.TrafficSignalNetwork(scenario, "gpu", ...)`` is built from the plugin
classes in ``pytsc_b200/backend`` through pytsc's own registry, with the device
engine replaced by a CPU test double (``helpers.FakeDeviceEngine``, oracle
backed).  pytsc's action spaces, observation spaces, reward functions and
metrics then run unchanged on top of the plugin, and everything they return is
compared with the golden fixtures recorded from the CityFlow backend.  Skipped
where the reference package is not importable.
"""
import numpy as np
import pytest

from helpers import FakeDeviceEngine, load_golden, reference_pytsc

pytsc = reference_pytsc()
pytestmark = pytest.mark.skipif(pytsc is None, reason="reference pytsc not importable here")


@pytest.fixture
def fake_engine(monkeypatch):
    from pytsc_b200.backend import simulator as sim_mod
    state = {}

    def factory(scenario, n_replicas, device=0, vehicle_capacity=0):
        return FakeDeviceEngine(scenario, n_replicas, device, vehicle_capacity,
                                port_kwargs=state["kwargs"], scenario_name=state["scenario"])
    monkeypatch.setattr(sim_mod, "Engine", factory)
    return state


def test_backend_is_registered():
    assert "gpu" in pytsc.SUPPORTED_SIMULATOR_BACKENDS
    mods = pytsc.SIMULATOR_MODULES["gpu"]
    assert set(mods) >= {"config", "metrics_parser", "network_parser", "retriever", "simulator", "traffic_signal"}


@pytest.mark.parametrize("case", ["syn_1x1__pm_queue_switch", "hangzhou_4_4__lf_pressure_select",
                                  "hangzhou_4_4__pm_queue_select_rr", "jinan_3_4__lf_queue_select"])
def test_reference_facade_on_gpu_plugin(fake_engine, case):
    g = load_golden(case)
    fake_engine.update(scenario=g["scenario"], kwargs=g["kwargs"])
    net = pytsc.TrafficSignalNetwork(g["scenario"], "gpu", **g["kwargs"])
    assert net.n_agents == len(g["signal_ids"])
    assert list(net.traffic_signals) == [str(x) for x in g["signal_ids"]]
    assert np.array_equal(np.asarray(net.get_action_mask(), np.uint8), g["mask0"])
    assert net.get_observation_size() == g["obs"].shape[-1]
    for t in range(min(int(g["n_steps"]), 36)):
        r, done, info = net.step([int(a) for a in g["actions"][t]])
        assert r == g["reward_global"][t]
        assert np.array_equal(np.asarray(net.get_rewards(), np.float64), g["reward"][t])
        assert np.array_equal(np.asarray(net.get_action_mask(), np.uint8), g["mask"][t])
        assert np.array_equal(np.asarray(net.get_observations(), np.float64), g["obs"][t])
        assert np.array_equal(np.asarray(net.get_state(), np.float64), g["state"][t])
        for k, i in (("n_queued", 0), ("mean_speed", 1), ("mean_delay", 2), ("density", 3), ("pressure", 4),
                     ("network_flow", 5)):
            assert info[k] == pytest.approx(g["metrics"][t][i], rel=1e-12, abs=1e-15)
        assert info["average_travel_time"] == g["sim"][t][1] and info["time_step"] == g["sim"][t][2]
        assert done == ((t + 1) % 72 == 0)
    net.simulator.close_simulator()


def test_rule_based_controller_and_wrapper(fake_engine):
    """FixedTimeController through TrafficSignal.get_controller_action, and the
    epymarl wrapper's reset / step / get_avail_actions on the gpu backend."""
    kw = dict(signal=dict(observation_space="lane_features", reward_function="queue_length",
                          action_space="phase_selection", round_robin=False))
    fake_engine.update(scenario="syn_1x1", kwargs=kw)
    net = pytsc.TrafficSignalNetwork("syn_1x1", "gpu", **kw)
    seq = []
    for _ in range(8):
        acts = [ts.get_controller_action("fixed_time") for ts in net.traffic_signals.values()]
        net.step(acts)
        seq.append(acts[0])
    assert seq == [0, 0, 0, 0, 0, 1, 2, 2]
    stats = net.get_env_stats()
    assert "n_vehicles" in stats and "lane" not in stats
    from pytsc.wrappers.epymarl import EPyMARLTrafficSignalNetwork
    env = EPyMARLTrafficSignalNetwork(map_name="syn_1x1", simulator_backend="gpu", **kw)
    info = env.get_env_info()
    assert info["n_agents"] == 1 and info["obs_shape"] == 212 and info["n_actions"] == 16
    obs, state = env.reset()
    avail = env.get_avail_actions()
    step_out = env.step([int(np.flatnonzero(a)[0]) for a in avail])
    assert len(step_out) == 5 and len(step_out[0]) == info["n_agents"]


def test_density_map_tables_follow_neighbors_lanes():
    """The compiled density-map tables (dm_off / dm_lane / dm_adjacency) hold parsed_network.neighbors_lanes in agent
    order and the adjacency matrix as the reference adds it (backends/cityflow/metrics.py:170-199)."""
    import numpy as np
    from helpers import build_scenario
    for name in ("hangzhou_4_4", "jinan_3_4", "manhattan_16_3", "syn_1x1"):
        cfg, parser, cs = build_scenario(name)
        ids = list(parser.traffic_signals.keys())
        A = len(ids)
        off, lanes = np.asarray(cs.dm_off), np.asarray(cs.dm_lane)
        assert off[-1] == cs.n_dm_total and len(off) == A * A + 1
        for i, ti in enumerate(ids):
            for j, tj in enumerate(ids):
                want = (parser.neighbors_lanes.get(ti) or {}).get(tj) or []
                got = [cs.lane_ids[l] for l in lanes[off[i * A + j]:off[i * A + j + 1]]]
                assert got == list(want), (name, ti, tj)
        assert np.array_equal(np.asarray(cs.dm_adjacency).reshape(A, A), np.asarray(parser.adjacency_matrix, np.float64))
