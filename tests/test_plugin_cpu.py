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


def test_batched_density_map_matches_metrics_parser():
    """env.batched_density_map (torch, [B, L] -> [B, A, A]) against the formula MetricsParser.density_map
    applies lane by lane (backends/cityflow/metrics.py:170-199), random occupancies incl. values above 1."""
    import numpy as np
    import torch
    from helpers import build_scenario
    from pytsc_b200.env import batched_density_map, density_map_operator
    for name in ("hangzhou_4_4", "jinan_3_4", "syn_1x1"):
        cfg, parser, cs = build_scenario(name)
        W, adj = density_map_operator(parser)
        rng = np.random.RandomState(5)
        occ = rng.uniform(0, 1.6, size=(3, len(parser.lanes)))
        got = batched_density_map(torch.from_numpy(occ), torch.from_numpy(W), torch.from_numpy(adj)).numpy()
        ids = list(parser.traffic_signals.keys())
        for b in range(3):
            lanes = {l: {"occupancy": occ[b, k]} for k, l in enumerate(parser.lanes)}
            dm = np.zeros((len(ids), len(ids)))
            for i, ts in enumerate(ids):
                for j, other in enumerate(ids):
                    if parser.neighbors_lanes[ts] and other in parser.neighbors_lanes[ts]:
                        ls = parser.neighbors_lanes[ts][other]
                        dm[i, j] = np.clip(sum(lanes[l]["occupancy"] for l in ls) / len(ls), 0, 1)
            want = (dm + dm.T) / 2 + 1e-6 * adj
            assert np.allclose(got[b], want, rtol=1e-12, atol=1e-15), name
