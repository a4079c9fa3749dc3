"""GPU: the public Python API on the real device engine.

* ``BatchedTrafficSignalNetwork`` (device tensors in / out) replayed against the
  golden fixtures;
* the UNMODIFIED reference facade ``pytsc.TrafficSignalNetwork(scenario, "gpu")``
  on top of the plugin classes, compared with the same fixtures (runs where the
  reference package is importable -- ``baseline/_ref`` travels to the GPU box --
  and is skipped otherwise).
"""
import numpy as np
import pytest

from helpers import load_golden, reference_pytsc

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5


@pytest.mark.parametrize("case", ["hangzhou_4_4__lf_pressure_select", "syn_1x1__pm_queue_switch",
                                  "manhattan_16_3__lf_queue_select"])
def test_batched_env_replays_reference(cuda_lib, case):
    import torch
    from pytsc_b200 import BatchedTrafficSignalNetwork
    g = load_golden(case)
    kw = {k: dict(v) for k, v in g["kwargs"].items()}
    B = 5
    env = BatchedTrafficSignalNetwork(g["scenario"], n_replicas=B, device=0, **kw)
    assert env.n_agents == len(g["signal_ids"])
    assert env.get_observation_size() == g["obs"].shape[-1] and env.get_action_size() == g["mask"].shape[-1]
    obs, mask = env.reset()
    assert np.array_equal(mask[B - 1].cpu().numpy(), g["mask0"])
    T = int(g["n_steps"])
    for t in range(T):
        act = torch.from_numpy(np.repeat(g["actions"][t][None], B, 0).astype(np.int32)).cuda()
        r, done, info = env.step(act)
        assert done == ((t + 1) % env.episode_limit == 0)
        assert np.array_equal(env.get_observations()[B - 1].cpu().numpy().astype(np.float64), g["obs"][t])
        assert np.array_equal(env.get_action_mask()[0].cpu().numpy(), g["mask"][t])
        np.testing.assert_allclose(r.cpu().numpy(), np.full(B, g["reward_global"][t]), rtol=REL_TOL)
        np.testing.assert_allclose(env.get_rewards()[2].cpu().numpy(), g["reward"][t], rtol=REL_TOL)
        assert float(info["n_queued"][1]) == g["metrics"][t][0]
        env.restart()
    env.check()
    m = env.all_reduce_episode_metrics()
    assert m["replicas"] == B
    assert m["average_travel_time"] == pytest.approx(g["sim"][T - 1][1], rel=1e-12)
    assert m["finished_vehicles"] == B * g["sim"][T - 1][3]
    env.close()


def test_batched_env_in_kernel_fixed_time(cuda_lib):
    """controller="fixed_time" (in-kernel FixedTimeController) == feeding the same
    controller's actions from outside."""
    import torch
    from pytsc_b200 import BatchedTrafficSignalNetwork
    kw = dict(signal=dict(observation_space="lane_features", reward_function="queue_length",
                          action_space="phase_selection", round_robin=False))
    a = BatchedTrafficSignalNetwork("hangzhou_4_4", n_replicas=2, **kw)
    b = BatchedTrafficSignalNetwork("hangzhou_4_4", n_replicas=2, **kw)
    cur = torch.zeros((2, a.n_agents), dtype=torch.int64, device="cuda")
    top = torch.zeros_like(cur)
    for t in range(80):
        green = (cur % 2 == 0)
        nxt = torch.where(green & (top < 25), cur, (cur + 1) % 16)
        top = torch.where(nxt == cur, top + 5, torch.full_like(top, 5))
        cur = nxt
        a.step(controller="fixed_time", green_time=25)
        b.step(cur.to(torch.int32))
        assert torch.equal(a.get_observations(), b.get_observations()), t
        assert torch.equal(a.get_rewards(), b.get_rewards()), t
        assert torch.equal(a.get_action_mask(), b.get_action_mask()), t
    a.check(); b.check()
    a.close(); b.close()


@pytest.mark.parametrize("case", ["hangzhou_4_4__lf_pressure_select", "hangzhou_4_4__pm_queue_select_rr",
                                  "jinan_3_4__lf_queue_select"])
def test_reference_facade_on_device(cuda_lib, case):
    pytsc = reference_pytsc()
    if pytsc is None:
        pytest.skip("reference pytsc not importable on this machine")
    g = load_golden(case)
    kw = {k: dict(v) for k, v in g["kwargs"].items()}
    kw["gpu"] = dict(n_replicas=2, view_replica=1, vehicle_capacity=1280)
    net = pytsc.TrafficSignalNetwork(g["scenario"], "gpu", **kw)
    assert np.array_equal(np.asarray(net.get_action_mask(), np.uint8), g["mask0"])
    for t in range(int(g["n_steps"])):
        r, done, info = net.step([int(a) for a in g["actions"][t]])
        assert r == pytest.approx(g["reward_global"][t], rel=1e-12)
        np.testing.assert_allclose(np.asarray(net.get_rewards(), np.float64), g["reward"][t], rtol=1e-12)
        assert np.array_equal(np.asarray(net.get_action_mask(), np.uint8), g["mask"][t])
        assert np.array_equal(np.asarray(net.get_observations(), np.float64), g["obs"][t])
        assert np.array_equal(np.asarray(net.get_state(), np.float64), g["state"][t])
        assert info["n_queued"] == g["metrics"][t][0]
        assert info["average_travel_time"] == pytest.approx(g["sim"][t][1], rel=1e-12)
        if done:
            net.restart()
    net.simulator.close_simulator()


def test_batched_epymarl_wrapper_replays_reference(cuda_lib):
    """BASELINE config 3 shape: Jinan 3x4, queue reward, the smac-style API
    (epymarl.py:96-111) with a replica dimension -- common reward = global / n_agents."""
    import torch
    from pytsc_b200 import BatchedEPyMARLTrafficSignalNetwork
    g = load_golden("jinan_3_4__lf_queue_select")
    kw = {k: dict(v) for k, v in g["kwargs"].items()}
    B = 6
    env = BatchedEPyMARLTrafficSignalNetwork(map_name=g["scenario"], simulator_backend="gpu", n_replicas=B, **kw)
    info = env.get_env_info()
    assert info["n_agents"] == 12 and info["n_actions"] == g["mask"].shape[-1] and info["obs_shape"] == g["obs"].shape[-1]
    assert info["agents"] == [str(x) for x in g["signal_ids"]]
    obs, state = env.reset()
    assert np.array_equal(env.get_avail_actions()[0].cpu().numpy(), g["mask0"])
    for t in range(int(g["n_steps"])):
        act = torch.from_numpy(np.repeat(g["actions"][t][None], B, 0).astype(np.int32)).cuda()
        obs, reward, over, truncated, _ = env.step(act)
        assert not truncated and over == ((t + 1) % env.episode_limit == 0)
        assert np.array_equal(obs[B - 1].cpu().numpy().astype(np.float64), g["obs"][t])
        assert np.array_equal(env.get_state()[1].cpu().numpy().astype(np.float64), g["state"][t])
        assert np.array_equal(env.get_avail_actions()[2].cpu().numpy(), g["mask"][t])
        np.testing.assert_allclose(reward.cpu().numpy(), np.full(B, g["reward_global"][t] / 12), rtol=REL_TOL)
        np.testing.assert_allclose(env.get_local_rewards()[3].cpu().numpy(), g["reward"][t], rtol=REL_TOL)
        if over:
            env.reset()
    env.tsc_env.check()
    loc = BatchedEPyMARLTrafficSignalNetwork(map_name=g["scenario"], n_replicas=2, common_reward=False, **kw)
    act = torch.from_numpy(np.repeat(g["actions"][0][None], 2, 0).astype(np.int32)).cuda()
    _, reward, _, _, _ = loc.step(act)
    np.testing.assert_allclose(reward[0].cpu().numpy(), g["reward"][0], rtol=REL_TOL)
    env.close(); loc.close()


def test_batched_env_rule_based_controllers(cuda_lib):
    """env.step(controller=...) for every in-kernel controller == asking tsc_controller_act first and
    feeding its phase indices back as external actions."""
    import torch
    from pytsc_b200 import BatchedTrafficSignalNetwork
    kw = dict(signal=dict(observation_space="lane_features", reward_function="queue_length",
                          action_space="phase_selection", round_robin=False))
    for name, ckw in (("sotl", dict(theta=2, mu=3, phi_min=10)), ("greedy", dict(seed=5)), ("max_pressure", dict(seed=5)),
                      ("random", dict(seed=9))):
        a = BatchedTrafficSignalNetwork("jinan_3_4", n_replicas=3, **kw)
        b = BatchedTrafficSignalNetwork("jinan_3_4", n_replicas=3, **kw)
        for t in range(50):
            acts = b.controller_actions(name, **ckw)
            a.step(controller=name, **ckw)
            b.step(acts, controller="phase_index")
            assert torch.equal(a.get_observations(), b.get_observations()), (name, t)
            assert torch.equal(a.get_action_mask(), b.get_action_mask()), (name, t)
        a.check(); b.check()
        a.close(); b.close()


@pytest.mark.parametrize("scenario", ["hangzhou_4_4", "jinan_3_4", "manhattan_16_3"])
def test_graph_metrics_against_reference_metrics_parser(cuda_lib, scenario):
    """density_map (retrieve kernel) and mst (tsc_max_spanning_tree) through the unmodified reference facade on the gpu
    backend, against the reference's own MetricsParser (backends/cityflow/metrics.py:170-209) running its CityFlow
    backend over the oracle engine, in lock-step under the same actions."""
    import random
    from scipy.sparse.csgraph import connected_components
    pytsc = reference_pytsc()
    if pytsc is None:
        pytest.skip("reference pytsc not importable on this machine")
    import sys
    from pytsc_b200 import compat
    from oracle.engine import Engine as OracleEngine
    import pytsc.backends.cityflow.simulator as cf_sim
    cf_sim.cityflow.Engine = OracleEngine if getattr(sys.modules.get("cityflow"), "__pytsc_b200_stub__", False) else cf_sim.cityflow.Engine
    from bench import materialise_reference_scenario
    import tempfile
    import pytsc.backends.cityflow.config as cf_config
    import pytsc.common.config as base_config
    kw = dict(cityflow=dict(flow_rate_type="constant"), signal=dict(observation_space="lane_features", reward_function="queue_length",
                                                                   action_space="phase_selection", round_robin=False))
    root = tempfile.mkdtemp(prefix="tsc_refscn_")
    name = materialise_reference_scenario(dict(scenario=scenario, kw=kw), root)
    old = (base_config.CONFIG_DIR, cf_config.CONFIG_DIR)
    base_config.CONFIG_DIR, cf_config.CONFIG_DIR = root, root + "/cityflow"
    try:
        ref = pytsc.TrafficSignalNetwork(name, "cityflow", **kw)
    finally:
        base_config.CONFIG_DIR, cf_config.CONFIG_DIR = old
    gkw = dict(kw, gpu=dict(n_replicas=2, view_replica=1, vehicle_capacity=1600))
    net = pytsc.TrafficSignalNetwork(scenario, "gpu", **gkw)
    rng = random.Random(4)
    for t in range(60):
        mask = ref.get_action_mask()
        acts = [rng.choice([k for k, m in enumerate(row) if m]) for row in mask]
        ref.step(acts)
        net.step(acts)
        if t % 6 == 5:
            want, got = np.asarray(ref.metrics.density_map), np.asarray(net.metrics.density_map)
            assert want.shape == got.shape
            if t == 59:
                assert want.max() > 1e-3      # vehicles on the connecting roads by now: not just the 1e-6 adjacency term
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-15)
            tw, tg = np.asarray(ref.metrics.mst), np.asarray(net.metrics.mst)
            sym = lambda m: np.triu(m + m.T - np.diag(np.diag(m)))
            assert tw.sum() == pytest.approx(tg.sum(), rel=1e-12)                      # same (maximum) tree weight
            assert np.count_nonzero(sym(tg)) == np.count_nonzero(sym(tw))              # same number of edges
            assert ((sym(tg) != 0) <= (np.triu(want + want.T) != 0)).all()             # edges of the signal graph only
            n_comp, _ = connected_components(sym(tg) != 0, directed=False)
            assert n_comp == len(tg) - np.count_nonzero(sym(tg))                       # a forest
    net.simulator.close_simulator()


def test_batched_env_graph_metrics(cuda_lib):
    """BatchedTrafficSignalNetwork.get_density_map / get_mst: every replica, device tensors; equal to the plugin view."""
    import torch
    from pytsc_b200 import BatchedTrafficSignalNetwork
    kw = dict(signal=dict(observation_space="lane_features", reward_function="queue_length",
                          action_space="phase_selection", round_robin=False), gpu=dict(vehicle_capacity=1200))
    env = BatchedTrafficSignalNetwork("jinan_3_4", n_replicas=5, lane_outputs=True, **kw)
    for _ in range(50):
        env.step(controller="fixed_time", green_time=25)
    dm, mst = env.get_density_map(), env.get_mst()
    assert dm.shape == (5, 12, 12) and mst.shape == (5, 12, 12) and dm.dtype == torch.float64
    assert torch.equal(dm, dm.transpose(1, 2)) and torch.equal(dm[0], dm[4]) and torch.equal(mst[0], mst[3])
    occ = env.out["lane_occupancy"][0].double().cpu().numpy()
    ids = list(env.parsed_network.traffic_signals.keys())
    nl, lanes = env.parsed_network.neighbors_lanes, env.scenario.lane_ids
    want = np.zeros((12, 12))
    for i, ti in enumerate(ids):
        for j, tj in enumerate(ids):
            ls = (nl.get(ti) or {}).get(tj)
            if ls:
                want[i, j] = np.clip(sum(occ[lanes.index(l)] for l in ls) / len(ls), 0, 1)
    want = (want + want.T) / 2 + 1e-6 * np.asarray(env.parsed_network.adjacency_matrix)
    np.testing.assert_allclose(dm[0].cpu().numpy(), want, rtol=1e-6)      # lane_occupancy output is fp32
    m = mst[0].cpu().numpy()
    assert np.count_nonzero(m) == 11 and (m <= 0).all()
    env.close()
