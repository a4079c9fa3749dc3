#!/usr/bin/env python
"""Generate the golden fixtures ``tests/golden/*.npz``.

Run in the BUILD CONTAINER only (needs the reference checkout):

    python tests/golden/make_golden.py [--ref /root/reference] [case ...]

What is recorded: the *unmodified* reference stack -- ``pytsc.TrafficSignalNetwork``
with its CityFlow backend plugin (Retriever, TrafficSignal, MetricsParser),
action spaces, observation spaces and reward functions -- is driven for T
env-steps with seeded random-over-mask actions, on top of the CPU oracle engine
standing in for the absent ``cityflow`` module (oracle/engine.py).  After every
``network.step(actions)`` the script stores what the reference reports:
observations, states, local and global rewards, action masks, per-lane
measurements, per-signal statistics, position-matrix windows, simulation
scalars and the step statistics, plus the oracle's vehicle snapshot at a few
checkpoints (which pins the oracle itself against regressions).

The fixtures are therefore exact outputs of the reference's own Python code for
everything downstream of the engine; the engine underneath is the restatement
(parity of the dynamics against real CityFlow is unpinned).
"""
import argparse
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# case -> (scenario, kwargs for TrafficSignalNetwork, env-steps, snapshot checkpoints)
CASES = {
    "syn_1x1__pm_queue_switch": ("syn_1x1", dict(
        cityflow=dict(flow_rate_type="constant", flow_file="syn_1x1__gaussian_600_flows.json"),
        signal=dict(observation_space="position_matrix", reward_function="queue_length",
                    action_space="phase_switch", round_robin=True)), 120),
    "syn_1x1__lf_pressure_select": ("syn_1x1", dict(
        cityflow=dict(flow_rate_type="constant", flow_file="syn_1x1__gaussian_700_flows.json"),
        signal=dict(observation_space="lane_features", reward_function="max_pressure",
                    action_space="phase_selection", round_robin=False)), 120),
    "hangzhou_4_4__lf_pressure_select": ("hangzhou_4_4", dict(
        signal=dict(observation_space="lane_features", reward_function="max_pressure",
                    action_space="phase_selection", round_robin=False)), 144),
    "hangzhou_4_4__pm_queue_select_rr": ("hangzhou_4_4", dict(
        signal=dict(observation_space="position_matrix", reward_function="queue_length",
                    action_space="phase_selection", round_robin=True)), 72),
    "hangzhou_4_4_5816__lf_queue_switch": ("hangzhou_4_4", dict(
        cityflow=dict(flow_file="anon_4_4_hangzhou_real_5816.json"),
        signal=dict(observation_space="lane_features", reward_function="queue_length",
                    action_space="phase_switch", round_robin=True)), 72),
    "jinan_3_4__lf_queue_select": ("jinan_3_4", dict(
        signal=dict(observation_space="lane_features", reward_function="queue_length",
                    action_space="phase_selection", round_robin=False)), 96),
    "manhattan_16_3__lf_queue_select": ("manhattan_16_3", dict(
        cityflow=dict(flow_file="anon_16_3_newyork_real.json", flow_rate_type="constant"),
        signal=dict(observation_space="lane_features", reward_function="queue_length",
                    action_space="phase_selection", round_robin=False)), 60),
    "manhattan_16_3__pm_pressure_switch": ("manhattan_16_3", dict(
        cityflow=dict(flow_file="anon_16_3_newyork_real.json", flow_rate_type="constant"),
        signal=dict(observation_space="position_matrix", reward_function="max_pressure",
                    action_space="phase_switch", round_robin=True)), 40),
}


# controller case -> (scenario, kwargs, controller name, controller kwargs, env-steps)
# Recorded from the reference's own controller classes (pytsc/controllers/controllers.py) run the way
# controllers/evaluate.py:112-137 runs them: action space forced to phase_selection, one controller object
# per signal fed the simulator's step measurements.
CONTROLLER_CASES = {
    "ctl_sotl__hangzhou_4_4": ("hangzhou_4_4", dict(signal=dict(observation_space="lane_features",
                               reward_function="queue_length", action_space="phase_selection", round_robin=False)),
                               "sotl", dict(theta=3, mu=4, phi_min=5), 144),
    "ctl_greedy__hangzhou_4_4": ("hangzhou_4_4", dict(signal=dict(observation_space="lane_features",
                                 reward_function="queue_length", action_space="phase_selection", round_robin=False)),
                                 "greedy", {}, 144),
    "ctl_max_pressure__jinan_3_4": ("jinan_3_4", dict(signal=dict(observation_space="lane_features",
                                    reward_function="max_pressure", action_space="phase_selection", round_robin=False)),
                                    "max_pressure", {}, 120),
    "ctl_max_pressure_rr__syn_1x1": ("syn_1x1", dict(
        cityflow=dict(flow_rate_type="constant", flow_file="syn_1x1__gaussian_700_flows.json"),
        signal=dict(observation_space="lane_features", reward_function="max_pressure",
                    action_space="phase_selection", round_robin=True)), "max_pressure", {}, 160),
    "ctl_random__syn_1x1": ("syn_1x1", dict(
        cityflow=dict(flow_rate_type="constant", flow_file="syn_1x1__gaussian_600_flows.json"),
        signal=dict(observation_space="lane_features", reward_function="queue_length",
                    action_space="phase_selection", round_robin=False)), "random", {}, 100),
    "ctl_fixed_time__jinan_3_4": ("jinan_3_4", dict(signal=dict(observation_space="lane_features",
                                  reward_function="queue_length", action_space="phase_selection", round_robin=False)),
                                  "fixed_time", dict(green_time=25), 100),
}


def setup_reference(ref):
    from pytsc_b200 import compat
    from oracle.engine import Engine
    compat.install_stubs(engine_factory=Engine)
    os.environ.setdefault("PYTSC_REFERENCE", ref)
    where = compat.find_reference_pytsc()
    if where is None:
        raise SystemExit("reference pytsc not found")
    import logging
    logging.disable(logging.CRITICAL)
    import pytsc  # noqa: F401
    return where


def record_case(name, scenario, kwargs, T, seed=0):
    from pytsc import TrafficSignalNetwork
    net = TrafficSignalNetwork(scenario, "cityflow", **kwargs)
    eng = net.simulator.engine
    lane_ids = eng.lane_ids                       # engine (roadnet) lane order
    ts = list(net.traffic_signals.values())
    A, vis = len(ts), net.config.signal["visibility"]
    rng = random.Random(seed)
    rec = {k: [] for k in ("actions", "obs", "state", "reward", "reward_global", "mask", "lane_count",
                           "lane_queued", "lane_occupancy", "lane_mean_speed", "sig_stats", "pos_in",
                           "pos_out", "sim", "metrics", "phase_changed")}
    snaps = {}
    mask0 = np.asarray(net.get_action_mask(), np.int64)
    mask = mask0
    for t in range(T):
        acts = [rng.choices(range(len(m)), weights=[int(x) for x in m])[0] for m in mask]
        r_glob, done, info = net.step(acts)
        rec["actions"].append(acts)
        rec["obs"].append(np.asarray(net.get_observations(), np.float64))
        rec["state"].append(np.asarray(net.observation_space.get_state(), np.float64))
        rec["reward"].append(np.asarray(net.get_rewards(), np.float64))
        rec["reward_global"].append(float(r_glob))
        mask = np.asarray(net.get_action_mask(), np.int64)
        rec["mask"].append(mask)
        lm = net.simulator.step_measurements["lane"]
        rec["lane_count"].append([lm[l]["n_vehicles"] for l in lane_ids])
        rec["lane_queued"].append([lm[l]["n_queued"] for l in lane_ids])
        rec["lane_occupancy"].append([float(lm[l]["occupancy"]) for l in lane_ids])
        rec["lane_mean_speed"].append([float(lm[l]["mean_speed"]) for l in lane_ids])
        rec["sig_stats"].append([[s.n_queued, float(s.occupancy), float(s.mean_speed), float(s.mean_delay),
                                  float(s.outgoing_occupancy), float(s.pressure), float(s.time_on_phase),
                                  s.controller.current_phase_index] for s in ts])
        rec["phase_changed"].append([int(s.controller.program.phase_changed) for s in ts])
        rec["pos_in"].append([s.inc_position_matrices[l] for s in ts for l in s.incoming_lanes])
        rec["pos_out"].append([s.out_position_matrices[l] for s in ts for l in s.outgoing_lanes])
        sm = net.simulator.step_measurements["sim"]
        rec["sim"].append([sm["n_vehicles"], sm["average_travel_time"], sm["time_step"],
                           eng.get_finished_vehicle_count()])
        st = net.metrics.get_step_stats()
        rec["metrics"].append([st["n_queued"], st["mean_speed"], st["mean_delay"], st["density"], st["pressure"],
                               st["network_flow"], float(net.metrics.flickering_signal),
                               float(net.metrics.norm_mean_speed)])
        if t in (T // 4, T // 2, T - 1):
            s = eng.snapshot()
            for k in ("uid", "drivable", "distance", "speed"):
                snaps[f"snap{t}_{k}"] = s[k]
    out = dict(
        actions=np.asarray(rec["actions"], np.int32), obs=np.asarray(rec["obs"]), state=np.asarray(rec["state"]),
        reward=np.asarray(rec["reward"]), reward_global=np.asarray(rec["reward_global"]),
        mask0=mask0.astype(np.uint8), mask=np.asarray(rec["mask"], np.uint8),
        lane_count=np.asarray(rec["lane_count"], np.int32), lane_queued=np.asarray(rec["lane_queued"], np.int32),
        lane_occupancy=np.asarray(rec["lane_occupancy"]), lane_mean_speed=np.asarray(rec["lane_mean_speed"]),
        sig_stats=np.asarray(rec["sig_stats"], np.float64), phase_changed=np.asarray(rec["phase_changed"], np.uint8),
        pos_in=np.asarray(rec["pos_in"], np.float64), pos_out=np.asarray(rec["pos_out"], np.float64),
        sim=np.asarray(rec["sim"], np.float64), metrics=np.asarray(rec["metrics"], np.float64),
        lane_ids=np.asarray(lane_ids), signal_ids=np.asarray([s.id for s in ts]),
        scenario=np.asarray(scenario), kwargs=np.asarray(repr(kwargs)), n_steps=np.asarray(T),
        snap_steps=np.asarray([T // 4, T // 2, T - 1]), **snaps)
    assert out["pos_in"].shape[-1] == vis
    return out


def record_controller_case(name, scenario, kwargs, controller, ckw, T, seed=0):
    """The reference's rule-based controller in closed loop.  Per step: the mask each controller saw,
    the score the controller computes for every phase index (its own helper methods), the action it
    returned, then the outputs of ``network.step(actions)``."""
    from pytsc import TrafficSignalNetwork
    from pytsc.controllers import CONTROLLERS
    np.random.seed(seed)
    net = TrafficSignalNetwork(scenario, "cityflow", **kwargs)
    eng = net.simulator.engine
    lane_ids = eng.lane_ids
    ts = list(net.traffic_signals.values())
    ctl = [CONTROLLERS[controller](t, **ckw) for t in ts]
    P = max(t.controller.n_phases for t in ts)
    rec = {k: [] for k in ("mask", "scores", "actions", "cur", "time_on_phase", "obs", "reward_global", "lane_count",
                           "lane_queued", "sim")}
    snaps = {}
    MASKED = -2 ** 31
    for t in range(T):
        inp = net.simulator.step_measurements
        masks, scores, acts, cur, top = [], [], [], [], []
        for sig, c in zip(ts, ctl):
            m = [int(x) for x in sig.controller.get_allowable_phase_switches()]
            m += [0] * (P - len(m))
            sc = [0] * P
            green = sig.controller.current_phase_index in sig.controller.green_phase_indices
            if controller in ("greedy", "max_pressure"):
                f = c._compute_queue_for_phase if controller == "greedy" else c._compute_pressure_for_phase
                sc = [int(f(inp, p)) if (green and p < sig.controller.n_phases and m[p]) else MASKED for p in range(P)]
            elif controller == "sotl" and m[sig.controller.current_phase_index]:
                sc[0] = int(c._compute_flow_for_phase(inp, sig.controller.current_phase_index))
                sc[1] = int(c._compute_flow_for_phase(inp, sig.controller.next_green_phase_index))
            cur.append(sig.controller.current_phase_index)
            top.append(sig.controller.time_on_phase)
            masks.append(m)
            scores.append(sc)
            acts.append(int(c.get_action(inp)))
        r_glob, done, info = net.step(acts)
        rec["mask"].append(masks); rec["scores"].append(scores); rec["actions"].append(acts)
        rec["cur"].append(cur); rec["time_on_phase"].append(top)
        rec["obs"].append(np.asarray(net.get_observations(), np.float64))
        rec["reward_global"].append(float(r_glob))
        lm = net.simulator.step_measurements["lane"]
        rec["lane_count"].append([lm[l]["n_vehicles"] for l in lane_ids])
        rec["lane_queued"].append([lm[l]["n_queued"] for l in lane_ids])
        sm = net.simulator.step_measurements["sim"]
        rec["sim"].append([sm["n_vehicles"], sm["average_travel_time"], sm["time_step"], eng.get_finished_vehicle_count()])
        if t == T - 1:
            s = eng.snapshot()
            for k in ("uid", "drivable", "distance", "speed"):
                snaps[f"snap{t}_{k}"] = s[k]
    return dict(mask=np.asarray(rec["mask"], np.uint8), scores=np.asarray(rec["scores"], np.int64),
                actions=np.asarray(rec["actions"], np.int32), cur=np.asarray(rec["cur"], np.int32),
                time_on_phase=np.asarray(rec["time_on_phase"], np.int32), obs=np.asarray(rec["obs"]),
                reward_global=np.asarray(rec["reward_global"]), lane_count=np.asarray(rec["lane_count"], np.int32),
                lane_queued=np.asarray(rec["lane_queued"], np.int32), sim=np.asarray(rec["sim"], np.float64),
                lane_ids=np.asarray(lane_ids), signal_ids=np.asarray([s.id for s in ts]),
                scenario=np.asarray(scenario), kwargs=np.asarray(repr(kwargs)), controller=np.asarray(controller),
                controller_kwargs=np.asarray(repr(ckw)), n_steps=np.asarray(T), snap_steps=np.asarray([T - 1]), **snaps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("cases", nargs="*")
    args = ap.parse_args()
    setup_reference(args.ref)
    for name in (args.cases or list(CASES) + list(CONTROLLER_CASES)):
        if name in CONTROLLER_CASES:
            scenario, kwargs, controller, ckw, T = CONTROLLER_CASES[name]
            out = record_controller_case(name, scenario, kwargs, controller, ckw, T)
            path = os.path.join(HERE, name + ".npz")
            np.savez_compressed(path, **out)
            print(f"{name}: T={T} A={out['actions'].shape[1]} veh(end)={int(out['sim'][-1, 0])} "
                  f"switches={int((out['actions'] != out['cur']).sum())} -> {os.path.getsize(path) / 1024:.0f} KB")
            continue
        scenario, kwargs, T = CASES[name]
        out = record_case(name, scenario, kwargs, T)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: T={T} A={out['actions'].shape[1]} obs={out['obs'].shape[-1]} "
              f"veh(end)={int(out['sim'][-1, 0])} -> {os.path.getsize(path) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
