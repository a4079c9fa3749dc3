"""TEST INFRASTRUCTURE -- CPU restatement of pytsc's Python half of the hot path.

``PortEnv`` restates, on top of the CPU oracle engine (oracle/engine.py), what
one ``TrafficSignalNetwork.step(actions)`` of the reference's CityFlow backend
does (reference paths relative to the reference repo, ``pytsc/...``):

    apply phases         common/actions.py:99-108, 144-158; backends/cityflow/traffic_signal.py:51-59
    program bookkeeping  common/traffic_signal.py:83-109
    5 x next_step        backends/cityflow/simulator.py:80-89
    Retriever            backends/cityflow/retriever.py:20-112; common/utils.py:115-133
    per-signal stats     backends/cityflow/traffic_signal.py:101-141
    network metrics      backends/cityflow/metrics.py:24-167, 212-260
    rewards              common/reward.py:54-88, 102-136
    action masks         common/traffic_signal.py:329-361, 375-404; common/actions.py:119-131, 169-188
    observations/state   common/observations.py:72-160, 192-213, 305-329, 352-374; common/utils.py:91-112
    fixed-time control   controllers/controllers.py:39-54

It is deliberately written the way the reference works -- dictionaries keyed by
lane id, one ``get_vehicle_info`` call and two string->float conversions per
vehicle, Python lists padded through numpy -- because it serves two purposes:
the checker for the CUDA path at sizes where no golden fixture exists, and the
"port" CPU baseline of bench.py (the reference itself cannot travel to the GPU
box, and its engine, CityFlow, is not installable).  ``tests/test_port.py``
pins it against the golden fixtures, which were recorded from the reference's
own classes; every output must be identical.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU legs may import
this module.  The product package never does.
"""
from __future__ import annotations

import json
import os
import tempfile

import numpy as np

from pytsc_b200 import bundle
from pytsc_b200.backend.config import Config
from pytsc_b200.backend.network_parser import NetworkParser

from .engine import Engine

MAX_LANES_PER_DIRECTION = 6      # observations.py:57-61 / 227-231
MAX_LANE_SPEED = 15.0
MAX_LANE_LENGTH = 500
MAX_PHASES = 20
MAX_N_CONTROLLED_LANES = 16

_json_dirs = {}


def engine_config_file(config) -> str:
    """CityFlow engine cfg JSON (backends/cityflow/config.py:78-103) over plain
    JSON copies of the scenario's roadnet / flow bundle."""
    flow_path = config.create_and_save_cityflow_cfg()
    key = (config.cityflow_roadnet_file, flow_path, config.simulator["seed"])
    if key not in _json_dirs:
        d = tempfile.mkdtemp(prefix="tsc_port_")
        with open(os.path.join(d, "roadnet.json"), "w") as f:
            json.dump(bundle.load_roadnet(config.cityflow_roadnet_file), f)
        with open(os.path.join(d, "flow.json"), "w") as f:
            json.dump(bundle.load_flow(flow_path), f)
        cfg = dict(dir=d + os.sep, roadnetFile="roadnet.json", flowFile="flow.json",
                   interval=config.simulator["interval"], rlTrafficLight=config.simulator["rl_traffic_light"],
                   laneChange=config.simulator["lane_change"], seed=config.simulator["seed"], saveReplay=False)
        fn = os.path.join(d, "engine_cfg.json")
        with open(fn, "w") as f:
            json.dump(cfg, f)
        _json_dirs[key] = fn
    return _json_dirs[key]


def pad_list(values, size, pad_value=0):
    """common/utils.py:91-112 -- ``np.full(size, pad_value)`` takes the dtype of
    ``pad_value``: an int pad truncates float entries toward zero."""
    arr = np.asarray(np.array(values))
    if len(arr) >= size:
        return arr.tolist()
    out = np.full(size, pad_value)
    out[: len(arr)] = arr
    return out.tolist()


def vehicle_bin_index(n_bins, lane_length, position):
    """common/utils.py:115-133."""
    if position < 0:
        position = 0
    elif position > lane_length:
        position = lane_length
    bin_size = lane_length / n_bins
    idx = int(position // bin_size)
    return n_bins - 1 if idx >= n_bins else idx


class _Signal:
    """Program state + statistics of one traffic signal."""

    def __init__(self, ts_id, cfg):
        self.id = ts_id
        self.cfg = cfg
        self.phases = cfg["phases"]
        self.n_phases = cfg["n_phases"]
        self.green = cfg["green_phase_indices"]
        self.incoming_lanes = cfg["incoming_lanes"]
        self.outgoing_lanes = cfg["outgoing_lanes"]
        self.current_phase_index = 0      # set_initial_phase, common/traffic_signal.py:83-92
        self.time_on_phase = 0
        self.norm_time_on_phase = 0
        self.phase_changed = None         # undefined until the first update

    @property
    def current_phase(self):
        return self.phases[self.current_phase_index]

    def update_current_phase(self, idx):
        """common/traffic_signal.py:94-109."""
        if idx == self.current_phase_index:
            self.phase_changed = False
            self.time_on_phase += self.cfg["yellow_time"]
        else:
            self.phase_changed = True
            self.time_on_phase = self.cfg["yellow_time"]
        self.current_phase_index = idx
        self.norm_time_on_phase = self.time_on_phase / self.cfg["phases_min_max_times"][self.current_phase]["max_time"]

    def allowable_phase_switches(self):
        """common/traffic_signal.py:329-361 (free) and 375-404 (round robin)."""
        mask = [0] * self.n_phases
        cur, nxt = self.current_phase_index, (self.current_phase_index + 1) % self.n_phases
        if cur in self.green:
            mm = self.cfg["phases_min_max_times"][self.current_phase]
            if self.time_on_phase < mm["min_time"]:
                mask[cur] = 1
            elif self.time_on_phase < mm["max_time"]:
                mask[cur] = 1
                mask[nxt] = 1
            elif self.time_on_phase == mm["max_time"]:
                mask[nxt] = 1
            else:
                raise RuntimeError("time_on_phase beyond max_time")   # the reference hits breakpoint() here
        elif self.cfg["round_robin"]:
            mask[nxt] = 1
        else:
            for g in self.green:
                if g != cur - 1:
                    mask[g] = 1
        return mask


class PortEnv:
    """``TrafficSignalNetwork(scenario, "cityflow", **kwargs)`` restated."""

    def __init__(self, scenario, **kwargs):
        self.config = Config(scenario, **kwargs)
        self.parsed_network = NetworkParser(self.config)
        self.engine = Engine(engine_config_file(self.config), thread_num=1)
        sim, sig = self.config.simulator, self.config.signal
        for _ in range(sim["initial_wait_time"]):
            self.engine.next_step()
        self.visibility = sig["visibility"]
        self.v_size = sim["veh_size_min_gap"]
        self.lane_lengths = self.parsed_network.lane_lengths
        self.lane_max_speeds = self.parsed_network.lane_max_speeds
        self.signals = {}
        self.retrieve_step_measurements()
        for ts_id, cfg in self.parsed_network.traffic_signals.items():
            s = _Signal(ts_id, cfg)
            self.engine.set_tl_phase(ts_id, s.phases[0])     # backends/cityflow/traffic_signal.py:26-32
            self.signals[ts_id] = s
            self._update_stats(s)
        self.n_agents = len(self.signals)
        self.static_features = self._static_lane_features()
        self.episode_count = 0

    # ---- Retriever (backends/cityflow/retriever.py) -------------------------------------
    def _position_matrix(self, lane, vehicles):
        bins = int(self.lane_lengths[lane] / self.v_size)
        if bins > 0 and len(vehicles) > 0:
            mat = [-1.0] * bins
            for v in vehicles:
                info = self.engine.get_vehicle_info(v)
                b = vehicle_bin_index(bins, self.lane_lengths[lane], float(info["distance"]))
                if b is not None:
                    mat[b] += 1.0
                    mat[b] += float(info["speed"]) / self.lane_max_speeds[lane]
            if len(mat) < self.visibility:
                mat += [-1.0] * (self.visibility - len(mat))
        else:
            mat = [-1.0] * self.visibility
        return mat

    def retrieve_step_measurements(self):
        queued = self.engine.get_lane_waiting_vehicle_count()
        lane_vehicles = self.engine.get_lane_vehicles()
        speeds = self.engine.get_vehicle_speed()
        lanes = {}
        for lane, vehicles in lane_vehicles.items():
            n = len(vehicles)
            if n == 0:
                mean_speed = 0.0
            else:
                total = 0
                for v in vehicles:
                    total += speeds[v]
                mean_speed = total / n
            lanes[lane] = {
                "n_vehicles": n,
                "n_queued": queued[lane],
                "occupancy": n / (self.lane_lengths[lane] / self.v_size),
                "mean_speed": mean_speed,
                "position_matrix": self._position_matrix(lane, vehicles),
            }
        self.step_measurements = {
            "lane": lanes,
            "sim": {"n_vehicles": self.engine.get_vehicle_count(),
                    "average_travel_time": self.engine.get_average_travel_time(),
                    "time_step": self.engine.get_current_time()},
        }

    # ---- TrafficSignal.update_stats (backends/cityflow/traffic_signal.py:101-141) ---------
    def _update_stats(self, s):
        lanes = self.step_measurements["lane"]
        s.n_queued, s.occupancy, s.mean_speed, s.mean_delay = 0, 0, 0, 0
        s.inc_position_matrices, s.out_position_matrices = {}, {}
        for lane in s.incoming_lanes:
            r = lanes[lane]
            s.n_queued += r["n_queued"]
            s.occupancy += r["occupancy"]
            s.mean_speed += r["mean_speed"]
            s.mean_delay += 1 - r["mean_speed"] / self.lane_max_speeds[lane]
            s.inc_position_matrices[lane] = r["position_matrix"][-self.visibility:]
        s.occupancy /= len(s.incoming_lanes)
        s.mean_speed /= len(s.incoming_lanes)
        s.mean_delay /= len(s.incoming_lanes)
        s.outgoing_occupancy = 0
        for lane in s.outgoing_lanes:
            r = lanes[lane]
            s.outgoing_occupancy += r["occupancy"]
            s.out_position_matrices[lane] = r["position_matrix"][: self.visibility]
        s.outgoing_occupancy /= len(s.outgoing_lanes)
        s.stat_time_on_phase = s.norm_time_on_phase
        one_hot = [0] * s.n_phases
        one_hot[s.current_phase_index] = 1
        s.phase_id = np.asarray(one_hot)
        s.pressure = np.abs(s.occupancy - s.outgoing_occupancy).item()

    # ---- the step (pytsc/__init__.py:178-182) ---------------------------------------------
    @property
    def sim_step(self):
        return self.engine.get_current_time() - self.config.simulator["initial_wait_time"]

    @property
    def episode_over(self):
        st = self.sim_step
        return st % self.config.simulator["episode_limit"] == 0 if st > 0 else False

    @property
    def is_terminated(self):
        return self.sim_step == self.config.simulator["sim_length"]

    def step(self, actions, phase_indices=False):
        """``phase_indices=True``: ``actions`` are pytsc phase indices whatever the action
        space (what ``TSController.switch_phase`` receives)."""
        switch = self.config.signal["action_space"] == "phase_switch" and not phase_indices
        for i, s in enumerate(self.signals.values()):
            if switch:       # common/actions.py:144-158
                idx = (s.current_phase_index + 1) % s.n_phases if actions[i] == 1 else s.current_phase_index
            else:            # common/actions.py:99-108
                idx = actions[i]
            self.engine.set_tl_phase(s.id, s.phases[idx])
            s.update_current_phase(idx)
        for _ in range(self.config.simulator["delta_time"]):
            self.engine.next_step()
        self.retrieve_step_measurements()
        for s in self.signals.values():
            self._update_stats(s)
        return self.get_reward(), self.episode_over, self.get_env_info()

    # ---- metrics (backends/cityflow/metrics.py) -----------------------------------------------
    def step_stats(self):
        lanes = self.step_measurements["lane"]
        n_queued = 0
        for d in lanes.values():
            n_queued += d["n_queued"]
        total_vehicles = sum(d["n_vehicles"] for d in lanes.values())
        if total_vehicles == 0:
            mean_speed = 0.0
        else:
            mean_speed = sum(d["mean_speed"] * d["n_vehicles"] for d in lanes.values()) / total_vehicles
        density = sum(d["occupancy"] for d in lanes.values()).item() / len(lanes)
        norm_mean_speed = sum(d["mean_speed"] / self.lane_max_speeds[l] for l, d in lanes.items()) / len(lanes)
        pressure = np.sum([s.pressure for s in self.signals.values()]).item()
        return {
            "time_step": self.step_measurements["sim"]["time_step"],
            "average_travel_time": self.step_measurements["sim"]["average_travel_time"],
            "n_queued": n_queued, "mean_speed": mean_speed, "mean_delay": 1 - norm_mean_speed,
            "density": density, "pressure": pressure, "network_flow": density * norm_mean_speed,
        }

    @property
    def flickering_signal(self):
        return np.mean([s.phase_changed for s in self.signals.values()])

    @property
    def norm_mean_speed(self):
        lanes = self.step_measurements["lane"]
        return sum(d["mean_speed"] / self.lane_max_speeds[l] for l, d in lanes.items()) / len(lanes)

    def get_env_info(self):
        stats = self.step_stats()
        stats.update({"episode_count": self.episode_count,
                      "episode_limit": int(self.config.simulator["episode_limit"] / self.config.simulator["delta_time"])})
        return stats

    # ---- rewards (common/reward.py) ---------------------------------------------------------------
    def _metric(self, s):
        return s.n_queued if self.config.signal["reward_function"] == "queue_length" else s.pressure

    def get_reward(self):
        fc = self.config.misc["flickering_coef"]
        reward = 1e-6
        if self.config.signal["reward_function"] == "queue_length":
            reward += fc * self.flickering_signal
            reward += self.step_stats()["n_queued"]
            return -1 * reward
        reward -= fc * self.flickering_signal
        reward -= np.sum([s.pressure for s in self.signals.values()]).item()
        return reward

    def get_rewards(self):
        fc, gamma = self.config.misc["flickering_coef"], self.config.misc["reward_gamma"]
        khop = self.parsed_network.k_hop_neighbors
        local = {i: -fc * s.phase_changed - self._metric(s) - 1e-6 for i, s in self.signals.items()}
        out = {}
        for i in self.signals:
            out[i] = local[i]
            for k in range(1, len(self.signals)):
                for nb in khop[i].get(k, []):
                    out[i] += gamma**k * local[nb]
        return list(out.values())

    # ---- action masks (common/actions.py:119-131, 169-188) ---------------------------------------------
    def get_action_size(self):
        if self.config.signal["action_space"] == "phase_switch":
            return 2
        return max(s.n_phases for s in self.signals.values())

    def get_action_mask(self):
        masks = []
        for s in self.signals.values():
            allow = s.allowable_phase_switches()
            if self.config.signal["action_space"] == "phase_switch":
                nxt = (s.current_phase_index + 1) % s.n_phases
                masks.append([1 if allow[s.current_phase_index] else 0, 1 if allow[nxt] else 0])
            else:
                masks.append(pad_list(allow, self.get_action_size()))
        return masks

    # ---- observations (common/observations.py) ---------------------------------------------------------
    def _static_lane_features(self):
        feats = {}
        for lane in self.parsed_network.lanes:
            one_hot = [0.0] * MAX_LANES_PER_DIRECTION
            one_hot[self.parsed_network.lane_indices[lane]] = 1.0
            length = np.clip(self.lane_lengths[lane] / MAX_LANE_LENGTH, 0, 1)
            angle = np.clip(self.parsed_network.lane_angles[lane] / np.pi, -1, 1)
            vmax = np.clip(self.lane_max_speeds[lane] / MAX_LANE_SPEED, 0, 1)
            feats[lane] = [length, angle, vmax] + one_hot
        return feats

    def _lane_feature_vectors(self):
        out = []
        lanes = self.step_measurements["lane"]
        size = MAX_N_CONTROLLED_LANES * 12 + MAX_PHASES
        for s in self.signals.values():
            vec = []
            for lane in s.incoming_lanes:
                vec.extend(self.static_features[lane])
                r = lanes[lane]
                vec.extend([r["n_queued"], r["occupancy"], r["mean_speed"]])
            vec = pad_list(vec, size - MAX_PHASES, -1)
            vec.extend(pad_list(s.phase_id, MAX_PHASES))
            out.append(vec)
        return out

    def get_state(self):
        return self._lane_feature_vectors()          # observations.py:192-213 / 352-374

    def get_observations(self):
        if self.config.signal["observation_space"] == "lane_features":
            return self._lane_feature_vectors()      # observations.py:305-329 (no dropped lanes)
        out = []                                     # observations.py:140-160
        size = MAX_N_CONTROLLED_LANES * (self.visibility + 9) + MAX_PHASES
        for s in self.signals.values():
            vec = []
            for lane, mat in s.inc_position_matrices.items():
                vec.extend(self.static_features[lane])
                # _add_gaussian_noise at std 0 (observations.py:72-88): keeps entries > 0, clipped to [0, 1]
                vec.extend([np.clip(v + 0.0, 0.0, 1.0) for v in mat if v > 0])
            vec = pad_list(vec, size - MAX_PHASES, -1)
            vec.extend(pad_list(s.phase_id, MAX_PHASES, -1))
            out.append(vec)
        return out

    # ---- fixed-time controller (controllers/controllers.py:39-54) -------------------------------------------
    def fixed_time_actions(self, green_time=25):
        acts = []
        switch = self.config.signal["action_space"] == "phase_switch"
        for s in self.signals.values():
            if s.current_phase_index in s.green and s.time_on_phase < green_time:
                idx = s.current_phase_index
            else:
                idx = (s.current_phase_index + 1) % s.n_phases
            acts.append((1 if idx != s.current_phase_index else 0) if switch else idx)   # actions.py:200-211
        return acts
