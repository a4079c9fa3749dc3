"""Config for the `gpu` backend.

Mirrors the reference's three-level merge -- defaults <- scenario ``config.yaml``
<- keyword overrides (``pytsc/common/config.py:37-76``) -- and the CityFlow
backend's file handling (``pytsc/backends/cityflow/config.py:29-76``): the `gpu`
backend consumes the *same* ``cityflow:`` YAML section and the same roadnet /
flow files, plus an optional ``gpu:`` section (replica count, device,
capacities).  Default values are those of
``pytsc/scenarios/default/config.yaml``.
"""
from __future__ import annotations

import copy
import os
import random
from itertools import cycle

import yaml

DEFAULTS = {
    "network": {"network_type": "synthetic", "control_scheme": "decentralized"},
    "signal": {
        "action_space": "phase_switch", "observation_space": "position_matrix",
        "reward_function": "queue_length", "yellow_time": 5, "min_green_time": 5,
        "max_green_time": 60, "visibility": 10, "input_n_avg": 1, "round_robin": True,
        "obs_dropout_prob": 0.0,
    },
    "cityflow": {
        "seed": 0, "thread_num": 1, "interval": 1.0, "rl_traffic_light": True,
        "lane_change": False, "delta_time": 5, "episode_limit": 360, "initial_wait_time": 0,
        "vehicle_length": 5, "veh_size_min_gap": 7.5, "save_replay": False,
        "flow_rate_type": "constant", "sim_length": 3600,
    },
    "gpu": {
        "n_replicas": 1,          # B: scenario replicas stepped in lockstep on this device
        "device": 0,
        "vehicle_capacity": 0,    # 0 = a bound derived from the scenario (scenario.derive_vehicle_capacity); smaller = faster
        "reference_exact": True,  # reproduce pad_list's integer truncation in observations (utils.py:91-112)
    },
    "misc": {
        "max_wait_time": 1000, "pad_value": 0.0, "save_trip_info": False, "flickering_coef": 0.01,
        "max_hops": 1, "reward_gamma": 0.9, "return_agent_stats": False, "return_lane_stats": False,
        "grid_reduction_factor": 5.0,
    },
}


def recursively_update_dict(d, u):
    for k, v in u.items():
        if isinstance(v, dict):
            d[k] = recursively_update_dict(d.get(k, {}) or {}, v)
        else:
            d[k] = v
    return d


def scenario_search_path():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    paths = [os.path.join(here, "scenarios")]
    env = os.environ.get("PYTSC_B200_SCENARIOS")
    if env:
        paths = env.split(os.pathsep) + paths
    try:  # the reference's own scenario tree, when pytsc is importable
        import importlib.util
        spec = importlib.util.find_spec("pytsc")
        if spec and spec.submodule_search_locations:
            paths.append(os.path.join(list(spec.submodule_search_locations)[0], "scenarios", "cityflow"))
    except Exception:
        pass
    return paths


def find_scenario_dir(scenario):
    if os.path.isdir(scenario):
        return os.path.abspath(scenario)
    for p in scenario_search_path():
        d = os.path.join(p, scenario)
        if os.path.isdir(d):
            return d
    raise FileNotFoundError(f"scenario {scenario!r} not found in {scenario_search_path()}")


def resolve_data_file(directory, name):
    """``name`` as given, or with .gz / .npz in place of .json."""
    base = os.path.join(directory, name)
    stem = os.path.splitext(base)[0]
    for cand in (base, base + ".gz", stem + ".npz", stem + ".json", stem + ".json.gz"):
        if os.path.exists(cand):
            return cand
    raise FileNotFoundError(base)


class Config:
    """Same attribute protocol as the reference's ``Config``: ``.network``,
    ``.signal``, ``.misc``, ``.simulator`` (= the ``cityflow`` section), plus
    ``.gpu``."""

    def __init__(self, scenario, debug=False, **kwargs):
        self.debug = debug
        self.scenario = scenario
        self._additional_config = kwargs
        self.dir = find_scenario_dir(scenario)
        cfg = copy.deepcopy(DEFAULTS)
        scenario_file = os.path.join(self.dir, "config.yaml")
        if os.path.exists(scenario_file):
            with open(scenario_file, "r") as f:
                recursively_update_dict(cfg, yaml.safe_load(f) or {})
        overrides = {k: v for k, v in kwargs.items() if isinstance(v, dict)}
        recursively_update_dict(cfg, overrides)
        self.network = cfg["network"]
        self.signal = cfg["signal"]
        self.misc = cfg["misc"]
        self.simulator = cfg["cityflow"]
        self.gpu = cfg["gpu"]
        self.full = cfg
        random.seed(self.simulator["seed"])
        self.cityflow_roadnet_file = resolve_data_file(self.dir, self.simulator["roadnet_file"])
        self.flow_files_cycle = cycle(self.simulator.get("flow_files", []))
        self.flow_file = None
        assert self.signal["yellow_time"] == self.simulator["delta_time"], \
            "Delta time and yellow times must be fixed to 5 seconds."   # cityflow/config.py:58-61
        if self.simulator["lane_change"]:
            raise NotImplementedError("gpu backend: lane_change is not supported (pytsc default is False)")
        if not self.simulator["rl_traffic_light"]:
            raise NotImplementedError("gpu backend: rl_traffic_light must be True (pytsc default)")
        if float(self.simulator["interval"]) != 1.0:
            raise NotImplementedError("gpu backend: interval must be 1.0 (pytsc default)")
        # reference features the device path does not implement are refused, not ignored
        if float(self.signal.get("obs_noise_std", 0.0) or 0.0) != 0.0:
            raise NotImplementedError("gpu backend: signal.obs_noise_std (observations.py:70-88) is not supported")
        if float(self.signal.get("obs_dropout_prob", 0.0) or 0.0) != 0.0:
            raise NotImplementedError("gpu backend: signal.obs_dropout_prob is not supported")
        if self.network.get("control_scheme", "decentralized") != "decentralized":
            raise NotImplementedError("gpu backend: network.control_scheme must be 'decentralized' "
                                      "(CentralizedActionSpace enumerates the joint action set)")

    def _set_flow_file(self):
        """``cityflow/config.py:63-76``."""
        self.flow_rate_type = self.simulator.get("flow_rate_type", "constant")
        if self.flow_rate_type == "constant":
            self.flow_file = self.simulator["flow_file"]
        elif self.flow_rate_type == "random":
            self.flow_file = random.choice(self.simulator["flow_files"])
        elif self.flow_rate_type == "sequential":
            self.flow_file = next(self.flow_files_cycle)
        else:
            raise ValueError("Flow files order is not supported. "
                             "Flow files order must be `random` or `constant`")
        return resolve_data_file(self.dir, self.flow_file)

    # the CityFlow backend writes an engine cfg JSON here; the gpu backend has no
    # such file -- the method is kept so that callers can treat both alike
    def create_and_save_cityflow_cfg(self):
        self.cityflow_flow_file = self._set_flow_file()
        return self.cityflow_flow_file

    def flow_file_universe(self):
        """Every flow file ``_set_flow_file`` can pick, in a fixed order: the batched environment compiles them
        all into one scenario (``n_flow_sets``) and gives each replica the one drawn for it at reset."""
        if self.simulator.get("flow_rate_type", "constant") == "constant":
            return [self.simulator["flow_file"]]
        return list(dict.fromkeys(self.simulator["flow_files"]))

    def resolve_flow_file(self, name):
        return resolve_data_file(self.dir, name)


class DisruptedConfig(Config):
    """``DisruptedConfig`` (``pytsc/backends/cityflow/config.py:106-175``): the ``cityflow:`` section holds, per
    ``mode`` (train / test), ``{domain: {disruption value: [flow files]}}``; every new engine runs a flow file
    drawn from a (drawn or fixed) domain class, found under ``<mode>/<domain>/<value>/`` in the scenario
    directory."""

    def __init__(self, scenario, mode="train", debug=False, **kwargs):
        self.mode = mode
        self.domain_class = kwargs.get("domain_class", None)
        super().__init__(scenario, debug=debug, **kwargs)
        self.domains = list(self.simulator[mode].keys())
        self.disrup_values = {d: list(self.simulator[mode][d].keys()) for d in self.domains}
        self.domain_classes = [(d, v) for d in self.domains for v in self.disrup_values[d]]
        self.current_domain_class = None
        random.seed(self.simulator["seed"])

    def _set_flow_file(self):
        self.flow_rate_type = self.simulator.get("flow_rate_type", "constant")
        if self.domain_class is None:
            domain = random.choice(self.domains)
            value = random.choice(self.disrup_values[domain])
        else:
            domain, value = self.domain_class[0], self.domain_class[1]
        self.current_domain_class = self.domain_classes.index((domain, value))
        flow_file = random.choice(self.simulator[self.mode][domain][value])
        self.flow_file = os.path.join(self.mode, domain, value, flow_file)
        return resolve_data_file(self.dir, self.flow_file)

    def set_domain_class(self, domain_class):
        self.domain_class = domain_class

    def flow_file_universe(self):
        return [os.path.join(self.mode, d, v, f) for d, v in self.domain_classes for f in self.simulator[self.mode][d][v]]
