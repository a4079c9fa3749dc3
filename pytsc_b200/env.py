"""Batched environment: B replicas of one scenario, device tensors in and out.

The method names follow ``pytsc.TrafficSignalNetwork`` (``pytsc/__init__.py:
17-182``) and the smac-style wrapper ``EPyMARLTrafficSignalNetwork``
(``pytsc/wrappers/epymarl.py:11-111``), with a leading replica dimension:

    env = BatchedTrafficSignalNetwork("hangzhou_4_4", n_replicas=4096,
                                      signal={"reward_function": "max_pressure"})
    obs, mask = env.reset()
    reward, done, info = env.step(actions)        # actions: int32 [B, A] on the device
    obs, mask, rewards = env.get_observations(), env.get_action_mask(), env.get_rewards()

One ``step`` is one CUDA launch (``tsc_env_step``): phase program, delta_time
engine ticks, Retriever reductions, per-signal stats, rewards, masks and
observations.  Replicas shard across GPUs by giving each rank its own
instance (``device=LOCAL_RANK``); nothing on the step path communicates.
``all_reduce_episode_metrics`` is the one collective (NCCL, episode end).
"""
from __future__ import annotations

import numpy as np

from .backend.config import Config, DisruptedConfig
from .backend.network_parser import NetworkParser
from .binding import CONTROLLERS, Engine, sotl_arg
from .scenario import compile_scenario, derive_vehicle_capacity

def shard_replicas(total_replicas: int, world_size: int, rank: int):
    """Contiguous block of replicas owned by ``rank``: (first, count).  Replicas are
    independent scenario instances, so this is the whole partitioning scheme."""
    base, extra = divmod(int(total_replicas), int(world_size))
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def reduce_episode_metrics(vec, group=None):
    """SUM all-reduce of the episode-metric vector ``[sum ATT, finished, running, sum reward,
    sum queue, replica-steps, replicas]`` over the process group (NCCL on GPUs, gloo in the CPU
    tests); a no-op without an initialised group.  Returns the global metrics as a dict."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    v = vec.tolist()
    n, steps = max(v[6], 1.0), max(v[5], 1.0)
    return {"average_travel_time": v[0] / n, "finished_vehicles": v[1], "running_vehicles": v[2],
            "mean_global_reward": v[3] / steps, "mean_n_queued": v[4] / steps, "replicas": int(v[6])}


STEP_OUTPUTS = ("obs", "state", "reward", "reward_global", "mask", "sim", "metrics", "err")
LANE_OUTPUTS = ("lane_count", "lane_queued", "lane_occupancy", "lane_mean_speed")


class BatchedTrafficSignalNetwork:
    def __init__(self, scenario, n_replicas=None, device=None, lane_outputs=False, **kwargs):
        # pytsc/__init__.py:21-33: disrupted=True selects the DisruptedConfig (mode, domain_class in kwargs)
        self.disrupted = bool(kwargs.get("disrupted", False))
        self.config = (DisruptedConfig if self.disrupted else Config)(scenario, **kwargs)
        gpu = self.config.gpu
        self.parsed_network = NetworkParser(self.config)
        # every flow file the config can draw (flow_rate_type random / sequential, DisruptedConfig) is compiled
        # into the scenario as a flow set; each replica gets the one drawn for it whenever its engine restarts
        self._flow_files = self.config.flow_file_universe()
        self._flow_index = {f: k for k, f in enumerate(self._flow_files)}
        if len(self._flow_files) > 1:
            paths = [self.config.resolve_flow_file(f) for f in self._flow_files]
            self.scenario = compile_scenario(self.config, self.parsed_network, flow_sets=paths)
        else:
            self.scenario = compile_scenario(self.config, self.parsed_network)
        self.n_replicas = int(n_replicas if n_replicas is not None else gpu.get("n_replicas", 1))
        dev = int(device if device is not None else gpu.get("device", 0))
        cap = int(gpu.get("vehicle_capacity", 0)) or derive_vehicle_capacity(self.scenario)
        self.engine = Engine(self.scenario, self.n_replicas, dev, cap)
        self.flow_sets = np.zeros(self.n_replicas, np.int32)
        self._host = None
        self.torch = self.engine.torch
        self.device = self.engine.device
        self.n_agents = self.engine.A
        names = STEP_OUTPUTS + (LANE_OUTPUTS if lane_outputs else ())
        self.out = self.engine.alloc_outputs(names)
        sim = self.config.simulator
        self.delta_time = int(sim["delta_time"])
        self._episode_ticks = int(sim["episode_limit"])
        self._sim_length = int(sim["sim_length"])
        self._wait = int(sim["initial_wait_time"])
        self.episode_count = 0
        self.hour_count = 0
        self._acc = self.torch.zeros(4, dtype=self.torch.float64, device=self.device)
        self.reset()

    # ---- sizes (pytsc/__init__.py:118-138) --------------------------------------------
    def get_action_size(self):
        return self.engine.dims["n_actions"]

    def get_observation_size(self):
        return self.engine.dims["obs_dim"]

    def get_state_size(self):
        return self.engine.dims["state_dim"]

    @property
    def episode_limit(self):
        return int(self._episode_ticks / self.delta_time)

    @property
    def sim_step(self):
        return self._tick - self._wait

    @property
    def episode_over(self):
        return self.sim_step > 0 and self.sim_step % self._episode_ticks == 0

    @property
    def is_terminated(self):
        return self.sim_step == self._sim_length

    # ---- lifecycle ----------------------------------------------------------------------
    def reset(self):
        """Engine back to tick 0 (a new ``cityflow.Engine`` in the reference,
        pytsc/__init__.py:164-176), programs on phase 0, first measurements."""
        self.engine.check()      # a replica that overflowed its vehicle capacity froze: never hand out its stale rows silently
        self.engine.reset(self._draw_flow_sets())
        self.engine.init_program(0)
        self._tick = 0
        if self._wait:
            self.engine.step(self._wait)
            self._tick = self._wait
        self.engine.retrieve(self.out)
        return self.out["obs"], self.out["mask"]

    def _draw_flow_sets(self):
        """One ``Config._set_flow_file()`` draw per replica (backends/cityflow/config.py:63-76, 146-169): what B
        independent reference environments would do for their new engines.  None with a single flow file."""
        if len(self._flow_files) <= 1:
            return None
        for b in range(self.n_replicas):
            self.config._set_flow_file()
            self.flow_sets[b] = self._flow_index[self.config.flow_file]
        return self.flow_sets

    def set_domain_class(self, domain_class):
        """epymarl.py:85-89 -> DisruptedConfig.set_domain_class: takes effect at the next engine restart."""
        self.config.set_domain_class(domain_class)

    def restart(self):
        """pytsc/__init__.py:164-176."""
        if self.episode_over:
            self.episode_count += 1
        if self.is_terminated:
            self.hour_count += 1
            self.reset()

    def close(self):
        self.engine.close()

    # ---- the step (pytsc/__init__.py:178-182) ---------------------------------------------
    def step(self, actions=None, controller=None, green_time=25, seed=0, theta=3, mu=4, phi_min=5):
        """``actions``: int32 [B, A] device tensor in the configured action space; or ``controller`` =
        one of pytsc's rule-based controllers (``pytsc/controllers/controllers.py``) evaluated inside
        the launch: ``"fixed_time"`` (green_time), ``"greedy"`` / ``"max_pressure"`` / ``"random"``
        (``seed`` of the tie-break stream), ``"sotl"`` (theta, mu, phi_min)."""
        if controller is None or controller == "external":
            self.engine.env_step(actions, self.out, n_ticks=self.delta_time, controller=0)
        elif controller == "phase_index":
            self.engine.env_step(actions, self.out, n_ticks=self.delta_time, controller=CONTROLLERS["phase_index"])
        else:
            arg = {"fixed_time": green_time, "sotl": sotl_arg(theta, mu, phi_min)}.get(controller, seed)
            self.engine.env_step(None, self.out, n_ticks=self.delta_time, controller=CONTROLLERS[controller], controller_arg=arg)
        self._tick += self.delta_time
        self._acc[0] += self.out["reward_global"].sum()
        self._acc[1] += self.out["metrics"][:, 0].sum()
        self._acc[2] += self.n_replicas
        return self.out["reward_global"], self.episode_over, self.get_env_info()

    # ---- the same step through HOST arrays (what a CPU-side trainer / the reference's callers hold) ---------
    def register_host_buffers(self, threads=0):
        """Allocate and register the host result arrays of ``step_host``: ``obs`` float32 [B, A, obs_dim],
        ``reward`` [B, A], ``mask`` uint8 [B, A, n_actions], ``reward_global`` [B] (numpy).  Lane-feature
        observations only.  ``threads``: host workers that finish the rows (0 = automatic)."""
        d = self.engine.dims
        self._host = {"obs": np.empty((d["B"], d["A"], d["obs_dim"]), np.float32),
                      "reward": np.empty((d["B"], d["A"]), np.float32),
                      "mask": np.empty((d["B"], d["A"], d["n_actions"]), np.uint8),
                      "reward_global": np.empty((d["B"],), np.float32)}
        self.engine.host_register(threads=threads, **self._host)
        return self._host

    def step_host(self, actions=None, controller=None, green_time=25, seed=0, theta=3, mu=4, phi_min=5):
        """``step`` with numpy int32 [B, A] actions in and the registered numpy arrays out (synchronous): one
        launch, < 1 KB per replica over PCIe, rows finished by host threads while the launch runs."""
        self.step_host_begin(actions, controller, green_time, seed, theta, mu, phi_min)
        return self.step_host_wait()

    def step_host_begin(self, actions=None, controller=None, green_time=25, seed=0, theta=3, mu=4, phi_min=5):
        """First half of ``step_host``: returns as soon as the launch is queued.  With two environments of B / 2
        replicas stepped alternately (``a.step_host_begin(..); b.step_host_wait(); policy(b); b.step_host_begin(..);
        a.step_host_wait(); ...``) the host policy and the row finishing of one half overlap the launch of the other
        (double-buffered sampling).  ``actions`` must stay untouched until ``step_host_wait``."""
        if self._host is None:
            self.register_host_buffers()
        if controller is None or controller == "external":
            self.engine.env_step_registered_begin(actions, n_ticks=self.delta_time, controller=0)
        elif controller == "phase_index":
            self.engine.env_step_registered_begin(actions, n_ticks=self.delta_time, controller=CONTROLLERS["phase_index"])
        else:
            arg = {"fixed_time": green_time, "sotl": sotl_arg(theta, mu, phi_min)}.get(controller, seed)
            self.engine.env_step_registered_begin(None, n_ticks=self.delta_time, controller=CONTROLLERS[controller], controller_arg=arg)

    def step_host_wait(self):
        """Second half of ``step_host``: the registered numpy arrays hold the step's results on return."""
        self.engine.env_step_registered_wait()
        self._tick += self.delta_time
        return self._host["reward_global"], self.episode_over, self._host

    def controller_actions(self, controller, scores=False, **kw):
        """The phase indices ``controller`` would choose in the current state (``tsc_controller_act``)."""
        arg = {"fixed_time": kw.get("green_time", 25),
               "sotl": sotl_arg(kw.get("theta", 3), kw.get("mu", 4), kw.get("phi_min", 5))}.get(controller, kw.get("seed", 0))
        return self.engine.controller_act(controller, arg, scores=scores)

    # getters return the tensors the last launch wrote (no copies, no syncs)
    def get_observations(self):
        return self.out["obs"]

    def get_state(self):
        return self.out["state"]

    def get_action_mask(self):
        return self.out["mask"]

    def get_reward(self):
        return self.out["reward_global"]

    def get_rewards(self):
        return self.out["reward"]

    def get_density_map(self):
        """``MetricsParser.density_map`` (backends/cityflow/metrics.py:170-199) per replica: float64 [B, A, A] on the
        device, computed by the retrieve kernel from the current state."""
        if not hasattr(self, "_dm"):
            self._dm = self.engine.alloc_outputs(["density_map"])
        self.engine.retrieve(self._dm)
        return self._dm["density_map"]

    def get_mst(self):
        """``MetricsParser.mst`` (metrics.py:202-209; common/utils.py:158-161) per replica: float64 [B, A, A] on the device,
        the maximum spanning tree of the density map (``tsc_max_spanning_tree``)."""
        return self.engine.max_spanning_tree(self.get_density_map())

    def get_env_info(self):
        """Per-replica step statistics (backends/cityflow/metrics.py:221-232) as tensors."""
        m, s = self.out["metrics"], self.out["sim"]
        return {"time_step": s[:, 2], "average_travel_time": s[:, 1], "n_queued": m[:, 0], "mean_speed": m[:, 1],
                "mean_delay": m[:, 2], "density": m[:, 3], "pressure": m[:, 4], "network_flow": m[:, 5],
                "episode_count": self.episode_count, "episode_limit": self.episode_limit,
                "error_flags": self.out["err"]}      # non-zero = that replica froze (capacity exceeded): its rows are stale

    def check(self):
        """Synchronise and raise if any replica overflowed its vehicle capacity."""
        self.engine.check()

    # ---- multi-GPU: the one collective ------------------------------------------------------
    def all_reduce_episode_metrics(self, group=None):
        """Global episode metrics over every rank's replicas: a single NCCL
        all-reduce(SUM) of a 7-element fp64 vector."""
        torch = self.torch
        self.engine.check()
        s = self.out["sim"]
        vec = torch.stack([s[:, 1].sum(), s[:, 3].sum(), s[:, 0].sum(),
                           self._acc[0], self._acc[1], self._acc[2],
                           torch.tensor(float(self.n_replicas), dtype=torch.float64, device=self.device)])
        return reduce_episode_metrics(vec, group)


class BatchedEPyMARLTrafficSignalNetwork:
    """``EPyMARLTrafficSignalNetwork`` (``pytsc/wrappers/epymarl.py:11-111``) with a leading replica
    dimension: the smac-style API MARL trainers drive (``reset`` -> ``obs, state``; ``step(actions)``
    -> ``obs, reward, episode_over, truncated, info``; ``get_avail_actions``), every array a device
    tensor of shape ``[B, ...]``.  ``common_reward`` + ``reward_scalarization="mean"`` give the
    global reward divided by the number of agents (epymarl.py:103-110), otherwise the local rewards."""

    def __init__(self, map_name="hangzhou_4_4", simulator_backend="gpu", n_replicas=None, device=None, **kwargs):
        if simulator_backend != "gpu":
            raise ValueError("BatchedEPyMARLTrafficSignalNetwork drives the gpu backend only")
        kwargs.pop("scenario", None)
        self.common_reward = kwargs.pop("common_reward", True)
        self.reward_scalarization = kwargs.pop("reward_scalarization", "mean")
        self.tsc_env = BatchedTrafficSignalNetwork(map_name, n_replicas=n_replicas, device=device, **kwargs)
        self.episode_limit = self.tsc_env.episode_limit

    def get_avail_actions(self):
        return self.tsc_env.get_action_mask()

    def get_obs(self):
        return self.tsc_env.get_observations()

    def get_obs_size(self):
        return self.tsc_env.get_observation_size()

    def get_state(self):
        return self.tsc_env.get_state()

    def get_state_size(self):
        return self.tsc_env.get_state_size()

    def get_local_rewards(self):
        return self.tsc_env.get_rewards()

    def get_total_actions(self):
        return self.tsc_env.get_action_size()

    def get_network_flow(self):
        return self.tsc_env.out["metrics"][:, 5]

    def get_stats(self):
        return self.tsc_env.get_env_info()

    def get_env_info(self):
        env = self.tsc_env
        return {"agents": list(env.parsed_network.traffic_signals.keys()), "episode_limit": self.episode_limit,
                "n_actions": self.get_total_actions(), "adjacency_matrix": env.parsed_network.adjacency_matrix,
                "n_agents": env.n_agents, "obs_shape": self.get_obs_size(), "state_shape": self.get_state_size(),
                "n_replicas": env.n_replicas}

    def is_terminated(self):
        return self.tsc_env.is_terminated

    def sim_step(self):
        return self.tsc_env.sim_step

    def reset(self):
        """epymarl.py:96-101: count the episode, hand out the current observations, restart the
        simulation once its horizon is reached."""
        self.tsc_env.episode_count += 1
        obs, state = self.get_obs(), self.get_state()
        if self.tsc_env.episode_over:
            self.tsc_env.restart()
        return obs, state

    def step(self, actions):
        reward, episode_over, env_info = self.tsc_env.step(actions)
        if self.common_reward:
            if self.reward_scalarization == "mean":
                reward = reward / self.tsc_env.n_agents
        else:
            reward = self.get_local_rewards()
        return self.get_obs(), reward, episode_over, False, env_info

    def close(self):
        self.tsc_env.close()
