"""Simulator plugin of the `gpu` backend.

Same surface as the reference's CityFlow ``Simulator``
(``pytsc/backends/cityflow/simulator.py:8-95``; abstract base
``pytsc/common/simulator.py:4-65``): ``start_simulator``, ``simulator_step``,
``close_simulator``, ``retrieve_step_measurements``, ``is_terminated``,
``sim_step``, ``sim_time`` and the ``step_measurements`` dictionary.  Instead
of one ``cityflow.Engine`` it owns B replicas on one CUDA device behind the C
ABI (``binding.Engine``); pytsc's single-environment modules see replica
``gpu.view_replica`` (all replicas receive the same phases through this view,
so they stay in lock-step).  The batched tensor API for RL lives in
``pytsc_b200.env.BatchedTrafficSignalNetwork``.
"""
from __future__ import annotations

import numpy as np

from ..binding import Engine
from ..scenario import compile_scenario
from .retriever import Retriever

VIEW_OUTPUTS = ("lane_count", "lane_queued", "lane_meas64", "pos_in", "pos_out", "sig_stats64", "sim", "metrics", "density_map")


class Simulator:
    def __init__(self, parsed_network):
        self.parsed_network = parsed_network
        self.config = parsed_network.config
        self.engine = None
        self.step_measurements = None

    # ---- reference properties (simulator.py:19-50) -----------------------------------
    @property
    def is_terminated(self):
        return self.sim_step == self.config.simulator["sim_length"]

    @property
    def sim_step(self):
        return self.sim_time - self.config.simulator["initial_wait_time"]

    @property
    def sim_time(self):
        return float(self._tick) * float(self.config.simulator["interval"])

    # ---- lifecycle ----------------------------------------------------------------------
    def start_simulator(self):
        """simulator.py:64-78: pick the flow file, create the engine, run the initial
        wait, take the first measurements."""
        gpu = self.config.gpu
        self.scenario = compile_scenario(self.config, self.parsed_network)
        self.n_replicas = int(gpu.get("n_replicas", 1))
        self.view_replica = int(gpu.get("view_replica", 0))
        cap = int(gpu.get("vehicle_capacity", 0)) or 1024
        self.engine = Engine(self.scenario, self.n_replicas, int(gpu.get("device", 0)), vehicle_capacity=cap)
        self._bufs = self.engine.alloc_outputs(VIEW_OUTPUTS)
        self._tick = 0
        self._program_started = False
        self._pending = None
        self._signal_index = {t: i for i, t in enumerate(self.scenario.signal_ids)}
        self._raw0 = np.asarray(self.scenario.sig_phase_raw).reshape(self.engine.A, -1)
        self.retriever = Retriever(self)
        self.cityflow_retriever = self.retriever          # the name the reference uses (simulator.py:75)
        # cityflow.save_replay (backends/cityflow/config.py:88-99): CityFlow-format replay of the view replica
        self._replay = None
        self._raw_now = self._raw0[:, 0].copy()
        if self.config.simulator.get("save_replay", False):
            import os
            from ..replay import ReplayWriter
            out_dir = gpu.get("replay_dir") or self.config.dir
            path = lambda name: os.path.join(out_dir, name)
            for name in (self.config.simulator.get("replay_log_file", "replay_log_file.txt"),
                         self.config.simulator.get("roadnet_log_file", "roadnet_log_file.json")):
                os.makedirs(os.path.dirname(path(name)) or ".", exist_ok=True)
            self._replay = ReplayWriter(self.scenario, self.parsed_network.net,
                                        path(self.config.simulator.get("replay_log_file", "replay_log_file.txt")),
                                        path(self.config.simulator.get("roadnet_log_file", "roadnet_log_file.json")))
        wait = int(self.config.simulator["initial_wait_time"])
        if wait:
            self.engine.step(wait)
            self._tick += wait
        self.retrieve_step_measurements()

    def close_simulator(self):
        """simulator.py:91-95 (engine.reset())."""
        if getattr(self, "_replay", None) is not None:
            self._replay.close()
            self._replay = None
        if self.engine is not None:
            self.engine.close()
            self.engine = None

    # ---- signals ------------------------------------------------------------------------------
    def init_signal_program(self, ts_id, phase_index):
        """TSProgram._initialize_traffic_light_program (backends/cityflow/traffic_signal.py:26-32):
        every signal starts on its phase `phase_index`; one launch serves all of them."""
        if not self._program_started:
            self.engine.init_program(int(phase_index))
            self._program_started = True

    def queue_phase(self, ts_id, phase_index):
        """TSController.switch_phase (traffic_signal.py:51-59): recorded here, applied by the
        next ``simulator_step`` inside the fused env-step launch."""
        if self._pending is None:
            self._pending = np.full(self.engine.A, -1, np.int32)
        self._pending[self._signal_index[ts_id]] = int(phase_index)

    # ---- stepping -------------------------------------------------------------------------------
    def simulator_step(self, n_steps=None):
        """simulator.py:80-89."""
        if n_steps is None:
            n_steps = self.config.simulator["delta_time"]
        if not n_steps:
            return
        if self._pending is not None:
            if (self._pending < 0).any():
                raise RuntimeError("gpu backend: every signal must be given a phase before simulator_step")
            torch = self.engine.torch
            act = torch.from_numpy(np.repeat(self._pending[None], self.n_replicas, 0)).to(self.engine.device)
            self._raw_now = self._raw0[np.arange(self.engine.A), self._pending].copy()
            if self._replay is None:
                self.engine.env_step(act, self._bufs, n_ticks=int(n_steps), controller=2)
            else:                         # one tick per launch so that every tick can be logged (same trajectory)
                self.engine.env_step(act, None, n_ticks=1, controller=2)
                self._log_replay()
                for _ in range(int(n_steps) - 1):
                    self.engine.step(1)
                    self._log_replay()
                self.engine.retrieve(self._bufs)
            self._pending = None
        elif self._replay is not None:
            for _ in range(int(n_steps)):
                self.engine.step(1)
                self._log_replay()
            self.engine.retrieve(self._bufs)
        else:
            self.engine.step(int(n_steps))
            self.engine.retrieve(self._bufs)
        self._tick += int(n_steps)
        self._publish()

    def _log_replay(self):
        self._replay.log_step(self.engine.snapshot(self.view_replica), self._raw_now)

    def retrieve_step_measurements(self):
        """simulator.py:52-62."""
        self.engine.retrieve(self._bufs)
        self._publish()

    def _publish(self):
        self.engine.check()
        b = self.view_replica
        self.view = {k: v[b].cpu().numpy() for k, v in self._bufs.items()}
        self.step_measurements = {
            "lane": self.retriever.retrieve_lane_measurements(),
            "sim": self.retriever.retrieve_sim_measurements(),
        }
