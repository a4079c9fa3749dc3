"""TrafficSignal plugin of the `gpu` backend.

Attribute protocol of the reference's ``TrafficSignal`` / ``TSController`` /
``TSProgram`` (``pytsc/backends/cityflow/traffic_signal.py:10-150``,
``pytsc/common/traffic_signal.py:13-404``) as read by pytsc's action spaces,
observation spaces, reward functions, metrics and rule-based controllers.
Phase changes are queued on the Simulator and applied by the next fused
env-step launch; the per-signal statistics come from the device
(``sig_stats64``) instead of Python loops over lane dictionaries.
"""
from __future__ import annotations

import numpy as np


class TSProgram:
    """common/traffic_signal.py:60-112 + backends/cityflow/traffic_signal.py:10-32."""
    start_phase_index = 0

    def __init__(self, id, config, simulator):
        self.id = id
        self.phases = config["phases"]
        self.phase_indices = config["phase_indices"]
        self.phases_min_max_times = config["phases_min_max_times"]
        self.yellow_time = config["yellow_time"]
        simulator.init_signal_program(id, self.start_phase_index)
        self.set_initial_phase(self.start_phase_index)

    def set_initial_phase(self, phase_index):
        self.current_phase_index = phase_index
        self.current_phase = self.phases[phase_index]
        self.time_on_phase = 0
        self.norm_time_on_phase = 0

    def update_current_phase(self, phase_index):
        if phase_index == self.current_phase_index:
            self.phase_changed = False
            self.time_on_phase += self.yellow_time
        else:
            self.phase_changed = True
            self.time_on_phase = self.yellow_time
        self.current_phase_index = phase_index
        self.current_phase = self.phases[phase_index]
        self.norm_time_on_phase = self.time_on_phase / self.phases_min_max_times[self.current_phase]["max_time"]


class TSController:
    """common/traffic_signal.py:115-296 + backends/cityflow/traffic_signal.py:35-59."""

    def __init__(self, id, config, simulator):
        self.id = id
        self.config = config
        self.simulator = simulator
        self.phases = config["phases"]
        self.program = TSProgram(id, config, simulator)

    n_phases = property(lambda s: s.config["n_phases"])
    phase_indices = property(lambda s: s.config["phase_indices"])
    green_phase_indices = property(lambda s: s.config["green_phase_indices"])
    yellow_phase_indices = property(lambda s: s.config["yellow_phase_indices"])
    yellow_time = property(lambda s: s.config["yellow_time"])
    phases_min_max_times = property(lambda s: s.config["phases_min_max_times"])
    current_phase = property(lambda s: s.program.current_phase)
    current_phase_index = property(lambda s: s.program.current_phase_index)
    next_phase_index = property(lambda s: (s.program.current_phase_index + 1) % s.n_phases)
    next_green_phase_index = property(lambda s: (s.program.current_phase_index + 2) % s.n_phases)
    time_on_phase = property(lambda s: s.program.time_on_phase)
    norm_time_on_phase = property(lambda s: s.program.norm_time_on_phase)

    @property
    def phase_one_hot(self):
        one_hot = [0] * self.n_phases
        one_hot[self.current_phase_index] = 1
        return one_hot

    def get_allowable_phase_switches(self):
        """TLSFreePhaseSelectLogic / TLSRoundRobinPhaseSelectLogic
        (common/traffic_signal.py:329-361, 375-404)."""
        mask = [0] * self.n_phases
        cur, nxt, t = self.current_phase_index, self.next_phase_index, self.time_on_phase
        if cur in self.green_phase_indices:
            mm = self.phases_min_max_times[self.current_phase]
            if t < mm["min_time"]:
                mask[cur] = 1
            elif t < mm["max_time"]:
                mask[cur] = 1
                mask[nxt] = 1
            elif t == mm["max_time"]:
                mask[nxt] = 1
            else:
                raise RuntimeError(f"{self.id}: time_on_phase {t} beyond max_time {mm['max_time']}")
        elif self.config["round_robin"]:
            mask[nxt] = 1
        else:
            for g in self.green_phase_indices:
                if g != cur - 1:
                    mask[g] = 1
        return mask

    def switch_phase(self, phase_index):
        self.simulator.queue_phase(self.id, phase_index)       # engine.set_tl_phase, applied at the next step
        self.program.update_current_phase(phase_index)


class FixedTimeController:
    """controllers/controllers.py:26-54."""

    def __init__(self, traffic_signal, green_time=25):
        self.traffic_signal = traffic_signal
        self.green_time = green_time
        self.controller = traffic_signal.controller

    def get_action(self, inp):
        c = self.controller
        if c.current_phase_index in c.green_phase_indices and c.time_on_phase < self.green_time:
            return c.current_phase_index
        return c.next_phase_index


class TrafficSignal:
    def __init__(self, id, config, simulator):
        self.id = id
        self.config = config
        self.simulator = simulator
        self.n_phases = config["n_phases"]
        self.controller = TSController(id, config, simulator)
        self.incoming_lanes = config["incoming_lanes"]
        self.outgoing_lanes = config["outgoing_lanes"]
        self.sub_results = None
        self._index = simulator.scenario.signal_ids.index(id)
        self.init_rule_based_controllers()

    def __repr__(self):
        return f"TrafficSignal ({self.id})"

    def init_rule_based_controllers(self):
        """common/traffic_signal.py:46-55: pytsc's own rule-based controllers when the
        package is importable, else the fixed-time controller only."""
        try:
            from pytsc.controllers import (FixedTimeController as F, GreedyController, MaxPressureController,
                                           SOTLController)
            self.controllers = {"fixed_time": F(self), "greedy": GreedyController(self),
                                "max_pressure": MaxPressureController(self), "sotl": SOTLController(self)}
        except Exception:
            self.controllers = {"fixed_time": FixedTimeController(self)}

    def get_controller_action(self, controller):
        inp = self.simulator.step_measurements
        inp.update({"time": self.simulator.sim_time,
                    "current_phase_index": self.controller.program.current_phase_index,
                    "time_on_phase": self.controller.time_on_phase})
        return self.controllers[controller].get_action(inp)

    def update_stats(self, sub_results):
        """backends/cityflow/traffic_signal.py:101-141; the sums over incoming /
        outgoing lanes were taken on the device."""
        self.sub_results = sub_results
        vis = self.config["visibility"]
        st = self.simulator.view["sig_stats64"][self._index]
        self.n_queued = int(st[0])
        self.occupancy = st[1]
        self.mean_speed = float(st[2])
        self.mean_delay = float(st[3])
        self.outgoing_occupancy = st[4]
        self.pressure = float(st[5])
        lanes = sub_results["lane"]
        self.inc_position_matrices = {l: lanes[l]["position_matrix"][-vis:] for l in self.incoming_lanes}
        self.out_position_matrices = {l: lanes[l]["position_matrix"][:vis] for l in self.outgoing_lanes}
        self.time_on_phase = self.controller.norm_time_on_phase
        self.phase_id = np.asarray(self.controller.phase_one_hot)
        self.sim_step = self.simulator.sim_step / 3600

    def action_to_phase(self, phase_index):
        self.controller.switch_phase(phase_index)
