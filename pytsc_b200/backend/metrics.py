"""MetricsParser plugin of the `gpu` backend
(``pytsc/backends/cityflow/metrics.py:7-260``): the network-wide reductions are
one row of the device's ``metrics`` output."""
from __future__ import annotations

import numpy as np


class MetricsParser:
    def __init__(self, parsed_network, simulator, traffic_signals):
        self.config = parsed_network.config
        self.simulator = simulator
        self.parsed_network = parsed_network
        self.traffic_signals = traffic_signals

    def _m(self, k):
        return self.simulator.view["metrics"][k]

    @property
    def flickering_signal(self):
        """metrics.py:24-38 -- from the Python-side program state, so it is defined
        exactly when the reference's is."""
        return np.mean([ts.controller.program.phase_changed for ts in self.traffic_signals.values()])

    n_queued = property(lambda s: int(s._m(0)))                 # metrics.py:41-52
    mean_speed = property(lambda s: float(s._m(1)))             # :70-86
    mean_delay = property(lambda s: float(s._m(2)))             # :132-135
    density = property(lambda s: float(s._m(3)))                # :89-100
    norm_mean_speed = property(lambda s: float(s._m(7)))        # :113-129
    network_flow = property(lambda s: float(s._m(5)))           # :212-219
    average_travel_time = property(lambda s: s.simulator.step_measurements["sim"]["average_travel_time"])
    time_step = property(lambda s: s.simulator.step_measurements["sim"]["time_step"])

    @property
    def n_queued_norm(self):
        """metrics.py:55-67."""
        lanes = self.simulator.step_measurements["lane"]
        return sum(d["n_queued"] / self.parsed_network.lane_lengths[l] for l, d in lanes.items()) / len(lanes)

    @property
    def pressure(self):
        """metrics.py:148-157."""
        return np.sum([ts.pressure for ts in self.traffic_signals.values()]).item()

    @property
    def pressures(self):
        return [ts.pressure for ts in self.traffic_signals.values()]

    @property
    def density_map(self):
        """metrics.py:170-199 -- computed by the retrieve kernel (``density_map`` output) for every replica; this is
        the view replica's matrix."""
        return np.array(self.simulator.view["density_map"], dtype=np.float64)

    @property
    def mst(self):
        """metrics.py:202-209 / common/utils.py compute_max_spanning_tree -- ``tsc_max_spanning_tree`` on the device
        (Prim per replica).  Same tree weight as scipy's; among equally heavy edges the choice may differ."""
        sim = self.simulator
        return sim.engine.max_spanning_tree(sim._bufs["density_map"])[sim.view_replica].cpu().numpy()

    def get_step_stats(self):
        """metrics.py:221-260."""
        stats = {"time_step": self.time_step, "average_travel_time": self.average_travel_time,
                 "n_queued": self.n_queued, "mean_speed": self.mean_speed, "mean_delay": self.mean_delay,
                 "density": self.density, "pressure": self.pressure, "network_flow": self.network_flow}
        if self.config.misc["return_agent_stats"]:
            for ts in self.traffic_signals.values():
                stats.update({f"{ts.id}__phase": ts.controller.current_phase, f"{ts.id}__n_queued": ts.n_queued,
                              f"{ts.id}__mean_speed": ts.mean_speed, f"{ts.id}__mean_delay": ts.mean_delay,
                              f"{ts.id}__density": ts.occupancy, f"{ts.id}__pressure": ts.pressure})
        if self.config.misc["return_lane_stats"]:
            for lane, d in self.simulator.step_measurements["lane"].items():
                for k in ("n_vehicles", "n_queued", "mean_speed", "occupancy"):
                    stats[f"{lane}__{k}"] = d[k].item() if hasattr(d[k], "item") else d[k]
        return stats
