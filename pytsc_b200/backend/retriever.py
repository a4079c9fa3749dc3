"""Retriever plugin of the `gpu` backend.

The reference's Retriever (``pytsc/backends/cityflow/retriever.py:5-112``)
pulls three dictionaries out of CityFlow and loops over lanes and vehicles in
Python.  Here the per-lane reductions already happened on the device
(``tsc_retrieve``); this class only re-labels the arrays of the viewed replica
into the dictionary layout pytsc's common modules read.
"""
from __future__ import annotations

import math


class PositionMatrixView(list):
    """Per-lane position matrix as the reference exposes it
    (``retriever.py:20-52``), materialised from the two windows the device
    computes: the first and the last ``visibility`` bins -- the only slices any
    pytsc consumer takes (``traffic_signal.py:124,135``; ``controllers.py:109,
    169-175,240``).  Bins outside both windows are NaN on purpose."""


def _position_matrix(n_bins, vis, head, tail):
    n = max(n_bins, vis)
    mat = PositionMatrixView([math.nan] * n)
    if head is not None:
        mat[:vis] = [float(x) for x in head]
    if tail is not None:
        mat[n - vis:] = [float(x) for x in tail]
    if head is None and tail is None:
        mat[:] = [-1.0] * n
    return mat


class Retriever:
    def __init__(self, simulator):
        self.simulator = simulator
        self.parsed_network = simulator.parsed_network
        self.config = simulator.config
        self.visibility = self.config.signal["visibility"]
        self.v_size = self.config.simulator["veh_size_min_gap"]
        self.lane_lengths = self.parsed_network.lane_lengths
        self.lane_max_speeds = self.parsed_network.lane_max_speeds
        cs = simulator.scenario
        self._lane_ids = cs.lane_ids
        # sorted lane ids: the key order of CityFlow's dictionaries
        self._order = sorted(range(len(cs.lane_ids)), key=lambda i: cs.lane_ids[i])
        self._in_row = {int(l): e for e, l in enumerate(cs.sig_in_lane[: cs.n_in_total])}
        self._out_row = {int(l): e for e, l in enumerate(cs.sig_out_lane[: cs.n_out_total])}
        self._bins = [int(self.lane_lengths[l] / self.v_size) for l in cs.lane_ids]

    def retrieve_lane_measurements(self):
        """retriever.py:54-99."""
        v = self.simulator.view
        vis = self.visibility
        out = {}
        for i in self._order:
            head = v["pos_out"][self._out_row[i]] if i in self._out_row else None
            tail = v["pos_in"][self._in_row[i]] if i in self._in_row else None
            out[self._lane_ids[i]] = {
                "n_vehicles": int(v["lane_count"][i]),
                "n_queued": int(v["lane_queued"][i]),
                "occupancy": v["lane_meas64"][i, 0],            # np.float64, as in the reference
                "mean_speed": float(v["lane_meas64"][i, 1]),
                "position_matrix": _position_matrix(self._bins[i], vis, head, tail),
            }
        return out

    def retrieve_sim_measurements(self):
        """retriever.py:101-112."""
        s = self.simulator.view["sim"]
        return {"n_vehicles": int(s[0]), "average_travel_time": float(s[1]), "time_step": float(s[2])}

    def retrieve_ts_measurements(self):
        pass
