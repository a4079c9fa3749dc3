"""Network parser for the `gpu` backend.

Produces the attribute protocol that pytsc's common observation / reward /
action / metric modules read from a backend's parsed network:
``traffic_signals`` (ordered, one config dict per signalised intersection),
``lanes``, ``lane_lengths``, ``lane_max_speeds``, ``lane_indices``,
``lane_angles``, ``adjacency_matrix``, ``k_hop_neighbors``,
``traffic_signal_ids``, ``neighbors_lanes`` ...

Semantics follow ``pytsc/backends/cityflow/network_parser.py`` (line numbers in
the method docstrings); the implementation is eager and array based so that the
scenario compiler can lower it to device tables directly.
"""
from __future__ import annotations

import math

import numpy as np

from ..bundle import load_roadnet


class NetworkParser:
    def __init__(self, config):
        self.config = config
        self.net = load_roadnet(config.cityflow_roadnet_file)
        self.intersections = self.net["intersections"]
        self.roads = self.net["roads"]
        self._inter_by_id = {it["id"]: it for it in self.intersections}
        self._build_lanes()
        self._build_graph()
        self._build_traffic_signals()

    # ---- lanes -------------------------------------------------------------------
    def _build_lanes(self):
        """lanes (:98-113, sorted ids), lane_lengths = centre-to-centre distance of
        the road's end intersections (:325-352), lane_max_speeds (:355-369),
        lane_indices (:372-386), lane_angles in degrees (:389-408)."""
        lane_ids, self.lane_lengths, self.lane_max_speeds = [], {}, {}
        self.lane_indices, self.lane_angles = {}, {}
        for road in self.roads:
            s = self._inter_by_id[road["startIntersection"]]["point"]
            e = self._inter_by_id[road["endIntersection"]]["point"]
            length = np.linalg.norm(np.array([s["x"], s["y"]]) - np.array([e["x"], e["y"]]))
            angle = math.degrees(math.atan2(e["y"] - s["y"], e["x"] - s["x"]))
            for i, lane in enumerate(road["lanes"]):
                lid = f"{road['id']}_{i}"
                lane_ids.append(lid)
                self.lane_lengths[lid] = length
                self.lane_max_speeds[lid] = lane["maxSpeed"]
                self.lane_indices[lid] = i
                self.lane_angles[lid] = angle
        self.lanes = sorted(lane_ids)

    # ---- signal graph --------------------------------------------------------------
    def _build_graph(self):
        """traffic_signal_ids sorted (:116-134); adjacency over roads joining two
        signalised intersections, symmetric (:137-166); k-hop neighbours from
        matrix powers (:169-184, :598-614); neighbours' connecting lanes (:253-292)."""
        self.traffic_signal_ids = sorted(it["id"] for it in self.intersections if not it["virtual"])
        idx = {t: i for i, t in enumerate(self.traffic_signal_ids)}
        n = len(idx)
        if "neighbors" in self.config.network:
            adj = np.zeros((n, n))
            nb = self.config.network["neighbors"]
            for a in self.traffic_signal_ids:
                for b in self.traffic_signal_ids:
                    if a in nb and b in nb[a]:
                        adj[idx[a], idx[b]] = 1.0
        else:
            adj = np.zeros((n, n))
            for road in self.roads:
                a, b = road["startIntersection"], road["endIntersection"]
                if a in idx and b in idx:
                    adj[idx[a], idx[b]] = 1.0
                    adj[idx[b], idx[a]] = 1.0
        self.adjacency_matrix = adj
        max_hops = self.config.misc["max_hops"]
        self.k_hop_neighbors = {}
        for t in self.traffic_signal_ids:
            self.k_hop_neighbors[t] = {}
            for k in range(1, max_hops + 1):
                row = np.linalg.matrix_power(adj, k)[idx[t]]
                self.k_hop_neighbors[t][k] = [self.traffic_signal_ids[j] for j in np.where(row > 0)[0]]
        if "neighbors_lanes" in self.config.network:
            self.neighbors_lanes = self.config.network["neighbors_lanes"]
        else:
            self.neighbors_lanes = {}
            for t in self.traffic_signal_ids:
                self.neighbors_lanes[t] = {}
                for j in np.where(adj[idx[t]] > 0)[0]:
                    nb_id = self.traffic_signal_ids[j]
                    lanes = []
                    for road in self.roads:
                        if road["startIntersection"] == t and road["endIntersection"] == nb_id:
                            lanes += [f"{road['id']}_{i}" for i in range(len(road["lanes"]))]
                    self.neighbors_lanes[t][nb_id] = lanes
        self.in_degrees = adj.sum(axis=0)
        self.out_degrees = adj.sum(axis=1)

    # ---- per-signal configuration -----------------------------------------------------
    def _phase_plan(self, inter):
        """Green = has available road-links and lasts longer than 5 s; the rest is
        yellow.  pytsc phases interleave green, yellow, green, yellow ...
        (:631-703); min/max times from the signal config."""
        sig = self.config.signal
        program = inter["trafficLight"]["lightphases"]
        green, yellow, mm = [], [], {}
        for i, p in enumerate(program):
            if len(p["availableRoadLinks"]) and p["time"] > 5:
                green.append(i)
                mm[i] = {"min_time": sig["min_green_time"], "max_time": sig["max_green_time"]}
            else:
                yellow.append(i)
                mm[i] = {"min_time": sig["yellow_time"], "max_time": sig["yellow_time"]}
        ys = [yellow[0]] * len(green) if len(yellow) == 1 else yellow
        phases = [p for pair in zip(green, ys) for p in pair]
        g_idx = [phases.index(g) for g in green]
        y_idx = [g + 1 for g in g_idx]
        p_idx = [p for pair in zip(g_idx, y_idx) for p in pair]
        return phases, mm, p_idx, g_idx, y_idx

    def _build_traffic_signals(self):
        """One entry per non-virtual intersection **in roadnet JSON order**
        (:32-78): lane maps (:598-629 sorted unique), phase->(incoming lane ->
        outgoing lanes) (:212-250), phase plan, then every ``signal.*`` key."""
        xs = [it["point"]["x"] for it in self.intersections]
        ys = [it["point"]["y"] for it in self.intersections]
        self.network_boundary = ((min(xs), min(ys)), (max(xs), max(ys)))
        self.norm_network_boundary = [max(xs) - min(xs), max(ys) - min(ys)]
        self.ts_coordinates, self.ts_norm_coordinates = {}, {}
        self.ts_phase_to_inc_out_lanes = {}
        self.traffic_signals = {}
        for inter in self.intersections:
            if inter["virtual"]:
                continue
            ts_id = inter["id"]
            inc, out, mapping, by_roadlink = [], [], {}, []
            for rl in inter["roadLinks"]:
                pairs = []
                for ll in rl["laneLinks"]:
                    a = f"{rl['startRoad']}_{ll['startLaneIndex']}"
                    b = f"{rl['endRoad']}_{ll['endLaneIndex']}"
                    inc.append(a)
                    out.append(b)
                    mapping.setdefault(a, []).append(b)
                    pairs.append((a, b))
                by_roadlink.append(pairs)
            p2l = {}
            if "trafficLight" in inter:
                for i, ph in enumerate(inter["trafficLight"]["lightphases"]):
                    p2l[i] = {}
                    for r in ph["availableRoadLinks"]:
                        for a, b in by_roadlink[r]:
                            p2l[i].setdefault(a, []).append(b)
            self.ts_phase_to_inc_out_lanes[ts_id] = p2l
            self.ts_coordinates[ts_id] = [inter["point"]["x"], inter["point"]["y"]]
            self.ts_norm_coordinates[ts_id] = [inter["point"]["x"] / self.norm_network_boundary[0],
                                               inter["point"]["y"] / self.norm_network_boundary[1]]
            phases, mm, p_idx, g_idx, y_idx = self._phase_plan(inter)
            if "phase_sequence" in self.config.simulator:
                phases = self.config.simulator["phase_sequence"]
                p_idx = list(range(len(phases)))
                g_idx, y_idx = p_idx[0::2], p_idx[1::2]
            cfg = {
                "coordinates": self.ts_coordinates[ts_id],
                "norm_coordinates": self.ts_norm_coordinates[ts_id],
                "incoming_lanes": sorted(set(inc)),
                "outgoing_lanes": sorted(set(out)),
                "inc_to_out_lanes": mapping,
                "phase_to_inc_out_lanes": p2l,
                "phases": phases,
                "n_phases": len(phases),
                "phases_min_max_times": mm,
                "phase_indices": p_idx,
                "green_phase_indices": g_idx,
                "yellow_phase_indices": y_idx,
            }
            cfg.update(self.config.signal)
            self.traffic_signals[ts_id] = cfg

    # ---- derived helpers other pytsc modules use ---------------------------------------
    @property
    def neighbors_offsets(self):
        """(:295-321)."""
        out = {t: {} for t in self.traffic_signal_ids}
        dt = self.config.simulator["delta_time"]
        for t in self.traffic_signal_ids:
            nl = self.neighbors_lanes[t]
            for nb in self.traffic_signal_ids:
                if nl and nb in nl:
                    lanes = nl[nb]
                    tt = sum(self.lane_lengths[l] / self.lane_max_speeds[l] for l in lanes)
                    out[t][nb] = int(tt / len(lanes) / dt)
        return out

    @property
    def distance_matrix(self):
        """Hop distances between signals (:484-499) by breadth-first search."""
        n = len(self.traffic_signal_ids)
        dist = np.zeros((n, n))
        for s in range(n):
            seen, frontier, d = {s}, [s], 0
            while frontier:
                d += 1
                nxt = []
                for u in frontier:
                    for v in np.where(self.adjacency_matrix[u] > 0)[0]:
                        if v not in seen:
                            seen.add(int(v))
                            dist[s, v] = d
                            nxt.append(int(v))
                frontier = nxt
        return dist
