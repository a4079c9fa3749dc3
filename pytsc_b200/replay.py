"""CityFlow-format replay logs for a replica of the gpu backend.

The reference's configs carry ``cityflow.save_replay`` / ``replay_log_file`` / ``roadnet_log_file``
(``pytsc/backends/cityflow/config.py:88-99``) straight into CityFlow's engine config; with ``saveReplay`` on, CityFlow
writes two files its web frontend plays back:

* ``roadnetLogFile`` -- JSON ``{"static": {"nodes": [...], "edges": [...]}}``: every intersection (id, point, virtual flag,
  width) and every road (id, end points, polyline, lane count and widths);
* ``replayLogFile`` -- one text line per simulation step: ``x y direction id laneChange length width,`` for every running
  vehicle, then ``;``, then ``roadId s0 s1 ...,`` for every road that ends at a real intersection, ``s`` = ``g`` / ``r`` per
  lane (all of the lane's lane-links open in the current light phase or not; ``i`` for roads ending at an intersection
  without signal).

This module writes the same two files from the device engine's state (``Engine.snapshot`` of one replica after every
tick) -- a debugging / visualisation aid, host-side, off the step path (a snapshot synchronises the device).  The
format is restated from CityFlow's published frontend format; CityFlow itself is not installable here, so the files
are not byte-compared with real CityFlow output ("parity unpinned" for the engine side).

Geometry: a lane's centre line is its road's polyline shifted to the right of the driving direction by the widths of
the lanes further in plus half its own; a vehicle on a lane-link follows the link's own polyline.
"""
from __future__ import annotations

import json
import math


def _polyline_point(points, dist):
    """Point and heading at arc length ``dist`` along a polyline of (x, y) tuples."""
    rest = max(dist, 0.0)
    for (x0, y0), (x1, y1) in zip(points[:-1], points[1:]):
        seg = math.hypot(x1 - x0, y1 - y0)
        if rest <= seg or (x1, y1) == points[-1]:
            t = 0.0 if seg == 0 else min(rest / seg, 1.0)
            return x0 + t * (x1 - x0), y0 + t * (y1 - y0), math.atan2(y1 - y0, x1 - x0)
        rest -= seg
    x, y = points[-1]
    return x, y, 0.0


def _offset_polyline(points, offset):
    """The polyline shifted ``offset`` metres to the right of its direction (segment normals; joints averaged)."""
    out = []
    n = len(points)
    for k, (x, y) in enumerate(points):
        nx = ny = 0.0
        for a, b in ((k - 1, k), (k, k + 1)):
            if 0 <= a and b < n:
                dx, dy = points[b][0] - points[a][0], points[b][1] - points[a][1]
                l = math.hypot(dx, dy) or 1.0
                nx += dy / l
                ny += -dx / l
        l = math.hypot(nx, ny) or 1.0
        out.append((x + offset * nx / l, y + offset * ny / l))
    return out


def roadnet_log(net: dict) -> dict:
    """The ``roadnetLogFile`` content for a CityFlow roadnet dict."""
    nodes = [{"id": it["id"], "point": [it["point"]["x"], it["point"]["y"]], "virtual": bool(it.get("virtual", False)),
              **({"width": it["width"]} if not it.get("virtual", False) else {})} for it in net["intersections"]]
    edges = [{"id": r["id"], "from": r["startIntersection"], "to": r["endIntersection"],
              "points": [[p["x"], p["y"]] for p in r["points"]], "nLane": len(r["lanes"]),
              "laneWidths": [l["width"] for l in r["lanes"]]} for r in net["roads"]]
    return {"static": {"nodes": nodes, "edges": edges}}


class ReplayWriter:
    """Writes ``roadnet_log_file`` once and appends one ``replay_log_file`` line per ``log_step`` call.

    ``scenario``: the CompiledScenario the engine runs (drivable order, lane-link tables); ``net``: the roadnet dict it
    was compiled from (``parser.net``)."""

    def __init__(self, scenario, net, replay_log_file, roadnet_log_file=None):
        self.cs = scenario
        self.L = scenario.n_lanes
        road_of = {r["id"]: r for r in net["roads"]}
        inter_of = {it["id"]: it for it in net["intersections"]}
        # lane centre lines, in the engine's lane order (road order, lane order)
        self.lane_poly, self.lane_road = [], []
        for r in net["roads"]:
            pts = [(p["x"], p["y"]) for p in r["points"]]
            inner = 0.0
            for k, lane in enumerate(r["lanes"]):
                self.lane_poly.append(_offset_polyline(pts, inner + lane["width"] / 2))
                self.lane_road.append(r["id"])
                inner += lane["width"]
        # CityFlow trims a road at both (real) intersections by their width: distance 0 of a lane is that far in
        self.lane_trim = []
        for r in net["roads"]:
            a = inter_of[r["startIntersection"]]
            self.lane_trim += [0.0 if a.get("virtual", False) else float(a.get("width", 0.0))] * len(r["lanes"])
        # lane-link polylines, in the engine's lane-link order (intersection, road-link, lane-link)
        self.link_poly = []
        for it in net["intersections"]:
            for rl in it.get("roadLinks", []):
                for ll in rl["laneLinks"]:
                    self.link_poly.append([(p["x"], p["y"]) for p in ll["points"]])
        # light status: roads ending at a real intersection, their lanes' lane-links
        self.status_roads = [r for r in net["roads"] if not inter_of[r["endIntersection"]].get("virtual", False)]
        lane_index, k = {}, 0
        for r in net["roads"]:
            for i in range(len(r["lanes"])):
                lane_index[(r["id"], i)] = k
                k += 1
        self.lane_index = lane_index
        off, ll = scenario.lane_ll_off, scenario.lane_ll
        self.lane_links = [[int(x) for x in ll[off[l]:off[l + 1]]] for l in range(self.L)]
        self.names = scenario.vehicle_names
        tm = scenario.tmpl.reshape(-1, 12)
        self.veh_len = [float(tm[t, 0]) for t in scenario.veh_tmpl[: max(scenario.n_vehicles, 1)]]
        self.veh_width = [float(tm[t, 11]) for t in scenario.veh_tmpl[: max(scenario.n_vehicles, 1)]]
        self.f = open(replay_log_file, "w")
        if roadnet_log_file:
            with open(roadnet_log_file, "w") as g:
                json.dump(roadnet_log(net), g)

    def _fmt(self, v):
        return f"{v:.6f}".rstrip("0").rstrip(".") if v == v else "0"

    def line(self, snapshot, raw_phase):
        """One replay line from ``Engine.snapshot(b)`` and the replica's raw light phases (int per signal)."""
        cs, L = self.cs, self.L
        parts = []
        for uid, d, dist in zip(snapshot["uid"], snapshot["drivable"], snapshot["distance"]):
            d = int(d)
            if d < L:
                x, y, ang = _polyline_point(self.lane_poly[d], float(dist) + self.lane_trim[d])
            else:
                x, y, ang = _polyline_point(self.link_poly[d - L], float(dist))
            uid = int(uid)
            parts.append(f"{self._fmt(x)} {self._fmt(y)} {self._fmt(ang)} {self.names[uid]} 0 "
                         f"{self._fmt(self.veh_len[uid])} {self._fmt(self.veh_width[uid])},")
        out = "".join(parts) + ";"
        mask, max_raw = cs.sig_phase_mask, cs.max_raw_phases
        for r in self.status_roads:
            out += r["id"]
            for i in range(len(r["lanes"])):
                links = self.lane_links[self.lane_index[(r["id"], i)]]
                go = True
                for k in links:
                    sg, bit = int(cs.ll_signal[k]), int(cs.ll_roadlink[k])
                    if not (int(mask[sg * max_raw + int(raw_phase[sg])]) >> bit) & 1:
                        go = False
                        break
                out += " g" if go else " r"
            out += ","
        return out

    def log_step(self, snapshot, raw_phase):
        self.f.write(self.line(snapshot, raw_phase) + "\n")

    def close(self):
        self.f.close()
