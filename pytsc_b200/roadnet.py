"""CityFlow road-network geometry and routing, restated for the scenario compiler.

The `gpu` backend replaces `cityflow.Engine` (reference call site
``pytsc/backends/cityflow/simulator.py:71-74``), so it must derive from the same
roadnet / flow JSON the quantities the CityFlow engine derives internally
(SURVEY.md Appendix A.1, A.3, A.5):

* lane length = road polyline trimmed by the width of each non-virtual end
  intersection; lane-link length = length of its JSON polyline;
* crosses = first polyline intersection of every pair of lane-links of an
  intersection, with the distance along both links, sorted per link;
* per-route drivable sequence (lane, lane-link, lane, ...) from the Router's
  "closest end-lane index" rule;
* the spawn list: one entry per vehicle a Flow creates (tick, route, template,
  priority drawn from mt19937(seed)).

Everything here runs once per scenario on the host; the arithmetic is plain
IEEE fp64 in a fixed operation order (the CPU oracle computes the same numbers
independently in C++ and tests compare them bit for bit).
"""
from __future__ import annotations

import gzip
import json
import math
import os
from dataclasses import dataclass, field

import numpy as np

EPS = 1e-8
ROADLINK_TYPE = {"go_straight": 3, "turn_left": 2, "turn_right": 1}

# CityFlow VehicleInfo defaults (used when a flow's "vehicle" omits a key)
VEHICLE_DEFAULTS = dict(length=5.0, width=2.0, maxPosAcc=4.5, maxNegAcc=4.5, usualPosAcc=2.5,
                        usualNegAcc=2.5, minGap=2.0, maxSpeed=16.66667, headwayTime=1.0,
                        yieldDistance=5.0, turnSpeed=8.3333)
VEHICLE_KEYS = tuple(VEHICLE_DEFAULTS)


def load_json(path):
    if str(path).endswith(".gz"):
        with gzip.open(path, "rt") as f:
            return json.load(f)
    with open(path, "r") as f:
        return json.load(f)


# ---- small fp64 vector helpers (operation order matters) ----------------------
def _len(x, y):
    return math.sqrt(x * x + y * y)


def _unit(x, y):
    l = _len(x, y)
    return x / l, y / l


def _sgn(x):
    return (1 if x + EPS > 0 else 0) - (1 if x < EPS else 0)


def _cross(ax, ay, bx, by):
    return ax * by - ay * bx


def _dot(ax, ay, bx, by):
    return ax * bx + ay * by


def _polyline_length(pts):
    l = 0.0
    for i in range(len(pts) - 1):
        l += _len(pts[i + 1][0] - pts[i][0], pts[i + 1][1] - pts[i][1])
    return l


def _on_segment(A, B, P):
    v1 = _cross(B[0] - A[0], B[1] - A[1], P[0] - A[0], P[1] - A[1])
    v2 = _dot(P[0] - A[0], P[1] - A[1], P[0] - B[0], P[1] - B[1])
    return _sgn(v1) == 0 and _sgn(v2) <= 0


def _intersect_point(A, B, C, D):
    t = _cross(A[0] - C[0], A[1] - C[1], C[0] - D[0], C[1] - D[1]) / \
        _cross(A[0] - B[0], A[1] - B[1], C[0] - D[0], C[1] - D[1])
    return A[0] + (B[0] - A[0]) * t, A[1] + (B[1] - A[1]) * t


@dataclass
class Lane:
    index: int            # drivable index 0..L-1
    road: int
    lane_index: int
    id: str
    width: float
    max_speed: float
    length: float = 0.0
    lanelinks: list = field(default_factory=list)   # lane-link indices (0..K-1) leaving this lane


@dataclass
class LaneLink:
    index: int            # 0..K-1 (drivable index = L + index)
    inter: int            # intersection index (all intersections, JSON order)
    roadlink: int         # road-link index inside the intersection
    rl_type: int          # 3 straight, 2 left, 1 right
    start_lane: int
    end_lane: int
    points: list
    length: float
    crosses: list = field(default_factory=list)     # [(my_dist, foe_ll, foe_dist, cross_id)] sorted by my_dist


class RoadNet:
    """Parsed CityFlow roadnet with engine-side derived geometry."""

    def __init__(self, net: dict):
        self.net = net
        J_int, J_roads = net["intersections"], net["roads"]
        self.road_ids = [r["id"] for r in J_roads]
        self.road_index = {rid: i for i, rid in enumerate(self.road_ids)}
        self.inter_ids = [it["id"] for it in J_int]
        self.inter_index = {iid: i for i, iid in enumerate(self.inter_ids)}
        self.inter_virtual = [bool(it.get("virtual", False)) for it in J_int]
        self.inter_width = [float(it.get("width", 0)) for it in J_int]

        # lanes: road order, lane order
        self.lanes: list[Lane] = []
        self.road_lanes: list[list[int]] = []
        self.road_start = []
        self.road_end = []
        for ri, r in enumerate(J_roads):
            ids = []
            for li, jl in enumerate(r["lanes"]):
                ids.append(len(self.lanes))
                self.lanes.append(Lane(len(self.lanes), ri, li, f"{r['id']}_{li}",
                                       float(jl["width"]), float(jl["maxSpeed"])))
            self.road_lanes.append(ids)
            self.road_start.append(self.inter_index[r["startIntersection"]])
            self.road_end.append(self.inter_index[r["endIntersection"]])
        self._init_lane_lengths(J_roads)

        # lane-links: intersection order, road-link order, lane-link order
        self.lanelinks: list[LaneLink] = []
        self.inter_lanelinks: list[list[int]] = [[] for _ in J_int]
        self.inter_n_roadlinks = [0] * len(J_int)
        self.inter_phases: list[list[tuple[float, list[int]]]] = [[] for _ in J_int]
        for ii, it in enumerate(J_int):
            if self.inter_virtual[ii]:
                continue
            for rli, rl in enumerate(it["roadLinks"]):
                sr, er = self.road_index[rl["startRoad"]], self.road_index[rl["endRoad"]]
                for jll in rl["laneLinks"]:
                    pts = [(float(p["x"]), float(p["y"])) for p in jll.get("points", [])]
                    if len(pts) < 2:
                        raise ValueError("lane-links without explicit points are not supported")
                    ll = LaneLink(len(self.lanelinks), ii, rli, ROADLINK_TYPE[rl["type"]],
                                  self.road_lanes[sr][int(jll["startLaneIndex"])],
                                  self.road_lanes[er][int(jll["endLaneIndex"])],
                                  pts, _polyline_length(pts))
                    self.lanes[ll.start_lane].lanelinks.append(ll.index)
                    self.inter_lanelinks[ii].append(ll.index)
                    self.lanelinks.append(ll)
            self.inter_n_roadlinks[ii] = len(it["roadLinks"])
            for ph in it["trafficLight"]["lightphases"]:
                self.inter_phases[ii].append((float(ph["time"]), [int(a) for a in ph["availableRoadLinks"]]))
        self.n_crosses = 0
        self.inter_crosses: list[list[tuple[int, int, float, float]]] = [[] for _ in J_int]
        for ii in range(len(J_int)):
            self._init_crosses(ii)

    # -- A.1 --------------------------------------------------------------------
    def _init_lane_lengths(self, J_roads):
        for ri, r in enumerate(J_roads):
            rp = [(float(p["x"]), float(p["y"])) for p in r["points"]]
            si, ei = self.road_start[ri], self.road_end[ri]
            if not self.inter_virtual[si]:
                w = self.inter_width[si]
                ux, uy = _unit(rp[1][0] - rp[0][0], rp[1][1] - rp[0][1])
                rp[0] = (rp[0][0] + ux * w, rp[0][1] + uy * w)
            if not self.inter_virtual[ei]:
                w = self.inter_width[ei]
                ux, uy = _unit(rp[-1][0] - rp[-2][0], rp[-1][1] - rp[-2][1])
                rp[-1] = (rp[-1][0] - ux * w, rp[-1][1] - uy * w)
            dsum = 0.0
            n = len(rp)
            for li in self.road_lanes[ri]:
                lane = self.lanes[li]
                dmin, dmax = dsum, dsum + lane.width
                off = (dmin + dmax) / 2.0
                lpts = []
                for j in range(n):
                    if j == 0:
                        ux, uy = _unit(rp[1][0] - rp[0][0], rp[1][1] - rp[0][1])
                    elif j + 1 == n:
                        ux, uy = _unit(rp[j][0] - rp[j - 1][0], rp[j][1] - rp[j - 1][1])
                    else:
                        u1 = _unit(rp[j + 1][0] - rp[j][0], rp[j + 1][1] - rp[j][1])
                        u2 = _unit(rp[j][0] - rp[j - 1][0], rp[j][1] - rp[j - 1][1])
                        ux, uy = _unit(u1[0] + u2[0], u1[1] + u2[1])
                    # v = -normal(u) = (uy, -ux)
                    vx, vy = -uy * -1.0, ux * -1.0
                    lpts.append((rp[j][0] + vx * off, rp[j][1] + vy * off))
                lane.length = _polyline_length(lpts)
                dsum += lane.width

    # -- A.5 --------------------------------------------------------------------
    def _init_crosses(self, ii):
        lls = [self.lanelinks[k] for k in self.inter_lanelinks[ii]]
        out = self.inter_crosses[ii]
        n = len(lls)
        for i in range(n):
            la = lls[i]
            pa = la.points
            for j in range(i + 1, n):
                lb = lls[j]
                pb = lb.points
                disa = 0.0
                found = False
                for ia in range(len(pa) - 1):
                    A1, A2 = pa[ia], pa[ia + 1]
                    disb = 0.0
                    for ib in range(len(pb) - 1):
                        B1, B2 = pb[ib], pb[ib + 1]
                        seg_b = _len(B2[0] - B1[0], B2[1] - B1[1])
                        if _sgn(_cross(A2[0] - A1[0], A2[1] - A1[1], B2[0] - B1[0], B2[1] - B1[1])) == 0:
                            disb += seg_b
                            continue
                        P = _intersect_point(A1, A2, B1, B2)
                        if _on_segment(A1, A2, P) and _on_segment(B1, B2, P):
                            da = disa + _len(P[0] - A1[0], P[1] - A1[1])
                            db = disb + _len(P[0] - B1[0], P[1] - B1[1])
                            out.append((la.index, lb.index, da, db))
                            found = True
                            break
                        disb += seg_b
                    if found:
                        break
                    disa += _len(A2[0] - A1[0], A2[1] - A1[1])
        for cid, (a, b, da, db) in enumerate(out):
            gid = self.n_crosses + cid
            self.lanelinks[a].crosses.append((da, b, db, gid))
            self.lanelinks[b].crosses.append((db, a, da, gid))
        self.n_crosses += len(out)
        for ll in lls:
            ll.crosses.sort(key=lambda c: c[0])   # stable, like the oracle

    # -- A.3 routing --------------------------------------------------------------
    def lanelinks_to_road(self, lane: int, road: int):
        return [k for k in self.lanes[lane].lanelinks if self.lanes[self.lanelinks[k].end_lane].road == road]

    def route_valid(self, route):
        if not route:
            return False
        for a, b in zip(route[:-1], route[1:]):
            if not any(self.lanelinks_to_road(l, b) for l in self.road_lanes[a]):
                return False
        return True

    def first_lane_candidates(self, route):
        lanes = self.road_lanes[route[0]]
        if len(route) == 1:
            return list(lanes)
        return [l for l in lanes if self.lanelinks_to_road(l, route[1])]

    def _select_lanelink(self, cur_lane, cands):
        sel, best = None, 0
        ci = self.lanes[cur_lane].lane_index
        for k in cands:
            d = abs(self.lanes[self.lanelinks[k].end_lane].lane_index - ci)
            if sel is None or d < best:
                sel, best = k, d
        return sel

    def drivable_sequence(self, route, first_lane):
        """Drivable indices (lanes 0..L-1, lane-links L+k) a vehicle on `route`
        visits when it starts on `first_lane`."""
        L = len(self.lanes)
        seq = [first_lane]
        lane = first_lane
        for t in range(len(route) - 1):
            lls = self.lanelinks_to_road(lane, route[t + 1])
            if t + 2 < len(route):
                lls = [k for k in lls if self.lanelinks_to_road(self.lanelinks[k].end_lane, route[t + 2])]
            k = self._select_lanelink(lane, lls)
            if k is None:
                raise ValueError("route cannot be driven from lane %s" % self.lanes[lane].id)
            seq.append(L + k)
            lane = self.lanelinks[k].end_lane
            seq.append(lane)
        return seq


class MT19937:
    """std::mt19937 (32-bit Mersenne twister, init_genrand seeding)."""

    def __init__(self, seed):
        self.mt = [0] * 624
        self.mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            self.mt[i] = (1812433253 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.idx = 624

    def __call__(self):
        if self.idx >= 624:
            mt = self.mt
            for i in range(624):
                y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.idx = 0
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF


def expand_flows(rn: RoadNet, flows: list, interval: float, seed: int, horizon_ticks: int):
    """Run Flow::nextStep / planRoute symbolically for `horizon_ticks` ticks and
    return the spawn list in creation order (A.3, A.6).

    Returns dict with per-vehicle arrays (tick, flow, flow_cnt, route_key,
    template index, priority, first_lane) plus `routes` (unique road tuples) and
    `templates` (unique vehicle parameter tuples)."""
    rnd = MT19937(seed)
    templates, tmpl_index = [], {}
    routes, route_index = [], {}
    fl = []
    for fi, f in enumerate(flows):
        v = f.get("vehicle", {})
        t = tuple(float(v.get(k, VEHICLE_DEFAULTS[k])) for k in VEHICLE_KEYS)
        if t not in tmpl_index:
            tmpl_index[t] = len(templates)
            templates.append(t)
        route = tuple(rn.road_index[r] for r in f["route"])
        fl.append(dict(tmpl=tmpl_index[t], route=route, interval=float(f.get("interval", 1.0)),
                       start=float(f.get("startTime", 0)), end=float(f.get("endTime", -1)),
                       now=float(f.get("interval", 1.0)), cur=0.0, cnt=0, valid=True,
                       route_ok=rn.route_valid(route)))
    # event-driven: most flows are single-shot; keep only flows still able to fire
    out = dict(tick=[], flow=[], flow_cnt=[], route=[], tmpl=[], priority=[], first_lane=[])
    live = list(range(len(fl)))
    seen_prio = set()
    dup_prio = 0
    for tick in range(horizon_ticks):
        created = []
        nxt = []
        for fi in live:
            f = fl[fi]
            if not f["valid"]:
                continue
            if f["end"] != -1 and f["cur"] > f["end"]:
                continue
            if f["cur"] >= f["start"]:
                while f["now"] >= f["interval"]:
                    pr = rnd()
                    if pr in seen_prio:
                        dup_prio += 1
                    seen_prio.add(pr)
                    pr = pr - (1 << 32) if pr >= (1 << 31) else pr   # (int) cast
                    created.append((fi, f["cnt"], pr))
                    f["cnt"] += 1
                    f["now"] -= f["interval"]
                f["now"] += interval
            f["cur"] += interval
            nxt.append(fi)
        live = nxt
        # planRoute: roads in roadnet order, each road's buffer in creation order
        first_lane = {}
        for fi, cnt, pr in sorted(created, key=lambda c: fl[c[0]]["route"][0]):   # stable
            if not fl[fi]["route_ok"]:      # planRoute drops the vehicle and disables the flow
                fl[fi]["valid"] = False
                continue
            cands = rn.first_lane_candidates(fl[fi]["route"])
            first_lane[(fi, cnt)] = cands[rnd() % len(cands)]
        for fi, cnt, pr in created:
            f = fl[fi]
            if (fi, cnt) not in first_lane:
                continue
            key = (f["route"], first_lane[(fi, cnt)])
            if key not in route_index:
                route_index[key] = len(routes)
                routes.append(key)
            out["tick"].append(tick)
            out["flow"].append(fi)
            out["flow_cnt"].append(cnt)
            out["route"].append(route_index[key])
            out["tmpl"].append(f["tmpl"])
            out["priority"].append(pr)
            out["first_lane"].append(first_lane[(fi, cnt)])
    res = {k: np.asarray(v, dtype=np.int64) for k, v in out.items()}
    res["routes"] = routes
    res["templates"] = templates
    res["duplicate_priorities"] = dup_prio
    res["invalid_flows"] = [i for i, f in enumerate(fl) if not f["valid"]]
    return res


def pytsc_lane_geometry(net: dict):
    """Per-lane quantities as *pytsc* defines them (different from CityFlow's):
    length = centre-to-centre distance of the end intersections
    (``backends/cityflow/network_parser.py:325-352``), angle in degrees
    (``:389-408``)."""
    pts = {it["id"]: (it["point"]["x"], it["point"]["y"]) for it in net["intersections"]}
    length, angle = {}, {}
    for r in net["roads"]:
        s, e = pts[r["startIntersection"]], pts[r["endIntersection"]]
        l = float(np.linalg.norm(np.array([s[0], s[1]]) - np.array([e[0], e[1]])))
        a = math.degrees(math.atan2(e[1] - s[1], e[0] - s[0]))
        for i in range(len(r["lanes"])):
            length[f"{r['id']}_{i}"] = l
            angle[f"{r['id']}_{i}"] = a
    return length, angle


def scenario_dir_candidates(scenario: str):
    here = os.path.dirname(os.path.abspath(__file__))
    yield os.path.join(here, "scenarios", scenario)
    env = os.environ.get("PYTSC_B200_SCENARIOS")
    if env:
        yield os.path.join(env, scenario)
