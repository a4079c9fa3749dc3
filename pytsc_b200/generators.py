"""Scenario generation for the gpu backend: grid roadnets and Gaussian-arrival flows.

Two pieces the reference obtains elsewhere:

* ``grid_roadnet`` -- the reference shells out to CityFlow's own
  ``tools/generator/generate_grid_scenario.py`` (``pytsc/backends/cityflow/
  grid_generator.py:39-73``), which is not part of the reference checkout.  This
  emitter writes the same roadnet JSON schema; it is pinned against the grids the
  reference ships (``scenarios/cityflow/syn_1x1``, ``syn_3x3``): same ids, same
  ordering, same road-link / lane-link tables, same signal plan, lane-link
  geometry equal to 1e-9 (``tests/test_generators.py``).
* ``GridTripGenerator`` -- ``CityFlowTripGenerator`` (``pytsc/backends/cityflow/
  trip_generator.py:45-286``): per incoming fringe road, Gaussian inter-arrival
  times; per vehicle, a random walk over road links weighted by turn
  probability x edge weight until an outgoing fringe road is reached.  The two
  random streams are consumed in the reference's order, so a given seed yields
  the reference's flow list (pinned by ``tests/golden/trips_syn_3x3.npz``).

Host-side, once per scenario; nothing here runs on the step path.
"""
from __future__ import annotations

import random
from collections import deque

import numpy as np

# direction k of road_{i}_{j}_{k}: 0 east, 1 north, 2 west, 3 south
_DX = (1, 0, -1, 0)
_DY = (0, 1, 0, -1)


def _hermite(p0, d0, p1, d1, scale, n=10):
    """n + 1 points of the cubic Hermite curve from p0 (tangent d0 * scale) to p1 (tangent d1 * scale)."""
    pts = []
    for s in range(n + 1):
        t = s / n
        h00 = 2 * t ** 3 - 3 * t ** 2 + 1
        h10 = t ** 3 - 2 * t ** 2 + t
        h01 = -2 * t ** 3 + 3 * t ** 2
        h11 = t ** 3 - t ** 2
        pts.append({"x": h00 * p0[0] + h10 * scale * d0[0] + h01 * p1[0] + h11 * scale * d1[0],
                    "y": h00 * p0[1] + h10 * scale * d0[1] + h01 * p1[1] + h11 * scale * d1[1]})
    return pts


def grid_roadnet(rows, cols, row_distance=300.0, col_distance=300.0, intersection_width=20.0,
                 n_left_lanes=1, n_straight_lanes=1, n_right_lanes=1, lane_max_speed=11.11, lane_width=4.0,
                 green_time=30, yellow_time=5):
    """A rows x cols signalised grid with a ring of virtual (source / sink) intersections, in CityFlow's
    roadnet schema.  Intersections are ``intersection_{i}_{j}`` (i = column 1..cols, j = row 1..rows,
    0 and cols+1 / rows+1 the virtual ring); roads ``road_{i}_{j}_{k}`` leave intersection (i, j) in
    direction k.  Lanes are numbered from the road's centre line outwards: left-turn lanes first, then
    straight, then right.  Every real intersection gets the 8-phase plan of the shipped grids
    (one all-red-but-right-turns phase followed by eight green phases)."""
    n_lanes = n_left_lanes + n_straight_lanes + n_right_lanes
    lane_of = {"turn_left": list(range(0, n_left_lanes)),
               "go_straight": list(range(n_left_lanes, n_left_lanes + n_straight_lanes)),
               "turn_right": list(range(n_left_lanes + n_straight_lanes, n_lanes))}

    def exists(i, j):
        inside_i, inside_j = 1 <= i <= cols, 1 <= j <= rows
        if inside_i and inside_j:
            return True
        return (inside_i and j in (0, rows + 1)) or (inside_j and i in (0, cols + 1))

    def virtual(i, j):
        return not (1 <= i <= cols and 1 <= j <= rows)

    def point(i, j):
        return ((i - 1) * col_distance, (j - 1) * row_distance)

    def _num(v):    # the shipped files hold ints where the value is integral
        return int(v) if float(v).is_integer() else v

    def has_road(i, j, k):
        ti, tj = i + _DX[k], j + _DY[k]
        return exists(i, j) and exists(ti, tj) and not (virtual(i, j) and virtual(ti, tj))

    roads, inters = [], []
    for j in range(rows + 2):
        for i in range(cols + 2):
            if not exists(i, j):
                continue
            for k in range(4):
                if has_road(i, j, k):
                    ti, tj = i + _DX[k], j + _DY[k]
                    (x0, y0), (x1, y1) = point(i, j), point(ti, tj)
                    roads.append({"id": f"road_{i}_{j}_{k}",
                                  "points": [{"x": _num(x0), "y": _num(y0)}, {"x": _num(x1), "y": _num(y1)}],
                                  "lanes": [{"width": _num(lane_width), "maxSpeed": lane_max_speed} for _ in range(n_lanes)],
                                  "startIntersection": f"intersection_{i}_{j}",
                                  "endIntersection": f"intersection_{ti}_{tj}"})
    green_plan_pairs = (("go_straight", 0, "go_straight", 2), ("go_straight", 1, "go_straight", 3),
                        ("turn_left", 0, "turn_left", 2), ("turn_left", 1, "turn_left", 3),
                        ("go_straight", 0, "turn_left", 0), ("go_straight", 2, "turn_left", 2),
                        ("go_straight", 1, "turn_left", 1), ("go_straight", 3, "turn_left", 3))
    for j in range(rows + 2):
        for i in range(cols + 2):
            if not exists(i, j):
                continue
            cx, cy = point(i, j)
            isv = virtual(i, j)
            incoming = []     # (road id, travel direction) in the order W, S, E, N approaches
            for k in range(4):
                si, sj = i - _DX[k], j - _DY[k]
                if has_road(si, sj, k):
                    incoming.append((f"road_{si}_{sj}_{k}", k))
            outgoing = [(f"road_{i}_{j}_{k}", k) for k in range(4) if has_road(i, j, k)]
            road_links = []
            if not isv:
                w = intersection_width
                for rid, k in incoming:
                    for oid, ok in outgoing:
                        if ok == (k + 2) % 4:
                            continue      # no U-turns
                        kind = "go_straight" if ok == k else ("turn_left" if ok == (k + 1) % 4 else "turn_right")
                        d0, d1 = (_DX[k], _DY[k]), (_DX[ok], _DY[ok])
                        r0, r1 = (_DY[k], -_DX[k]), (_DY[ok], -_DX[ok])     # unit vector to the right of travel
                        lane_links = []
                        for a in lane_of[kind]:
                            for b in range(n_lanes):
                                off_a, off_b = (a + 0.5) * lane_width, (b + 0.5) * lane_width
                                p0 = (cx - d0[0] * w + r0[0] * off_a, cy - d0[1] * w + r0[1] * off_a)
                                p1 = (cx + d1[0] * w + r1[0] * off_b, cy + d1[1] * w + r1[1] * off_b)
                                lane_links.append({"startLaneIndex": a, "endLaneIndex": b,
                                                   "points": _hermite(p0, d0, p1, d1, w)})
                        road_links.append({"type": kind, "startRoad": rid, "endRoad": oid, "direction": k,
                                           "laneLinks": lane_links, "_in": k})
            rights = [n for n, rl in enumerate(road_links) if rl["type"] == "turn_right"]
            phases = [{"time": yellow_time, "availableRoadLinks": _first_phase_order(road_links, rights) if not isv else []}]
            for ta, ka, tb, kb in green_plan_pairs:
                if isv:
                    phases.append({"time": green_time, "availableRoadLinks": []})
                    continue
                sel = [n for n, rl in enumerate(road_links)
                       if (rl["type"] == ta and rl["_in"] == ka) or (rl["type"] == tb and rl["_in"] == kb)]
                phases.append({"time": green_time, "availableRoadLinks": sorted(set(sel) | set(rights))})
            for rl in road_links:
                del rl["_in"]
            # a virtual intersection lists the road arriving from its neighbour, then the one going back
            road_list = [r for r, _ in incoming] + [r for r, _ in outgoing]
            inters.append({"id": f"intersection_{i}_{j}", "point": {"x": _num(cx), "y": _num(cy)},
                           "width": 0 if isv else _num(intersection_width), "roads": road_list, "roadLinks": road_links,
                           "trafficLight": {"roadLinkIndices": list(range(len(road_links))), "lightphases": phases},
                           "virtual": isv})
    return {"intersections": inters, "roads": roads}


def _first_phase_order(road_links, rights):
    """The shipped grids list the right turns of the all-red phase starting from the north approach
    ([10, 2, 3, 6] on a four-way junction): keep that order so the files compare equal."""
    if not rights:
        return []
    return [rights[-1]] + rights[:-1]


# ---------------------------------------------------------------------------------------------
class GridTripGenerator:
    """``CityFlowTripGenerator`` (trip_generator.py:45-286) over a roadnet dict.

    ``generate()`` returns the flow list (CityFlow flow JSON entries, sorted by start time); the
    numpy and ``random`` streams are both seeded with ``seed`` and consumed exactly as the reference
    consumes its global generators."""

    turns = ["turn_left", "turn_right", "go_straight"]
    vehicle_data = {"length": 5.0, "width": 2.0, "maxPosAcc": 2.0, "maxNegAcc": 4.5, "usualPosAcc": 2.0,
                    "usualNegAcc": 4.5, "minGap": 2.5, "maxSpeed": 11.11, "headwayTime": 1.5}

    def __init__(self, roadnet, start_time, end_time, inter_mu, inter_sigma, seed=0, edge_weights=None,
                 turn_probs=(0.1, 0.3, 0.6)):
        self.net = roadnet
        self.start_time, self.end_time = start_time, end_time
        self.inter_mu, self.inter_sigma = inter_mu, inter_sigma
        self.turn_probabilities = list(turn_probs)
        self.np_rng = np.random.RandomState(seed)      # == np.random.seed(seed) + the global functions
        self.py_rng = random.Random(seed)              # == random.seed(seed) + the global functions
        self.max_trip_length = self._max_trip_length()
        self.lane_connectivity_map = {}
        for inter in roadnet["intersections"]:
            for rl in inter.get("roadLinks", []):
                self.lane_connectivity_map.setdefault(rl["startRoad"], {})[rl["type"]] = rl["endRoad"]
        self._set_edge_weights(edge_weights)
        virt = {it["id"] for it in roadnet["intersections"] if it["virtual"]}
        self.incoming_edges = [r["id"] for r in roadnet["roads"]
                               if r["startIntersection"] in virt and r["endIntersection"] not in virt]
        self.outgoing_edges = [r["id"] for r in roadnet["roads"]
                               if r["startIntersection"] not in virt and r["endIntersection"] in virt]

    def _max_trip_length(self):
        """nx.diameter of the intersection digraph + 2 (trip_generator.py:101-116), by BFS."""
        adj = {it["id"]: set() for it in self.net["intersections"]}
        for r in self.net["roads"]:
            adj[r["startIntersection"]].add(r["endIntersection"])
        diameter = 0
        for src in adj:
            dist = {src: 0}
            q = deque([src])
            while q:
                u = q.popleft()
                for v in adj[u]:
                    if v not in dist:
                        dist[v] = dist[u] + 1
                        q.append(v)
            if len(dist) != len(adj):        # not strongly connected: the reference's fallback
                return sum(1 for it in self.net["intersections"] if not it["virtual"]) + 1
            diameter = max(diameter, max(dist.values()))
        return diameter + 2

    def _set_edge_weights(self, given):
        self.edge_weights = {}
        if given is not None:
            for road in self.net["roads"]:
                self.edge_weights[road["id"]] = given.get(road["id"], 0.0)
            return
        speeds = {r["id"]: [l["maxSpeed"] for l in r["lanes"]] for r in self.net["roads"]}
        gmax = max(max(v) for v in speeds.values())
        for road in self.net["roads"]:
            v = speeds[road["id"]]
            self.edge_weights[road["id"]] = np.round(sum(v) / len(v) / gmax, 2)

    def _choose_next_edge(self, cur):
        if cur not in self.lane_connectivity_map:
            return None
        cands, weights = [], []
        for i, direction in enumerate(self.turns):
            nxt = self.lane_connectivity_map[cur].get(direction)
            if nxt:
                cands.append(nxt)
                weights.append(self.turn_probabilities[i] * self.edge_weights.get(nxt, 1.0))
        if not cands:
            return None
        total = sum(weights)
        if total == 0:
            return None
        return self.py_rng.choices(cands, weights=[w / total for w in weights], k=1)[0]

    def _generate_route(self, start_edge):
        route, cur = [start_edge], start_edge
        while True:
            nxt = self._choose_next_edge(cur)
            attempts = 0
            while nxt in route:
                nxt = self._choose_next_edge(cur)
                attempts += 1
                if attempts >= len(self.turns):
                    nxt = None
                    break
            if nxt is None:
                break
            route.append(nxt)
            cur = nxt
            if nxt in self.outgoing_edges:
                break
        return route

    # ---- the arrival loop shared by the whole family (trip_generator.py:251-286 and its six variations):
    #      per start edge, vehicles at Gaussian inter-arrival times; the subclasses change which edges start
    #      vehicles, how the next gap is drawn, and how a route is picked ----
    def _start_edges(self):
        return self.incoming_edges

    def _gap(self, edge, now):
        return self.np_rng.normal(self.inter_mu, self.inter_sigma)

    def _pick_route(self, edge):
        route = [edge]
        while len(route) <= 1 or len(route) > self.max_trip_length:
            route = self._generate_route(edge)
        return route

    def _entry(self, route, t0):
        return {"vehicle": self.vehicle_data, "route": route, "interval": 1.0, "startTime": t0, "endTime": t0}

    def _arrivals(self, edge, now, until, flows):
        """Vehicles entering on ``edge`` from ``now`` until ``until``; returns the time of the last one."""
        while now < until:
            t0 = int(now + max(0, self._gap(edge, now)))
            if t0 >= until or t0 >= self.end_time:
                break
            flows.append(self._entry(self._pick_route(edge), t0))
            now = t0
        return now

    def generate(self):
        flows = []
        for edge in self._start_edges():
            self._arrivals(edge, self.start_time, self.end_time, flows)
        return sorted(flows, key=lambda f: f["startTime"])


class LinkDisruptedTripGenerator(GridTripGenerator):
    """``LinkDisruptedCityFlowTripGenerator`` (trip_generator.py:289-388): a share of the interior roads is closed
    (``int(disruption_ratio * n_signals)`` of them, drawn by shuffling the non-fringe roads); routes avoid them."""

    def __init__(self, roadnet, start_time, end_time, inter_mu, inter_sigma, disruption_ratio=0.1, **kw):
        super().__init__(roadnet, start_time, end_time, inter_mu, inter_sigma, **kw)
        self.disruption_ratio = disruption_ratio
        n_signals = sum(1 for it in roadnet["intersections"] if not it["virtual"])
        fringe = set(self.incoming_edges) | set(self.outgoing_edges)
        interior = [r["id"] for r in roadnet["roads"] if r["id"] not in fringe]
        self.py_rng.shuffle(interior)
        self.disrupted_links = set(interior[: int(disruption_ratio * n_signals)])

    def _choose_next_edge(self, cur):
        if cur not in self.lane_connectivity_map:
            return None
        cands, weights = [], []
        for i, direction in enumerate(self.turns):
            nxt = self.lane_connectivity_map[cur].get(direction)
            if nxt and nxt not in self.disrupted_links:
                cands.append(nxt)
                weights.append(self.turn_probabilities[i] * self.edge_weights.get(nxt, 1.0))
        if not cands:
            return None
        total = sum(weights)
        if total == 0:
            return None
        return self.py_rng.choices(cands, weights=[w / total for w in weights], k=1)[0]


class FlowDisruptedTripGenerator(GridTripGenerator):
    """``FlowDisruptedCityFlowTripGenerator`` (trip_generator.py:391-489): a share of the incoming fringe roads sees
    a ten-minute burst (arrival rate x 4) starting at a random time within the first twenty minutes.

    The reference draws the burst times while iterating a ``set`` of road ids, whose order follows Python's
    per-process string hashing; here they are drawn in the order the roads were sampled (identical whenever
    one road is disrupted, e.g. ``disruption_ratio`` 0.1 on a 3 x 3 grid)."""
    burst_start_min, burst_start_max, burst_duration, burst_multiplier = 0, 20, 10, 4

    def __init__(self, roadnet, start_time, end_time, inter_mu, inter_sigma, disruption_ratio=0.1, **kw):
        super().__init__(roadnet, start_time, end_time, inter_mu, inter_sigma, **kw)
        self.disruption_ratio = disruption_ratio
        picked = self.py_rng.sample(self.incoming_edges, int(len(self.incoming_edges) * disruption_ratio))
        self.disrupted_links = set(picked)
        self.burst_timings = {}
        for link in picked:
            t0 = self.py_rng.randint(self.burst_start_min * 60, self.burst_start_max * 60)
            self.burst_timings[link] = (t0, t0 + self.burst_duration * 60)

    def _gap(self, edge, now):
        b0, b1 = self.burst_timings.get(edge, (float("inf"), float("inf")))
        if edge in self.disrupted_links and b0 <= now < b1:
            return self.np_rng.normal(self.inter_mu / self.burst_multiplier, self.inter_sigma / self.burst_multiplier)
        return self.np_rng.normal(self.inter_mu, self.inter_sigma)


def weibull_flow_rates(np_rng, shape, scale, max_rate, num_segments):
    """``generate_weibull_flow_rates`` (common/utils.py:136-155): a Gaussian-shaped profile of mean inter-arrival
    times over the hour's segments, rolled to a random peak (the Weibull draws only advance the generator)."""
    np_rng.weibull(shape, 1000)
    peak = np_rng.randint(0, num_segments)
    x = np.linspace(-2, 2, num_segments)
    rates = np.exp(-(x ** 2))
    return np.roll(rates / max(rates) * max_rate, peak)


class IntervalTripGenerator(GridTripGenerator):
    """``IntervalCityFlowTripGenerator`` (trip_generator.py:492-554): the mean inter-arrival time changes every
    ``interval_duration`` seconds along a Weibull-placed profile."""

    def generate(self, interval_duration=360, shape=1.5, scale=300):
        self._segment_mean = weibull_flow_rates(self.np_rng, shape, scale, self.inter_mu, int(3600 / interval_duration))
        flows = []
        n_intervals = (self.end_time - self.start_time) // interval_duration
        for edge in self._start_edges():
            now = self.start_time
            for k in range(n_intervals):
                self._mean_now = self._segment_mean[k]
                now = self._arrivals(edge, now, self.start_time + (k + 1) * interval_duration, flows)
        return sorted(flows, key=lambda f: f["startTime"])

    def _gap(self, edge, now):
        return self.np_rng.normal(self._mean_now, self.inter_sigma)


class VariableDemandTripGenerator(GridTripGenerator):
    """``VariableDemandTripGenerator`` (trip_generator.py:557-666): per-edge mean / sigma of the inter-arrival time,
    scaled by a ten-minute demand profile.  The reference does not seed: its ``random`` stream is the one
    ``Config`` left (``random.seed(cityflow.seed)``), its numpy stream whatever the caller set -- ``config_seed`` and
    ``seed`` here."""
    demand_profile = [0.5, 0.6, 0.75, 1.0, 1.0, 0.5, 0.5, 0.3, 0.3, 1e-6]

    def __init__(self, roadnet, start_time, end_time, inter_mus, inter_sigmas, edge_weights, turn_probs=(1 / 3, 1 / 3, 1 / 3),
                 seed=None, config_seed=0):
        super().__init__(roadnet, start_time, end_time, None, None, seed=seed, edge_weights=edge_weights, turn_probs=turn_probs)
        self.py_rng = random.Random(config_seed)
        self.inter_mus, self.inter_sigmas = inter_mus, inter_sigmas

    def _start_edges(self):
        return [e for e in self.incoming_edges if e in self.inter_mus]

    def _gap(self, edge, now):
        slot = (now % 3600) // 600
        return self.np_rng.normal(self.inter_mus[edge] / self.demand_profile[slot], self.inter_sigmas[edge] / self.demand_profile[slot])


class OneWayTripGenerator(GridTripGenerator):
    """``CityFlowOneWayTripGenerator`` (trip_generator.py:669-802): one-way grids; north-south and east-west entry
    roads have their own arrival rates, every vehicle goes straight."""

    def __init__(self, roadnet, start_time, end_time, inter_mu_ns, inter_sigma_ns, inter_mu_ew, inter_sigma_ew,
                 edge_weights=None, seed=0):
        super().__init__(roadnet, start_time, end_time, inter_mu_ns, inter_sigma_ns, seed=seed, edge_weights=edge_weights,
                         turn_probs=(0.0, 0.0, 1.0))
        self.rates = {"ns": (inter_mu_ns, inter_sigma_ns), "ew": (inter_mu_ew, inter_sigma_ew)}
        self._kind = {}
        for r in roadnet["roads"]:
            p0, p1 = r["points"][0], r["points"][-1]
            if p0["x"] == p1["x"] and p0["y"] > p1["y"]:
                self._kind[r["id"]] = "ns"
            elif p0["y"] == p1["y"] and p0["x"] > p1["x"]:
                self._kind[r["id"]] = "ew"

    def _start_edges(self):      # all north-south entries first, then the east-west ones (roadnet order within each)
        inc = set(self.incoming_edges)
        roads = [r["id"] for r in self.net["roads"] if r["id"] in inc]
        return [e for e in roads if self._kind.get(e) == "ns"] + [e for e in roads if self._kind.get(e) == "ew"]

    def _gap(self, edge, now):
        return self.np_rng.normal(*self.rates[self._kind[edge]])


def turn_direction(prev_road, cur_road):
    """``detect_turn_direction`` (trip_generator.py:26-42): from the second field of the road ids."""
    a, b = prev_road.split("_")[1], cur_road.split("_")[1]
    if a == b:
        return "go_straight"
    return "turn_right" if int(b) > int(a) else "turn_left"


class RandomizedTripGenerator(GridTripGenerator):
    """``CityFlowRandomizedTripGenerator`` (trip_generator.py:805-1027): re-samples an existing flow file -- per entry
    road the observed inter-arrival statistics (scaled by ``flow_type``) and the observed routes with their
    frequencies.  ``base_flows`` is that flow file's content.  Streams as for ``VariableDemandTripGenerator``."""

    def __init__(self, roadnet, base_flows, start_time, end_time, seed=None, config_seed=0):
        self.flow_info, self.stored_routes, self.route_proportions, turn_ratios = self._flow_statistics(base_flows)
        super().__init__(roadnet, start_time, end_time, None, None, seed=seed,
                         turn_probs=(turn_ratios["turn_left"], turn_ratios["turn_right"], turn_ratios["go_straight"]))
        self.py_rng = random.Random(config_seed)
        self.max_trip_length = max(v["max_route_length"] for v in self.flow_info.values())

    @staticmethod
    def _flow_statistics(base_flows):
        counts, starts, lengths, routes, route_counts = {}, {}, {}, {}, {}
        turns = {"go_straight": 0, "turn_right": 0, "turn_left": 0}
        all_starts = []
        for veh in base_flows:
            route, t0 = veh["route"], veh["startTime"]
            e = route[0]
            routes.setdefault(e, [])
            if route not in routes[e]:
                routes[e].append(route)
            rc = route_counts.setdefault(e, {})
            rc[tuple(route)] = rc.get(tuple(route), 0) + 1
            counts[e] = counts.get(e, 0) + 1
            starts.setdefault(e, []).append(t0)
            lengths.setdefault(e, []).append(len(route))
            all_starts.append(t0)
            for i in range(1, len(route)):
                turns[turn_direction(route[i - 1], route[i])] += 1
        proportions = {e: {r: n / sum(rc.values()) for r, n in rc.items()} for e, rc in route_counts.items()}
        total_turns = sum(turns.values())
        turns = {k: v / total_turns for k, v in turns.items()}
        hours = (max(all_starts) - min(all_starts)) / 3600
        info = {}
        for e in counts:
            d = np.diff(sorted(starts[e]))
            info[e] = {"flow_rate": counts[e] / hours, "arrival_diff_mean": np.mean(d), "arrival_diff_std": np.std(d),
                       "mean_route_length": np.mean(lengths[e]), "min_route_length": np.min(lengths[e]),
                       "max_route_length": np.max(lengths[e])}
        return info, routes, proportions, turns

    def _start_edges(self):
        return [e for e in self.incoming_edges if e in self.flow_info]

    def _gap(self, edge, now):
        m = self.flow_info[edge]["arrival_diff_mean"]
        if self.flow_type == "low":
            m = m + self.np_rng.uniform(-0.5, 0.0)
        elif self.flow_type == "medium":
            m = m * 0.75
        elif self.flow_type == "high":
            m = m * 0.5
        else:
            raise ValueError("Invalid flow type")
        return self.np_rng.normal(m, self.flow_info[edge]["arrival_diff_std"])

    def _pick_route(self, edge):
        routes = self.stored_routes[edge]
        return self.py_rng.choices(routes, weights=[self.route_proportions[edge][tuple(r)] for r in routes], k=1)[0]

    def generate(self, flow_type="low"):
        self.flow_type = flow_type
        return super().generate()


def synthetic_grid_scenario(rows, cols, vehicles_per_hour_per_road=900, horizon=3600, sigma=0.8, seed=0,
                            turn_probs=(0.1, 0.3, 0.6), **roadnet_kwargs):
    """Roadnet + flows of a synthetic grid (BASELINE config 4: 16 x 16, heavy demand = 900 vehicles per
    hour on every incoming fringe road)."""
    net = grid_roadnet(rows, cols, **roadnet_kwargs)
    gen = GridTripGenerator(net, 0, horizon, 3600.0 / vehicles_per_hour_per_road, sigma, seed=seed, turn_probs=turn_probs)
    return net, gen.generate()


def write_grid_scenario(directory, rows, cols, vehicles_per_hour_per_road=900, horizon=3600, sigma=0.8, seed=0,
                        turn_probs=(0.1, 0.3, 0.6), signal=None, **roadnet_kwargs):
    """Write ``<directory>/syn_{rows}x{cols}/`` -- roadnet and flow bundles plus a ``config.yaml`` in the
    layout of the shipped scenarios -- and return its path, usable wherever a scenario name is
    (``Config(path)``, ``BatchedTrafficSignalNetwork(path, ...)``)."""
    import os

    import yaml

    from . import bundle
    net, flows = synthetic_grid_scenario(rows, cols, vehicles_per_hour_per_road, horizon, sigma, seed, turn_probs,
                                         **roadnet_kwargs)
    d = os.path.join(str(directory), f"syn_{rows}x{cols}")
    os.makedirs(d, exist_ok=True)
    roadnet_file = f"{rows}x{cols}_roadnet.npz"
    flow_file = f"syn_{rows}x{cols}__gaussian_{int(vehicles_per_hour_per_road)}_flows.npz"
    bundle.save_npz(os.path.join(d, roadnet_file), bundle.pack_roadnet(net))
    bundle.save_npz(os.path.join(d, flow_file), bundle.pack_flow(flows, net))
    cfg = {"cityflow": {"roadnet_file": roadnet_file, "flow_file": flow_file, "flow_rate_type": "constant",
                        "save_replay": False, "sim_length": int(horizon), "episode_limit": min(int(horizon), 3600)},
           "signal": dict({"action_space": "phase_selection", "round_robin": False, "reward_function": "queue_length"},
                          **(signal or {}))}
    with open(os.path.join(d, "config.yaml"), "w") as f:
        f.write("# generated by pytsc_b200.generators.write_grid_scenario\n")
        yaml.safe_dump(cfg, f, sort_keys=False)
    return d
