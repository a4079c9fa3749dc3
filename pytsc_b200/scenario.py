"""Scenario compiler: roadnet + flow + pytsc signal configuration -> the flat,
read-only tables of ``tsc_scenario_t`` (include/tsc_b200.h).

Host-side, once per scenario.  Two independent sources are lowered here:

* the CityFlow engine view (``roadnet.RoadNet`` / ``roadnet.expand_flows``):
  drivable lengths, lane-link topology, crosses, phase -> road-link masks,
  per-route drivable sequences, the spawn list;
* the pytsc view (``backend.network_parser.NetworkParser``): incoming /
  outgoing lanes per signal in pytsc's order, the 2*n_green phase plan with
  min/max times, static lane features, k-hop reward neighbourhoods.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import bundle
from .roadnet import RoadNet, VEHICLE_DEFAULTS, VEHICLE_KEYS, expand_flows

ABI_VERSION = 3
T_STRIDE = 12
REWARD_TYPES = {"queue_length": 0, "max_pressure": 1}
OBS_TYPES = {"lane_features": 0, "position_matrix": 1}
ACTION_SPACES = {"phase_selection": 0, "phase_switch": 1}

# observation layout constants (pytsc/common/observations.py:57-61, 227-231)
MAX_LANES_PER_DIRECTION = 6
MAX_LANE_SPEED = 15.0
MAX_LANE_LENGTH = 500
MAX_PHASES = 20
MAX_N_CONTROLLED_LANES = 16


class tsc_scenario_t(C.Structure):
    _i, _pd, _pi, _pu8, _pu32 = C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
    _fields_ = [
        ("abi_version", _i), ("n_lanes", _i), ("n_lanelinks", _i), ("n_signals", _i), ("n_vehicles", _i),
        ("n_templates", _i), ("n_route_seq", _i), ("n_cross_entries", _i), ("horizon_ticks", _i),
        ("max_raw_phases", _i), ("max_phases", _i), ("n_in_total", _i), ("n_out_total", _i), ("n_nbr_total", _i),
        ("n_ctl_total", _i), ("n_dm_total", _i), ("n_flow_sets", _i),
        ("drv_length", _pd), ("drv_max_speed", _pd), ("lane_ll_off", _pi), ("lane_ll", _pi),
        ("lane_spawn_off", _pi), ("lane_spawn_vid", _pi), ("ll_start_lane", _pi), ("ll_end_lane", _pi),
        ("ll_signal", _pi), ("ll_roadlink", _pi), ("ll_type", _pi), ("ll_cross_off", _pi),
        ("xr_dist", _pd), ("xr_foe_ll", _pi), ("xr_foe_dist", _pd), ("sig_phase_mask", _pu32),
        ("sig_n_raw_phases", _pi), ("route_seq", _pi), ("veh_tick", _pi), ("veh_seq_start", _pi),
        ("veh_tmpl", _pi), ("veh_priority", _pi), ("tmpl", _pd),
        ("lane_pytsc_length", _pd), ("lane_feat", _pd), ("sig_in_off", _pi), ("sig_in_lane", _pi),
        ("sig_out_off", _pi), ("sig_out_lane", _pi), ("sig_n_phases", _pi), ("sig_phase_raw", _pi),
        ("sig_phase_green", _pu8), ("sig_min_time", _pi), ("sig_max_time", _pi),
        ("nbr_off", _pi), ("nbr_idx", _pi), ("nbr_weight", _pd),
        ("ctl_off", _pi), ("ctl_in_lane", _pi), ("ctl_out_lane", _pi),
        ("dm_off", _pi), ("dm_lane", _pi), ("dm_adjacency", _pd),
        ("reward_type", _i), ("obs_type", _i), ("action_space", _i), ("round_robin", _i), ("visibility", _i),
        ("yellow_time", _i), ("obs_dim", _i), ("state_dim", _i), ("n_actions", _i), ("reference_exact", _i),
        ("max_lanes_per_signal", _i), ("max_obs_phases", _i),
        ("veh_size_min_gap", C.c_double), ("flickering_coef", C.c_double), ("interval", C.c_double),
    ]


_CT = {np.dtype(np.float64): C.c_double, np.dtype(np.int32): C.c_int32, np.dtype(np.uint8): C.c_uint8,
       np.dtype(np.uint32): C.c_uint32}


@dataclass
class CompiledScenario:
    arrays: dict
    scalars: dict
    lane_ids: list            # engine lane order (roadnet order)
    signal_ids: list          # agent order (roadnet order of non-virtual intersections)
    vehicle_names: list = field(default_factory=list)
    stats: dict = field(default_factory=dict)

    def __getattr__(self, k):
        if k in ("arrays", "scalars"):
            raise AttributeError(k)
        if k in self.arrays:
            return self.arrays[k]
        if k in self.scalars:
            return self.scalars[k]
        raise AttributeError(k)

    def to_struct(self) -> tsc_scenario_t:
        s = tsc_scenario_t()
        for name, ctype in tsc_scenario_t._fields_:
            if name in self.arrays:
                a = self.arrays[name]
                assert a.flags["C_CONTIGUOUS"], name
                setattr(s, name, a.ctypes.data_as(C.POINTER(_CT[a.dtype])))
            else:
                setattr(s, name, self.scalars[name])
        s._keepalive = self
        return s


def _csr(lists, dtype=np.int32):
    off = np.zeros(len(lists) + 1, np.int32)
    for i, l in enumerate(lists):
        off[i + 1] = off[i] + len(l)
    flat = np.asarray([x for l in lists for x in l], dtype=dtype)
    if flat.size == 0:
        flat = np.zeros(1, dtype)
    return off, flat


def static_lane_features(parser):
    """observations.py:90-116 / 260-286: [len/500, angle/pi, vmax/15] clipped, + one-hot lane index (6)."""
    feats = {}
    for lane in parser.lanes:
        one_hot = [0.0] * MAX_LANES_PER_DIRECTION
        one_hot[parser.lane_indices[lane]] = 1.0
        ll = np.clip(parser.lane_lengths[lane] / MAX_LANE_LENGTH, 0, 1)
        la = np.clip(parser.lane_angles[lane] / np.pi, -1, 1)
        ls = np.clip(parser.lane_max_speeds[lane] / MAX_LANE_SPEED, 0, 1)
        feats[lane] = [float(ll), float(la), float(ls)] + one_hot
    return feats


def derive_vehicle_capacity(cs, margin=1.1) -> int:
    """Upper bound on simultaneously running vehicles per replica when ``gpu.vehicle_capacity`` is 0: no more
    than a flow set creates in total, and no more than bumper to bumper on every drivable (vehicle length +
    minGap per vehicle, the standing spacing CityFlow's car-following rule converges to), with a margin.
    Safe rather than tight: a smaller explicit capacity keeps more replicas per SM (bench.py passes one);
    exceeding it freezes the replica and is reported by ``Engine.check()`` / the ``err`` output."""
    off = cs.stats.get("flow_set_off", [0, int(cs.n_vehicles)])
    created = max(int(off[k + 1] - off[k]) for k in range(len(off) - 1))
    tm = np.asarray(cs.tmpl).reshape(-1, T_STRIDE)
    spacing = float((tm[:, 0] + tm[:, 5]).min()) if tm.size else 7.5
    packed = int(sum(np.ceil(np.asarray(cs.drv_length) / max(spacing, 1.0)) + 1))
    return int(min(32000, max(64, min(created, int(np.ceil(packed * margin))))))


def merge_flow_sets(sps):
    """Spawn lists of several flow files (``expand_flows`` results) as ONE vehicle table, set-major: the
    vehicles of set f follow those of set f-1; routes and vehicle templates are shared.  Adds ``set_off``
    ([F+1] vehicle index ranges)."""
    if len(sps) == 1:
        sp = dict(sps[0])
        sp["set_off"] = [0, len(sp["tick"])]
        return sp
    routes, route_index, templates, tmpl_index = [], {}, [], {}
    out = {k: [] for k in ("tick", "flow", "flow_cnt", "route", "tmpl", "priority", "first_lane")}
    set_off, dup, invalid = [0], 0, []
    for sp in sps:
        rmap = []
        for key in sp["routes"]:
            if key not in route_index:
                route_index[key] = len(routes)
                routes.append(key)
            rmap.append(route_index[key])
        tmap = []
        for t in sp["templates"]:
            if t not in tmpl_index:
                tmpl_index[t] = len(templates)
                templates.append(t)
            tmap.append(tmpl_index[t])
        for k in ("tick", "flow", "flow_cnt", "priority", "first_lane"):
            out[k].append(np.asarray(sp[k], np.int64))
        out["route"].append(np.asarray([rmap[r] for r in sp["route"]], np.int64))
        out["tmpl"].append(np.asarray([tmap[t] for t in sp["tmpl"]], np.int64))
        set_off.append(set_off[-1] + len(sp["tick"]))
        dup += sp["duplicate_priorities"]
        invalid.append(sp["invalid_flows"])
    res = {k: np.concatenate(v) if v else np.zeros(0, np.int64) for k, v in out.items()}
    res.update(routes=routes, templates=templates, duplicate_priorities=dup, invalid_flows=invalid, set_off=set_off)
    return res


def compile_scenario(config, parser, flows=None, flow_file=None, flow_sets=None) -> CompiledScenario:
    """``config``: backend.config.Config; ``parser``: backend.network_parser.NetworkParser.

    One flow set by default -- ``flows`` (a CityFlow flow list), else ``flow_file``, else the file the config
    picks (``backends/cityflow/config.py:63-76``).  ``flow_sets`` = a list of flow lists / file paths compiles
    several alternatives into one scenario (``n_flow_sets``): replicas are assigned one each at reset
    (``tsc_reset_flows``), which is how ``flow_rate_type: random | sequential`` and ``DisruptedConfig`` map
    onto a batch."""
    sim, sig, misc = config.simulator, config.signal, config.misc
    rn = RoadNet(parser.net)
    horizon = int(sim["sim_length"]) + int(sim["initial_wait_time"])
    if flow_sets is None:
        if flows is None:
            flows = bundle.load_flow(flow_file or config.create_and_save_cityflow_cfg())
        flow_sets = [flows]
    flow_sets = [bundle.load_flow(f) if isinstance(f, str) else f for f in flow_sets]
    if not flow_sets:
        raise ValueError("compile_scenario: empty flow_sets")
    # every flow file is what a fresh cityflow.Engine(seed) would make of it: the generator restarts per set
    sp = merge_flow_sets([expand_flows(rn, f, float(sim["interval"]), int(sim["seed"]), horizon) for f in flow_sets])
    F = len(flow_sets)
    if not sp["templates"]:      # no vehicle at all: CityFlow's default vehicle stands in (the tables must be well formed)
        sp["templates"] = [tuple(float(VEHICLE_DEFAULTS[k]) for k in VEHICLE_KEYS)]
    L, K = len(rn.lanes), len(rn.lanelinks)
    a, s = {}, {}
    f64, i32 = np.float64, np.int32

    # ---- engine tables ---------------------------------------------------------
    a["drv_length"] = np.asarray([l.length for l in rn.lanes] + [l.length for l in rn.lanelinks], f64)
    a["drv_max_speed"] = np.asarray([l.max_speed for l in rn.lanes] + [10000.0] * K, f64)
    a["lane_ll_off"], a["lane_ll"] = _csr([l.lanelinks for l in rn.lanes])
    signal_ids = [iid for iid, v in zip(rn.inter_ids, rn.inter_virtual) if not v]
    sig_of_inter = {rn.inter_index[iid]: k for k, iid in enumerate(signal_ids)}
    A = len(signal_ids)
    a["ll_start_lane"] = np.asarray([l.start_lane for l in rn.lanelinks], i32)
    a["ll_end_lane"] = np.asarray([l.end_lane for l in rn.lanelinks], i32)
    a["ll_signal"] = np.asarray([sig_of_inter[l.inter] for l in rn.lanelinks], i32)
    a["ll_roadlink"] = np.asarray([l.roadlink for l in rn.lanelinks], i32)
    a["ll_type"] = np.asarray([l.rl_type for l in rn.lanelinks], i32)
    a["ll_cross_off"], _ = _csr([l.crosses for l in rn.lanelinks])
    xs = [c for l in rn.lanelinks for c in l.crosses]
    a["xr_dist"] = np.asarray([c[0] for c in xs] or [0.0], f64)
    a["xr_foe_ll"] = np.asarray([c[1] for c in xs] or [0], i32)
    a["xr_foe_dist"] = np.asarray([c[2] for c in xs] or [0.0], f64)
    max_raw = max(len(rn.inter_phases[rn.inter_index[t]]) for t in signal_ids)
    mask = np.zeros((A, max_raw), np.uint32)
    nraw = np.zeros(A, i32)
    for k, t in enumerate(signal_ids):
        ii = rn.inter_index[t]
        if rn.inter_n_roadlinks[ii] > 32:
            raise ValueError("more than 32 road-links at one intersection")
        nraw[k] = len(rn.inter_phases[ii])
        for p, (_, avail) in enumerate(rn.inter_phases[ii]):
            for r in avail:
                mask[k, p] |= np.uint32(1 << r)
    a["sig_phase_mask"], a["sig_n_raw_phases"] = mask.reshape(-1), nraw
    # routes: -1 separated drivable sequences
    seq, starts = [-1], []
    for route, first_lane in sp["routes"]:
        starts.append(len(seq))
        seq += rn.drivable_sequence(list(route), first_lane)
        seq.append(-1)
    a["route_seq"] = np.asarray(seq, i32)
    N = len(sp["tick"])
    a["veh_tick"] = sp["tick"].astype(i32) if N else np.zeros(1, i32)
    a["veh_seq_start"] = np.asarray([starts[r] for r in sp["route"]] or [0], i32)
    a["veh_tmpl"] = sp["tmpl"].astype(i32) if N else np.zeros(1, i32)
    a["veh_priority"] = sp["priority"].astype(i32) if N else np.zeros(1, i32)
    per_lane = [[] for _ in range(L)]      # union over the flow sets (which lanes spawn at all)
    spawn_off = np.zeros((F, L + 1), i32)
    spawn_vid = []
    for f in range(F):
        lanes_f = [[] for _ in range(L)]
        for v in range(sp["set_off"][f], sp["set_off"][f + 1]):
            lanes_f[int(sp["first_lane"][v])].append(v)
            per_lane[int(sp["first_lane"][v])].append(v)
        spawn_off[f, 0] = len(spawn_vid)
        for l in range(L):
            spawn_vid += lanes_f[l]
            spawn_off[f, l + 1] = len(spawn_vid)
    a["lane_spawn_off"] = spawn_off.reshape(-1)
    a["lane_spawn_vid"] = np.asarray(spawn_vid or [0], i32)
    interval = float(sim["interval"])
    tm = np.zeros((max(len(sp["templates"]), 1), T_STRIDE), f64)
    for i, t in enumerate(sp["templates"]):
        p = dict(zip(VEHICLE_KEYS, t))
        tm[i, :10] = [p["length"], p["maxPosAcc"], p["maxNegAcc"], p["usualPosAcc"], p["usualNegAcc"], p["minGap"],
                      p["maxSpeed"], p["headwayTime"], p["yieldDistance"], p["turnSpeed"]]
        tm[i, 10] = p["maxSpeed"] * p["maxSpeed"] / p["usualNegAcc"] / 2 + p["maxSpeed"] * interval * 2
        tm[i, 11] = p["width"]
    a["tmpl"] = tm.reshape(-1)
    # the data-parallel spawn step assumes a freshly inserted vehicle can never be
    # another lane's look-ahead leader inside the same tick
    for i, t in enumerate(sp["templates"]):
        thr = tm[i, 10]
        for l in range(L):
            if not per_lane[l]:
                continue
            nxt = min([rn.lanelinks[k].length for k in rn.lanes[l].lanelinks] or [float("inf")])
            if rn.lanes[l].length + nxt <= thr:
                raise ValueError(f"spawn lane {rn.lanes[l].id} is shorter than the look-ahead horizon")

    # ---- pytsc tables ------------------------------------------------------------
    lane_idx = {l.id: l.index for l in rn.lanes}
    lane_ids = [l.id for l in rn.lanes]
    if list(parser.traffic_signals.keys()) != signal_ids:
        raise ValueError("signal order mismatch between engine and pytsc views")
    a["lane_pytsc_length"] = np.asarray([parser.lane_lengths[l] for l in lane_ids], f64)
    feats = static_lane_features(parser)
    a["lane_feat"] = np.asarray([feats[l] for l in lane_ids], f64).reshape(-1)
    ts = parser.traffic_signals
    a["sig_in_off"], a["sig_in_lane"] = _csr([[lane_idx[l] for l in ts[t]["incoming_lanes"]] for t in signal_ids])
    a["sig_out_off"], a["sig_out_lane"] = _csr([[lane_idx[l] for l in ts[t]["outgoing_lanes"]] for t in signal_ids])
    P = max(ts[t]["n_phases"] for t in signal_ids)
    nph = np.zeros(A, i32)
    raw = np.zeros((A, P), i32)
    grn = np.zeros((A, P), np.uint8)
    mn = np.zeros((A, P), i32)
    mx = np.ones((A, P), i32)
    for k, t in enumerate(signal_ids):
        cfg = ts[t]
        nph[k] = cfg["n_phases"]
        for p, r in enumerate(cfg["phases"]):
            raw[k, p] = r
            grn[k, p] = 1 if p in cfg["green_phase_indices"] else 0
            mn[k, p] = cfg["phases_min_max_times"][r]["min_time"]
            mx[k, p] = cfg["phases_min_max_times"][r]["max_time"]
    a["sig_n_phases"], a["sig_phase_raw"], a["sig_phase_green"] = nph, raw.reshape(-1), grn.reshape(-1)
    a["sig_min_time"], a["sig_max_time"] = mn.reshape(-1), mx.reshape(-1)
    # reward neighbourhoods in the reference's summation order (reward.py:81-88):
    # k = 1 .. n_signals-1, neighbours as listed by k_hop_neighbors[ts][k]
    gamma = misc["reward_gamma"]
    sidx = {t: k for k, t in enumerate(signal_ids)}
    nbr, wts = [], []
    for t in signal_ids:
        ids, w = [], []
        for k in range(1, len(signal_ids)):
            for nb in parser.k_hop_neighbors[t].get(k, []):
                ids.append(sidx[nb])
                w.append(gamma ** k)
        nbr.append(ids)
        wts.append(w)
    a["nbr_off"], a["nbr_idx"] = _csr(nbr)
    _, a["nbr_weight"] = _csr(wts, f64)
    # rule-based controllers (controllers/controllers.py:95-114, 151-176, 222-238): the incoming lanes a
    # pytsc phase serves and, per incoming lane, the LAST outgoing lane listed for it -- MaxPressure's
    # inner loop overwrites instead of accumulating, so only that one counts
    ctl_in, ctl_out = [], []
    for t in signal_ids:
        cfg = ts[t]
        for p in range(P):
            ins, outs = [], []
            if p < cfg["n_phases"]:
                for inc_lane, out_lanes in cfg["phase_to_inc_out_lanes"].get(cfg["phases"][p], {}).items():
                    ins.append(lane_idx[inc_lane])
                    outs.append(lane_idx[out_lanes[-1]] if out_lanes else -1)
            ctl_in.append(ins)
            ctl_out.append(outs)
    a["ctl_off"], a["ctl_in_lane"] = _csr(ctl_in)
    _, a["ctl_out_lane"] = _csr(ctl_out)

    # MetricsParser.density_map (backends/cityflow/metrics.py:170-199): lanes from signal i to signal j in agent order,
    # and the adjacency matrix exactly as the reference adds it (its rows follow the SORTED ids: SURVEY B5)
    nl = parser.neighbors_lanes
    a["dm_off"], a["dm_lane"] = _csr([[lane_idx[l] for l in ((nl.get(ti) or {}).get(tj) or [])]
                                      for ti in signal_ids for tj in signal_ids])
    a["dm_adjacency"] = np.asarray(parser.adjacency_matrix, f64).reshape(-1)

    vis = int(sig["visibility"])
    obs_type = OBS_TYPES[sig["observation_space"]]
    state_dim = MAX_N_CONTROLLED_LANES * 12 + MAX_PHASES
    obs_dim = state_dim if obs_type == 0 else MAX_N_CONTROLLED_LANES * (vis + 9) + MAX_PHASES
    act = ACTION_SPACES[sig["action_space"]]
    s.update(abi_version=ABI_VERSION, n_lanes=L, n_lanelinks=K, n_signals=A, n_vehicles=N,
             n_templates=len(sp["templates"]) or 1, n_route_seq=len(seq), n_cross_entries=len(xs),
             horizon_ticks=horizon, max_raw_phases=max_raw, max_phases=P,
             n_in_total=int(a["sig_in_off"][-1]), n_out_total=int(a["sig_out_off"][-1]),
             n_nbr_total=int(a["nbr_off"][-1]), n_ctl_total=int(a["ctl_off"][-1]), n_dm_total=int(a["dm_off"][-1]), n_flow_sets=F,
             reward_type=REWARD_TYPES[sig["reward_function"]], obs_type=obs_type, action_space=act,
             round_robin=int(bool(sig["round_robin"])), visibility=vis, yellow_time=int(sig["yellow_time"]),
             obs_dim=obs_dim, state_dim=state_dim, n_actions=(P if act == 0 else 2),
             reference_exact=int(bool(config.gpu.get("reference_exact", True))),
             max_lanes_per_signal=MAX_N_CONTROLLED_LANES, max_obs_phases=MAX_PHASES,
             veh_size_min_gap=float(sim["veh_size_min_gap"]), flickering_coef=float(misc["flickering_coef"]),
             interval=interval)
    for k, v in a.items():
        a[k] = np.ascontiguousarray(v)
    names = [f"flow_{f}_{c}" for f, c in zip(sp["flow"], sp["flow_cnt"])]
    stats = dict(flow_set_off=[int(x) for x in sp["set_off"]],
                 n_routes=len(sp["routes"]), n_crosses=rn.n_crosses, duplicate_priorities=sp["duplicate_priorities"],
                 invalid_flows=sp["invalid_flows"], n_spawn_lanes=sum(1 for p in per_lane if p))
    return CompiledScenario(a, s, lane_ids, signal_ids, names, stats)
