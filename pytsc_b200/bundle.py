"""Compact, lossless (for the fields the engine and pytsc read) ``.npz`` form of
CityFlow roadnet / flow JSON files.

The real-city inputs the reference ships (``pytsc/scenarios/cityflow/*``) are
1-3 MB of JSON each; the same information as typed arrays is 10-50x smaller and
travels with the repo to machines where the reference checkout does not exist.
``unpack_*`` rebuilds a dict in the CityFlow JSON schema,
so the rest of the package has a single input path.
"""
from __future__ import annotations

import io

import numpy as np

from .roadnet import VEHICLE_KEYS

_RL_TYPES = ("turn_right", "turn_left", "go_straight")   # index + 1 == CityFlow enum value


def _ragged(lists, dtype):
    off = np.zeros(len(lists) + 1, np.int64)
    for i, l in enumerate(lists):
        off[i + 1] = off[i] + len(l)
    flat = np.asarray([x for l in lists for x in l], dtype=dtype)
    return off, flat


def pack_roadnet(net: dict) -> dict:
    ints, roads = net["intersections"], net["roads"]
    road_idx = {r["id"]: i for i, r in enumerate(roads)}
    int_idx = {it["id"]: i for i, it in enumerate(ints)}
    a = {}
    a["inter_id"] = np.asarray([it["id"] for it in ints])
    a["inter_xy"] = np.asarray([[it["point"]["x"], it["point"]["y"]] for it in ints], np.float64)
    a["inter_width"] = np.asarray([it.get("width", 0) for it in ints], np.float64)
    a["inter_virtual"] = np.asarray([bool(it.get("virtual", False)) for it in ints])
    a["inter_roads_off"], a["inter_roads"] = _ragged([[road_idx[r] for r in it["roads"]] for it in ints], np.int32)
    rl_inter, rl_type, rl_sr, rl_er, rl_dir = [], [], [], [], []
    ll_rl, ll_s, ll_e, ll_pts = [], [], [], []
    ph_inter, ph_time, ph_avail = [], [], []
    tl_idx = []
    for ii, it in enumerate(ints):
        tl_idx.append(list(it.get("trafficLight", {}).get("roadLinkIndices", [])))
        for rl in it.get("roadLinks", []):
            rid = len(rl_inter)
            rl_inter.append(ii)
            rl_type.append(_RL_TYPES.index(rl["type"]))
            rl_sr.append(road_idx[rl["startRoad"]])
            rl_er.append(road_idx[rl["endRoad"]])
            rl_dir.append(int(rl.get("direction", 0)))
            for ll in rl["laneLinks"]:
                ll_rl.append(rid)
                ll_s.append(int(ll["startLaneIndex"]))
                ll_e.append(int(ll["endLaneIndex"]))
                ll_pts.append([(p["x"], p["y"]) for p in ll.get("points", [])])
        for ph in it.get("trafficLight", {}).get("lightphases", []):
            ph_inter.append(ii)
            ph_time.append(ph["time"])
            ph_avail.append([int(x) for x in ph["availableRoadLinks"]])
    a["rl_inter"] = np.asarray(rl_inter, np.int32)
    a["rl_type"] = np.asarray(rl_type, np.int8)
    a["rl_start_road"] = np.asarray(rl_sr, np.int32)
    a["rl_end_road"] = np.asarray(rl_er, np.int32)
    a["rl_direction"] = np.asarray(rl_dir, np.int32)
    a["ll_rl"] = np.asarray(ll_rl, np.int32)
    a["ll_start"] = np.asarray(ll_s, np.int16)
    a["ll_end"] = np.asarray(ll_e, np.int16)
    a["ll_pts_off"], pts = _ragged(ll_pts, np.float64)
    a["ll_pts"] = pts.reshape(-1, 2)
    a["ph_inter"] = np.asarray(ph_inter, np.int32)
    a["ph_time"] = np.asarray(ph_time, np.float64)
    a["ph_avail_off"], a["ph_avail"] = _ragged(ph_avail, np.int16)
    a["tl_idx_off"], a["tl_idx"] = _ragged(tl_idx, np.int16)
    a["road_id"] = np.asarray([r["id"] for r in roads])
    a["road_start"] = np.asarray([int_idx[r["startIntersection"]] for r in roads], np.int32)
    a["road_end"] = np.asarray([int_idx[r["endIntersection"]] for r in roads], np.int32)
    a["road_pts_off"], rp = _ragged([[(p["x"], p["y"]) for p in r["points"]] for r in roads], np.float64)
    a["road_pts"] = rp.reshape(-1, 2)
    a["lane_off"], lw = _ragged([[(l["width"], l["maxSpeed"]) for l in r["lanes"]] for r in roads], np.float64)
    a["lane_wm"] = lw.reshape(-1, 2)
    return a


def _num(x):
    x = float(x)
    return int(x) if x == int(x) and abs(x) < 1e15 else x


def unpack_roadnet(a) -> dict:
    road_ids = [str(x) for x in a["road_id"]]
    inter_ids = [str(x) for x in a["inter_id"]]
    roads = []
    for i, rid in enumerate(road_ids):
        p0, p1 = a["road_pts_off"][i], a["road_pts_off"][i + 1]
        l0, l1 = a["lane_off"][i], a["lane_off"][i + 1]
        roads.append({
            "id": rid,
            "points": [{"x": _num(x), "y": _num(y)} for x, y in a["road_pts"][p0:p1]],
            "lanes": [{"width": _num(w), "maxSpeed": float(m)} for w, m in a["lane_wm"][l0:l1]],
            "startIntersection": inter_ids[a["road_start"][i]],
            "endIntersection": inter_ids[a["road_end"][i]],
        })
    ints = []
    for i, iid in enumerate(inter_ids):
        r0, r1 = a["inter_roads_off"][i], a["inter_roads_off"][i + 1]
        ints.append({
            "id": iid,
            "point": {"x": _num(a["inter_xy"][i, 0]), "y": _num(a["inter_xy"][i, 1])},
            "width": _num(a["inter_width"][i]),
            "roads": [road_ids[k] for k in a["inter_roads"][r0:r1]],
            "roadLinks": [],
            "trafficLight": {"roadLinkIndices": [int(x) for x in a["tl_idx"][a["tl_idx_off"][i]:a["tl_idx_off"][i + 1]]],
                             "lightphases": []},
            "virtual": bool(a["inter_virtual"][i]),
        })
    rl_objs = []
    for k in range(len(a["rl_inter"])):
        rl = {"type": _RL_TYPES[a["rl_type"][k]], "startRoad": road_ids[a["rl_start_road"][k]],
              "endRoad": road_ids[a["rl_end_road"][k]], "direction": int(a["rl_direction"][k]), "laneLinks": []}
        ints[a["rl_inter"][k]]["roadLinks"].append(rl)
        rl_objs.append(rl)
    for k in range(len(a["ll_rl"])):
        p0, p1 = a["ll_pts_off"][k], a["ll_pts_off"][k + 1]
        rl_objs[a["ll_rl"][k]]["laneLinks"].append({
            "startLaneIndex": int(a["ll_start"][k]), "endLaneIndex": int(a["ll_end"][k]),
            "points": [{"x": float(x), "y": float(y)} for x, y in a["ll_pts"][p0:p1]]})
    for k in range(len(a["ph_inter"])):
        v0, v1 = a["ph_avail_off"][k], a["ph_avail_off"][k + 1]
        ints[a["ph_inter"][k]]["trafficLight"]["lightphases"].append(
            {"time": _num(a["ph_time"][k]), "availableRoadLinks": [int(x) for x in a["ph_avail"][v0:v1]]})
    return {"intersections": ints, "roads": roads}


def pack_flow(flows: list, net: dict) -> dict:
    road_idx = {r["id"]: i for i, r in enumerate(net["roads"])}
    tmpls, tidx = [], {}
    ft, routes, iv, st, en = [], [], [], [], []
    for f in flows:
        v = f.get("vehicle", {})
        key = tuple((k, float(v[k])) for k in sorted(v))
        if key not in tidx:
            tidx[key] = len(tmpls)
            tmpls.append(dict(key))
        ft.append(tidx[key])
        routes.append([road_idx[r] for r in f["route"]])
        iv.append(f.get("interval", 1.0))
        st.append(f.get("startTime", 0))
        en.append(f.get("endTime", -1))
    keys = sorted({k for t in tmpls for k in t})
    tv = np.full((len(tmpls), len(keys)), np.nan)
    for i, t in enumerate(tmpls):
        for j, k in enumerate(keys):
            if k in t:
                tv[i, j] = t[k]
    a = {"tmpl_keys": np.asarray(keys), "tmpl_vals": tv, "flow_tmpl": np.asarray(ft, np.int32),
         "flow_interval": np.asarray(iv, np.float64), "flow_start": np.asarray(st, np.float64),
         "flow_end": np.asarray(en, np.float64), "road_id": np.asarray([r["id"] for r in net["roads"]])}
    a["route_off"], a["route"] = _ragged(routes, np.int32)
    return a


def unpack_flow(a) -> list:
    road_ids = [str(x) for x in a["road_id"]]
    keys = [str(k) for k in a["tmpl_keys"]]
    tmpls = []
    for row in a["tmpl_vals"]:
        tmpls.append({k: _num(v) if k in ("headwayTime",) else float(v)
                      for k, v in zip(keys, row) if not np.isnan(v)})
    out = []
    for i in range(len(a["flow_tmpl"])):
        r0, r1 = a["route_off"][i], a["route_off"][i + 1]
        out.append({"vehicle": dict(tmpls[a["flow_tmpl"][i]]),
                    "route": [road_ids[k] for k in a["route"][r0:r1]],
                    "interval": float(a["flow_interval"][i]),
                    "startTime": _num(a["flow_start"][i]), "endTime": _num(a["flow_end"][i])})
    return out


def save_npz(path, arrays: dict):
    buf = io.BytesIO()
    np.savez_compressed(buf, **arrays)
    with open(path, "wb") as f:
        f.write(buf.getvalue())


def load_roadnet(path) -> dict:
    """Roadnet dict from ``*.json``, ``*.json.gz`` or a ``*.npz`` bundle."""
    from .roadnet import load_json
    if str(path).endswith(".npz"):
        with np.load(path, allow_pickle=False) as z:
            return unpack_roadnet({k: z[k] for k in z.files})
    return load_json(path)


def load_flow(path) -> list:
    from .roadnet import load_json
    if str(path).endswith(".npz"):
        with np.load(path, allow_pickle=False) as z:
            return unpack_flow({k: z[k] for k in z.files})
    return load_json(path)
