"""TEST INFRASTRUCTURE -- ``cityflow.Engine``-shaped adapter over the CPU oracle.

Gives ``oracle/cityflow_oracle.cpp`` the eleven-method call surface that the
reference uses (``pytsc/backends/cityflow/simulator.py:50,71-77,88,95``,
``retriever.py:35,95-97,109-111``, ``traffic_signal.py:31,58``) so that the
*unmodified* reference stack can run on it:

    from pytsc_b200 import compat
    from oracle.engine import Engine
    compat.install_stubs(engine_factory=Engine)   # becomes cityflow.Engine

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs may
import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcityflow_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with g++ (Makefile in this directory)."""
    src = os.path.join(_HERE, "cityflow_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, ci, cd, cc = C.c_void_p, C.c_int, C.c_double, C.c_char_p
        pi, pd = C.POINTER(C.c_int), C.POINTER(C.c_double)
        sig = {
            "cfo_last_error": (cc, []),
            "cfo_create": (vp, [cc, ci]),
            "cfo_destroy": (None, [vp]),
            "cfo_next_step": (None, [vp]),
            "cfo_next_steps": (None, [vp, ci]),
            "cfo_get_current_time": (cd, [vp]),
            "cfo_reset": (None, [vp, ci]),
            "cfo_set_tl_phase": (ci, [vp, cc, ci]),
            "cfo_set_tl_phase_idx": (ci, [vp, ci, ci]),
            "cfo_get_vehicle_count": (ci, [vp]),
            "cfo_get_finished_count": (ci, [vp]),
            "cfo_get_created_count": (ci, [vp]),
            "cfo_get_non_fifo_events": (C.c_longlong, [vp]),
            "cfo_get_average_travel_time": (cd, [vp]),
            "cfo_num_lanes": (ci, [vp]),
            "cfo_num_lanelinks": (ci, [vp]),
            "cfo_num_intersections": (ci, [vp]),
            "cfo_lane_id": (cc, [vp, ci]),
            "cfo_intersection_id": (cc, [vp, ci]),
            "cfo_drivable_length": (cd, [vp, ci]),
            "cfo_num_crosses": (ci, [vp, ci]),
            "cfo_get_crosses": (ci, [vp, ci, pi, pi, pd, pd]),
            "cfo_lane_counts": (None, [vp, pi, pi]),
            "cfo_lane_vehicles": (ci, [vp, pi, ci]),
            "cfo_running_vehicles": (ci, [vp, pi, pi, pd, pd, pi, pd, pi, pi, pi, ci]),
            "cfo_vehicle_name": (cc, [vp, ci]),
            "cfo_vehicle_info": (ci, [vp, ci, pd, pd, pi]),
            "cfo_waiting_buffer_sizes": (None, [vp, pi]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Engine:
    """Drop-in for ``cityflow.Engine(config_file=..., thread_num=...)``."""

    def __init__(self, config_file, thread_num=1):
        self._L = lib()
        self._h = self._L.cfo_create(str(config_file).encode(), int(thread_num))
        if not self._h:
            raise RuntimeError("oracle engine: " + self._L.cfo_last_error().decode())
        L = self._L
        self.n_lanes = L.cfo_num_lanes(self._h)
        self.n_lanelinks = L.cfo_num_lanelinks(self._h)
        self.lane_ids = [L.cfo_lane_id(self._h, i).decode() for i in range(self.n_lanes)]
        self.intersection_ids = [L.cfo_intersection_id(self._h, i).decode()
                                 for i in range(L.cfo_num_intersections(self._h))]
        self._sorted_lane_order = sorted(range(self.n_lanes), key=lambda i: self.lane_ids[i])
        self._names = {}
        self._uid_of = {}
        self._cap = 1 << 15

    def __del__(self):
        try:
            if self._h:
                self._L.cfo_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- the eleven methods pytsc calls -------------------------------------
    def next_step(self):
        self._L.cfo_next_step(self._h)

    def get_current_time(self):
        return self._L.cfo_get_current_time(self._h)

    def reset(self, seed=False):
        self._L.cfo_reset(self._h, int(bool(seed)))
        self._names.clear()
        self._uid_of.clear()

    def set_tl_phase(self, intersection_id, phase_id):
        if self._L.cfo_set_tl_phase(self._h, intersection_id.encode(), int(phase_id)) != 0:
            raise RuntimeError(f"set_tl_phase({intersection_id!r}, {phase_id}) rejected")

    def get_vehicle_count(self):
        return self._L.cfo_get_vehicle_count(self._h)

    def get_average_travel_time(self):
        return self._L.cfo_get_average_travel_time(self._h)

    def get_lane_waiting_vehicle_count(self):
        nv = np.empty(self.n_lanes, np.int32)
        nw = np.empty(self.n_lanes, np.int32)
        self._L.cfo_lane_counts(self._h, _ip(nv), _ip(nw))
        return {self.lane_ids[i]: int(nw[i]) for i in self._sorted_lane_order}

    def get_lane_vehicle_count(self):
        nv = np.empty(self.n_lanes, np.int32)
        nw = np.empty(self.n_lanes, np.int32)
        self._L.cfo_lane_counts(self._h, _ip(nv), _ip(nw))
        return {self.lane_ids[i]: int(nv[i]) for i in self._sorted_lane_order}

    def _name(self, uid):
        n = self._names.get(uid)
        if n is None:
            n = self._L.cfo_vehicle_name(self._h, int(uid)).decode()
            self._names[uid] = n
            self._uid_of[n] = uid
        return n

    def get_lane_vehicles(self):
        nv = np.empty(self.n_lanes, np.int32)
        nw = np.empty(self.n_lanes, np.int32)
        self._L.cfo_lane_counts(self._h, _ip(nv), _ip(nw))
        total = int(nv.sum())
        uids = np.empty(max(total, 1), np.int32)
        self._L.cfo_lane_vehicles(self._h, _ip(uids), len(uids))
        off = np.concatenate([[0], np.cumsum(nv)])
        return {self.lane_ids[i]: [self._name(int(u)) for u in uids[off[i]:off[i + 1]]]
                for i in self._sorted_lane_order}

    def snapshot(self):
        """All running vehicles, drivable-major in list order (numpy arrays)."""
        while True:
            cap = self._cap
            a = {k: np.empty(cap, np.int32) for k in
                 ("uid", "drivable", "leader", "blocker", "enter_ll_time", "priority")}
            a["distance"] = np.empty(cap, np.float64)
            a["speed"] = np.empty(cap, np.float64)
            a["gap"] = np.empty(cap, np.float64)
            n = self._L.cfo_running_vehicles(
                self._h, _ip(a["uid"]), _ip(a["drivable"]), _dp(a["distance"]), _dp(a["speed"]),
                _ip(a["leader"]), _dp(a["gap"]), _ip(a["blocker"]), _ip(a["enter_ll_time"]),
                _ip(a["priority"]), cap)
            if n <= cap:
                return {k: v[:n].copy() for k, v in a.items()}
            self._cap = 2 * n

    def get_vehicle_speed(self):
        s = self.snapshot()
        return {self._name(int(u)): float(v) for u, v in zip(s["uid"], s["speed"])}

    def get_vehicle_info(self, vehicle_id):
        uid = self._uid_of.get(vehicle_id)
        if uid is None:  # name not seen through get_lane_vehicles yet: "flow_<i>_<c>"
            for u in range(self._L.cfo_get_created_count(self._h)):
                if self._name(u) == vehicle_id:
                    uid = u
                    break
        if uid is None:
            raise RuntimeError(f"Vehicle '{vehicle_id}' not found")
        d, s, dr = C.c_double(), C.c_double(), C.c_int()
        r = self._L.cfo_vehicle_info(self._h, uid, C.byref(d), C.byref(s), C.byref(dr))
        if r < 0:
            raise RuntimeError(f"Vehicle '{vehicle_id}' not found")
        if r == 0:
            return {"running": "0"}
        # CityFlow returns std::to_string(double) == "%f"
        info = {"running": "1", "distance": "%f" % d.value, "speed": "%f" % s.value}
        if dr.value < self.n_lanes:
            lane = self.lane_ids[dr.value]
            info["drivable"] = lane
            info["road"] = lane.rsplit("_", 1)[0]
        else:
            info["drivable"] = f"lanelink_{dr.value - self.n_lanes}"
        return info

    # ---- extras used by tests -------------------------------------------------
    def next_steps(self, n):
        self._L.cfo_next_steps(self._h, int(n))

    def set_tl_phase_idx(self, inter_index, phase):
        if self._L.cfo_set_tl_phase_idx(self._h, int(inter_index), int(phase)) != 0:
            raise RuntimeError(f"set_tl_phase_idx({inter_index}, {phase}) rejected")

    def get_finished_vehicle_count(self):
        return self._L.cfo_get_finished_count(self._h)

    def get_created_vehicle_count(self):
        return self._L.cfo_get_created_count(self._h)

    def non_fifo_events(self):
        return self._L.cfo_get_non_fifo_events(self._h)

    def drivable_lengths(self):
        n = self.n_lanes + self.n_lanelinks
        return np.array([self._L.cfo_drivable_length(self._h, i) for i in range(n)])

    def crosses(self, inter_index):
        n = self._L.cfo_num_crosses(self._h, inter_index)
        ll0 = np.empty(max(n, 1), np.int32)
        ll1 = np.empty(max(n, 1), np.int32)
        d0 = np.empty(max(n, 1), np.float64)
        d1 = np.empty(max(n, 1), np.float64)
        self._L.cfo_get_crosses(self._h, inter_index, _ip(ll0), _ip(ll1), _dp(d0), _dp(d1))
        return ll0[:n], ll1[:n], d0[:n], d1[:n]

    def waiting_buffer_sizes(self):
        out = np.empty(self.n_lanes, np.int32)
        self._L.cfo_waiting_buffer_sizes(self._h, _ip(out))
        return out
