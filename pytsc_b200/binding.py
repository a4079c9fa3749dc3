"""ctypes binding of ``libtsc_b200.so`` (include/tsc_b200.h).

Thin on purpose: torch owns device buffers and streams, this module only passes
``data_ptr()`` values across the C ABI.  There is no CPU fallback -- if the
library is missing or no CUDA device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .scenario import CompiledScenario, tsc_scenario_t

_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtsc_b200.so")

ERRORS = {-1: "TSC_EINVAL", -2: "TSC_ECUDA", -3: "TSC_ENOMEM", -4: "TSC_EOVERFLOW", -5: "TSC_EORDER"}

# every symbol include/tsc_b200.h declares
SYMBOLS = ("tsc_abi_version", "tsc_last_error", "tsc_create", "tsc_destroy", "tsc_get_dims", "tsc_reset",
           "tsc_set_phase", "tsc_init_program", "tsc_step", "tsc_retrieve", "tsc_env_step", "tsc_env_step_host",
           "tsc_snapshot", "tsc_load_snapshot", "tsc_check", "tsc_counters", "tsc_launch_count", "tsc_kernel_info",
           "tsc_debug_timing", "tsc_controller_act", "tsc_kernel_variant", "tsc_reset_replicas", "tsc_state_bytes",
           "tsc_save_state", "tsc_load_state", "tsc_reset_flows", "tsc_reset_replicas_flows", "tsc_host_register",
           "tsc_host_unregister", "tsc_env_step_registered", "tsc_host_packet_bytes", "tsc_max_spanning_tree",
           "tsc_env_step_registered_begin", "tsc_env_step_registered_wait", "tsc_host_threads")

# tsc_env_step / tsc_controller_act controller codes (include/tsc_b200.h)
CONTROLLERS = {"external": 0, "fixed_time": 1, "phase_index": 2, "greedy": 3, "max_pressure": 4, "sotl": 5, "random": 6}
SCORE_MASKED = -2 ** 31


def sotl_arg(theta=3, mu=4, phi_min=5):
    """TSC_SOTL_ARG: the SOTLController parameters (controllers/controllers.py:190-197) in one int32."""
    return (theta & 0xFF) | ((mu & 0xFF) << 8) | ((phi_min & 0xFFFF) << 16)


class tsc_outputs_t(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "lane_count", "lane_queued", "lane_occupancy", "lane_mean_speed", "lane_meas64", "pos_in", "pos_out",
        "sig_stats64", "obs", "state", "reward", "reward_global", "mask", "sim", "metrics", "density_map", "err")]


class TscError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


def load_library(path=None):
    """Load the CUDA library; raises if it has not been built (see ``_build.py``)."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    path = path or os.environ.get("TSC_B200_LIB") or LIB_PATH      # TSC_B200_LIB: another build of the same ABI (A/B measurements)
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -m pytsc_b200._build` "
                           "(the gpu backend has no CPU fallback)")
    L = C.CDLL(path)
    vp, i32, pi32 = C.c_void_p, C.c_int32, C.POINTER(C.c_int32)
    L.tsc_abi_version.restype = C.c_int
    L.tsc_last_error.restype = C.c_char_p
    L.tsc_create.argtypes = [C.POINTER(tsc_scenario_t), i32, i32, i32, C.POINTER(vp)]
    L.tsc_destroy.argtypes = [vp]
    L.tsc_destroy.restype = None
    L.tsc_get_dims.argtypes = [vp] + [pi32] * 9
    L.tsc_reset.argtypes = [vp, vp]
    L.tsc_reset_replicas.argtypes = [vp, vp, i32, vp]
    L.tsc_reset_flows.argtypes = [vp, vp, vp]
    L.tsc_reset_replicas_flows.argtypes = [vp, vp, vp, i32, vp]
    L.tsc_host_register.argtypes = [vp, vp, vp, vp, vp]
    L.tsc_host_unregister.argtypes = [vp]
    L.tsc_env_step_registered.argtypes = [vp, vp, i32, i32, i32]
    L.tsc_env_step_registered_begin.argtypes = [vp, vp, i32, i32, i32]
    L.tsc_env_step_registered_wait.argtypes = [vp]
    L.tsc_host_threads.argtypes = [vp, i32]
    L.tsc_host_packet_bytes.argtypes = [vp]
    L.tsc_host_packet_bytes.restype = C.c_int64
    L.tsc_max_spanning_tree.argtypes = [vp, vp, vp, vp]
    L.tsc_state_bytes.argtypes = [vp]
    L.tsc_state_bytes.restype = C.c_int64
    L.tsc_save_state.argtypes = [vp, vp, C.c_int64, vp]
    L.tsc_load_state.argtypes = [vp, vp, C.c_int64, vp]
    L.tsc_set_phase.argtypes = [vp, vp, vp]
    L.tsc_init_program.argtypes = [vp, i32, vp]
    L.tsc_step.argtypes = [vp, i32, vp]
    L.tsc_retrieve.argtypes = [vp, C.POINTER(tsc_outputs_t), vp]
    L.tsc_env_step.argtypes = [vp, vp, i32, i32, i32, C.POINTER(tsc_outputs_t), vp]
    L.tsc_env_step_host.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp]
    L.tsc_snapshot.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp]
    L.tsc_load_snapshot.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    L.tsc_check.argtypes = [vp, pi32]
    L.tsc_counters.argtypes = [vp, vp, vp, vp, vp]
    L.tsc_launch_count.argtypes = [vp]
    L.tsc_launch_count.restype = C.c_int64
    L.tsc_kernel_info.argtypes = [vp, pi32, pi32, pi32, pi32]
    L.tsc_debug_timing.argtypes = [vp, i32, vp, i32]
    L.tsc_controller_act.argtypes = [vp, i32, i32, vp, vp, vp]
    L.tsc_kernel_variant.argtypes = [vp, pi32, pi32, pi32]
    for n in SYMBOLS:
        getattr(L, n)
    if path == LIB_PATH:
        _LIB = L
    return L


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _np_ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


OUTPUT_SPECS = {   # name -> (shape builder, torch dtype name)
    "lane_count": (lambda d: (d["B"], d["L"]), "int32"),
    "lane_queued": (lambda d: (d["B"], d["L"]), "int32"),
    "lane_occupancy": (lambda d: (d["B"], d["L"]), "float32"),
    "lane_mean_speed": (lambda d: (d["B"], d["L"]), "float32"),
    "lane_meas64": (lambda d: (d["B"], d["L"], 2), "float64"),
    "pos_in": (lambda d: (d["B"], d["n_in"], d["vis"]), "float32"),
    "pos_out": (lambda d: (d["B"], d["n_out"], d["vis"]), "float32"),
    "sig_stats64": (lambda d: (d["B"], d["A"], 8), "float64"),
    "obs": (lambda d: (d["B"], d["A"], d["obs_dim"]), "float32"),
    "state": (lambda d: (d["B"], d["A"], d["state_dim"]), "float32"),
    "reward": (lambda d: (d["B"], d["A"]), "float32"),
    "reward_global": (lambda d: (d["B"],), "float32"),
    "mask": (lambda d: (d["B"], d["A"], d["n_actions"]), "uint8"),
    "sim": (lambda d: (d["B"], 4), "float64"),
    "metrics": (lambda d: (d["B"], 8), "float64"),
    "density_map": (lambda d: (d["B"], d["A"], d["A"]), "float64"),
    "err": (lambda d: (d["B"],), "int32"),
}


class Engine:
    """B replicas of one compiled scenario on one CUDA device."""

    def __init__(self, scenario: CompiledScenario, n_replicas: int, device: int = 0, vehicle_capacity: int = 0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("pytsc_b200: no CUDA device -- the gpu backend has no CPU fallback")
        self.torch = torch
        self.lib = load_library()
        self.scenario = scenario
        self.device = torch.device("cuda", device)
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.zeros(1, device=self.device)   # make sure the primary context exists
        self._struct = scenario.to_struct()
        h = C.c_void_p()
        self._check(self.lib.tsc_create(C.byref(self._struct), n_replicas, device, vehicle_capacity, C.byref(h)))
        self.h = h
        v = [C.c_int32() for _ in range(9)]
        self._check(self.lib.tsc_get_dims(self.h, *[C.byref(x) for x in v]))
        self.dims = dict(zip(("B", "L", "A", "obs_dim", "state_dim", "n_actions", "n_in", "n_out", "vis"),
                             [x.value for x in v]))
        self.B, self.L, self.A = self.dims["B"], self.dims["L"], self.dims["A"]

    def _check(self, rc):
        if rc != 0:
            raise TscError(rc, self.lib.tsc_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.tsc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- buffers ---------------------------------------------------------------------
    def alloc_outputs(self, names=None):
        torch = self.torch
        names = names or list(OUTPUT_SPECS)
        return {n: torch.zeros(OUTPUT_SPECS[n][0](self.dims), dtype=getattr(torch, OUTPUT_SPECS[n][1]),
                               device=self.device) for n in names}

    def _outputs(self, bufs):
        o = tsc_outputs_t()
        for n, t in (bufs or {}).items():
            exp = OUTPUT_SPECS[n]
            assert tuple(t.shape) == tuple(exp[0](self.dims)) and t.is_contiguous() and t.is_cuda, n
            assert str(t.dtype) == "torch." + exp[1], (n, t.dtype)
            setattr(o, n, t.data_ptr())
        return o

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    # ---- engine-level calls --------------------------------------------------------------
    def reset(self, flow_sets=None):
        """All replicas back to tick 0; ``flow_sets`` (int array [B]) = the flow set every replica restarts on
        (``tsc_reset_flows``; None keeps the current assignment)."""
        if flow_sets is None:
            self._check(self.lib.tsc_reset(self.h, self._stream()))
        else:
            fs = np.ascontiguousarray(np.asarray(flow_sets, np.int32).reshape(-1))
            assert fs.size == self.B, "one flow set per replica"
            self._check(self.lib.tsc_reset_flows(self.h, _np_ptr(fs), self._stream()))

    def reset_replicas(self, replicas, flow_sets=None):
        """Selected replicas back to tick 0 (host list / array of replica indices), optionally on new flow sets."""
        idx = np.ascontiguousarray(np.asarray(replicas, np.int32).reshape(-1))
        fs = None if flow_sets is None else np.ascontiguousarray(np.asarray(flow_sets, np.int32).reshape(-1))
        assert fs is None or fs.size == idx.size
        self._check(self.lib.tsc_reset_replicas_flows(self.h, _np_ptr(idx), _np_ptr(fs), len(idx), self._stream()))

    def save_state(self, device=False):
        """Snapshot of all replicas (CityFlow's engine.snapshot()): a uint8 tensor, pinned host or device."""
        n = int(self.lib.tsc_state_bytes(self.h))
        buf = (self.torch.empty(n, dtype=self.torch.uint8, device="cuda") if device
               else self.torch.empty(n, dtype=self.torch.uint8, pin_memory=True))
        self._check(self.lib.tsc_save_state(self.h, _ptr(buf), n, self._stream()))
        return buf

    def load_state(self, blob):
        """Restore a snapshot taken by ``save_state`` (CityFlow's engine.load(archive))."""
        assert blob.dtype == self.torch.uint8 and blob.is_contiguous()
        self._check(self.lib.tsc_load_state(self.h, _ptr(blob), blob.numel(), self._stream()))

    def set_phase(self, raw_phase):
        assert raw_phase.dtype == self.torch.int32 and tuple(raw_phase.shape) == (self.B, self.A)
        self._check(self.lib.tsc_set_phase(self.h, _ptr(raw_phase.contiguous()), self._stream()))

    def init_program(self, phase_index=0):
        self._check(self.lib.tsc_init_program(self.h, phase_index, self._stream()))

    def step(self, n_ticks=1):
        self._check(self.lib.tsc_step(self.h, n_ticks, self._stream()))

    def retrieve(self, bufs):
        o = self._outputs(bufs)
        self._check(self.lib.tsc_retrieve(self.h, C.byref(o), self._stream()))

    def env_step(self, actions, bufs, n_ticks=5, controller=0, controller_arg=0):
        if actions is not None:
            assert actions.dtype == self.torch.int32 and tuple(actions.shape) == (self.B, self.A) and actions.is_contiguous()
        o = self._outputs(bufs)
        self._check(self.lib.tsc_env_step(self.h, _ptr(actions), controller, controller_arg, n_ticks,
                                          C.byref(o) if bufs is not None else None, self._stream()))

    def controller_act(self, controller, controller_arg=0, scores=False):
        """What pytsc's rule-based ``controller`` would do now: int32 [B, A] phase indices (and the
        [B, A, P] scores behind them), without touching the programs."""
        torch = self.torch
        acts = torch.empty((self.B, self.A), dtype=torch.int32, device=self.device)
        sc = torch.empty((self.B, self.A, self.scenario.max_phases), dtype=torch.int32, device=self.device) if scores else None
        self._check(self.lib.tsc_controller_act(self.h, CONTROLLERS.get(controller, controller), controller_arg,
                                                _ptr(acts), _ptr(sc), self._stream()))
        return (acts, sc) if scores else acts

    def env_step_host(self, actions, obs=None, reward=None, mask=None, reward_global=None, n_ticks=5,
                      controller=0, controller_arg=0):
        """numpy in, numpy out (host buffers), synchronous: the end-to-end path."""
        self._check(self.lib.tsc_env_step_host(self.h, _np_ptr(actions), controller, controller_arg, n_ticks,
                                               _np_ptr(obs), _np_ptr(reward), _np_ptr(mask), _np_ptr(reward_global)))

    def host_register(self, obs=None, reward=None, mask=None, reward_global=None, threads=0):
        """Register the caller's HOST result arrays (numpy, C-contiguous) for ``env_step_registered``: one launch
        per step, compact per-replica packets over PCIe, rows finished by host threads (``tsc_host_register``;
        ``threads`` = 0 lets the library choose).
        The arrays must stay alive and must not be written by the caller until ``host_unregister``."""
        n = self.lib.tsc_host_threads(self.h, int(threads))      # (returns the worker count that will be used)
        if n < 0:
            self._check(n)
        d = self.dims
        for arr, shape, dt in ((obs, (d["B"], d["A"], d["obs_dim"]), np.float32), (reward, (d["B"], d["A"]), np.float32),
                               (mask, (d["B"], d["A"], d["n_actions"]), np.uint8), (reward_global, (d["B"],), np.float32)):
            assert arr is None or (arr.shape == shape and arr.dtype == dt and arr.flags["C_CONTIGUOUS"]), (shape, dt)
        self._registered = (obs, reward, mask, reward_global)
        self._check(self.lib.tsc_host_register(self.h, _np_ptr(obs), _np_ptr(reward), _np_ptr(mask), _np_ptr(reward_global)))

    def host_unregister(self):
        self._check(self.lib.tsc_host_unregister(self.h))
        self._registered = None

    def env_step_registered(self, actions=None, n_ticks=5, controller=0, controller_arg=0):
        """numpy int32 [B, A] actions in (None for the in-kernel controllers); the registered arrays hold the results
        on return (synchronous)."""
        if actions is not None:
            assert actions.dtype == np.int32 and actions.shape == (self.B, self.A) and actions.flags["C_CONTIGUOUS"]
        self._check(self.lib.tsc_env_step_registered(self.h, _np_ptr(actions), controller, controller_arg, n_ticks))

    def env_step_registered_begin(self, actions=None, n_ticks=5, controller=0, controller_arg=0):
        """First half of ``env_step_registered``: queue the action copy and the launch, wake the host workers, return.
        ``actions`` must stay untouched until ``env_step_registered_wait``."""
        if actions is not None:
            assert actions.dtype == np.int32 and actions.shape == (self.B, self.A) and actions.flags["C_CONTIGUOUS"]
        self._inflight_actions = actions
        self._check(self.lib.tsc_env_step_registered_begin(self.h, _np_ptr(actions), controller, controller_arg, n_ticks))

    def env_step_registered_wait(self):
        """Second half: on return the registered arrays hold the step's results."""
        self._check(self.lib.tsc_env_step_registered_wait(self.h))
        self._inflight_actions = None

    def max_spanning_tree(self, density_map):
        """``MetricsParser.mst`` for every replica: float64 [B, A, A] device tensor in (as the ``density_map`` output), the
        tree's edges out (-weight in the upper triangle, scipy's sign convention)."""
        torch = self.torch
        assert density_map.dtype == torch.float64 and tuple(density_map.shape) == (self.B, self.A, self.A) and density_map.is_contiguous()
        out = torch.empty_like(density_map)
        self._check(self.lib.tsc_max_spanning_tree(self.h, _ptr(density_map), _ptr(out), self._stream()))
        return out

    def host_packet_bytes(self):
        return int(self.lib.tsc_host_packet_bytes(self.h))

    # ---- debugging / parity ------------------------------------------------------------------
    def snapshot(self, replica=0):
        cap = 1 << 14
        while True:
            vid = np.empty(cap, np.int32); drv = np.empty(cap, np.int32)
            dist = np.empty(cap, np.float64); spd = np.empty(cap, np.float64)
            blk = np.empty(cap, np.int32); ellt = np.empty(cap, np.int32)
            n = self.lib.tsc_snapshot(self.h, replica, cap, _np_ptr(vid), _np_ptr(drv), _np_ptr(dist), _np_ptr(spd),
                                      _np_ptr(blk), _np_ptr(ellt))
            if n < 0:
                self._check(n)
            if n <= cap:
                return dict(uid=vid[:n].copy(), drivable=drv[:n].copy(), distance=dist[:n].copy(),
                            speed=spd[:n].copy(), blocker=blk[:n].copy(), enter_ll_time=ellt[:n].copy())
            cap = n

    def load_snapshot(self, replica, drivable, distance, speed, vid=None, route_pos=None):
        drivable = np.ascontiguousarray(drivable, np.int32)
        distance = np.ascontiguousarray(distance, np.float64)
        speed = np.ascontiguousarray(speed, np.float64)
        vid = None if vid is None else np.ascontiguousarray(vid, np.int32)
        route_pos = None if route_pos is None else np.ascontiguousarray(route_pos, np.int32)
        self._check(self.lib.tsc_load_snapshot(self.h, replica, len(drivable), _np_ptr(vid), _np_ptr(drivable),
                                               _np_ptr(distance), _np_ptr(speed), _np_ptr(route_pos)))

    def check(self):
        bad = C.c_int32(-1)
        self._check(self.lib.tsc_check(self.h, C.byref(bad)))

    def counters(self):
        out = {k: np.empty(self.B, np.int32) for k in ("tick", "n_running", "n_finished", "n_slots")}
        self._check(self.lib.tsc_counters(self.h, *[_np_ptr(out[k]) for k in ("tick", "n_running", "n_finished", "n_slots")]))
        return out

    PHASES = ("stage_in", "prologue", "spawn", "retrieve_lanes", "decisions", "retrieve_signals", "cross", "leave", "enter", "compact",
              "retrieve", "stage_out", "n_heads", "n_zone", "n_cross", "n_movers")

    def debug_timing(self, enable=True):
        """Read the per-phase cycle counters accumulated so far (dict), then enable / disable them."""
        buf = np.zeros(32, np.uint64)
        n = self.lib.tsc_debug_timing(self.h, int(enable), _np_ptr(buf), 32)
        if n < 0:
            self._check(n)
        return dict(zip(self.PHASES, (int(x) for x in buf[:n])))

    def launch_count(self):
        return int(self.lib.tsc_launch_count(self.h))

    def kernel_info(self):
        v = [C.c_int32() for _ in range(4)]
        self._check(self.lib.tsc_kernel_info(self.h, *[C.byref(x) for x in v]))
        w = [C.c_int32() for _ in range(3)]
        self._check(self.lib.tsc_kernel_variant(self.h, *[C.byref(x) for x in w]))
        return dict(zip(("smem_bytes", "threads", "grid", "regs", "fixed_capacity", "global_workspace", "blocks_per_sm"),
                        [x.value for x in v + w]))
