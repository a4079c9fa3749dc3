"""Import shims so that the *unmodified* reference package ``pytsc`` can be
imported on a machine that has neither CityFlow nor SUMO installed.

``pytsc/__init__.py:3-4`` eagerly imports both simulator backends, and those
modules import ``cityflow`` (``backends/cityflow/simulator.py:1``), ``traci``,
``traci.constants``, ``sumolib`` (``backends/sumo/*.py``) and even
``from turtle import pos`` (``backends/sumo/traffic_signal.py:1``); the SUMO
modules ``sys.exit`` when ``SUMO_HOME`` is unset
(``backends/sumo/simulator.py:7-14``).  None of that is needed by the ``gpu``
backend, so we install inert stand-ins -- **only** for modules that are really
missing; a real CityFlow / SUMO install is never shadowed.

Nothing in this file is on the simulation path.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import tempfile
import types

_TRACI_CONSTANTS = (
    "LAST_STEP_VEHICLE_NUMBER", "LAST_STEP_VEHICLE_HALTING_NUMBER",
    "LAST_STEP_MEAN_SPEED", "LAST_STEP_OCCUPANCY", "LAST_STEP_VEHICLE_ID_LIST",
    "VAR_WAITING_TIME", "VAR_ACCUMULATED_WAITING_TIME", "VAR_SPEED",
    "VAR_LANEPOSITION", "VAR_LANE_ID", "VAR_TIME", "VAR_TIME_STEP",
    "VAR_DEPARTED_VEHICLES_IDS", "VAR_ARRIVED_VEHICLES_IDS",
    "VAR_DEPARTED_VEHICLES_NUMBER", "VAR_ARRIVED_VEHICLES_NUMBER",
    "VAR_COLLIDING_VEHICLES_NUMBER", "VAR_MIN_EXPECTED_VEHICLES",
    "VAR_PENDING_VEHICLES", "VAR_LENGTH", "VAR_MAXSPEED", "VAR_POSITION",
    "LAST_STEP_TIME_SINCE_DETECTION", "VAR_ALLOWED_SPEED", "VAR_DISTANCE",
)


class _Anything(types.ModuleType):
    """A module whose every missing attribute resolves to an inert integer /
    callable, enough for ``import``-time references in the SUMO backend."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name.isupper():
            return hash(name) & 0xFF
        def _missing(*a, **k):
            raise RuntimeError(
                f"{self.__name__}.{name} is a stand-in: the real package is not installed")
        return _missing


def _have(mod: str) -> bool:
    if mod in sys.modules:
        return True
    try:
        return importlib.util.find_spec(mod) is not None
    except (ImportError, ValueError):
        return False


def _stub(name: str, **attrs) -> types.ModuleType:
    m = _Anything(name)
    m.__dict__.update(attrs)
    m.__pytsc_b200_stub__ = True
    sys.modules[name] = m
    return m


def install_stubs(engine_factory=None) -> dict:
    """Install stand-ins for whatever of cityflow / traci / sumolib / turtle /
    smac / matplotlib is missing.  ``engine_factory`` (optional) becomes
    ``cityflow.Engine`` when CityFlow itself is absent -- tests pass the CPU
    oracle's Engine adapter here; the product never does.

    Returns a dict ``{module: "real" | "stub"}`` for reporting."""
    report = {}
    if _have("cityflow") and not getattr(sys.modules.get("cityflow"), "__pytsc_b200_stub__", False):
        report["cityflow"] = "real"
    else:
        def _no_engine(*a, **k):
            raise RuntimeError("CityFlow is not installed; use the `gpu` backend")
        m = sys.modules.get("cityflow") or _stub("cityflow")
        m.Engine = engine_factory or getattr(m, "Engine", None) or _no_engine
        report["cityflow"] = "stub"
    if "SUMO_HOME" not in os.environ:
        os.environ["SUMO_HOME"] = tempfile.gettempdir()
    if _have("traci") and not getattr(sys.modules.get("traci"), "__pytsc_b200_stub__", False):
        report["traci"] = "real"
    else:
        tc = _stub("traci.constants", **{k: i for i, k in enumerate(_TRACI_CONSTANTS)})
        t = _stub("traci", constants=tc)
        t.__path__ = []
        report["traci"] = "stub"
    if _have("sumolib") and not getattr(sys.modules.get("sumolib"), "__pytsc_b200_stub__", False):
        report["sumolib"] = "real"
    else:
        s = _stub("sumolib", checkBinary=lambda b: b)
        s.__path__ = []
        s.net = _stub("sumolib.net", readNet=lambda *a, **k: None)
        s.miscutils = _stub("sumolib.miscutils", getFreeSocketPort=lambda: 0)
        report["sumolib"] = "stub"
    try:  # `from turtle import pos` needs tkinter
        importlib.import_module("turtle")
        report["turtle"] = "real"
    except Exception:
        _stub("turtle", pos=None)
        report["turtle"] = "stub"
    if not _have("smac"):
        class MultiAgentEnv:  # minimal smac base class
            def __init__(self, *a, **k):
                pass
        s = _stub("smac")
        s.__path__ = []
        s.env = _stub("smac.env", MultiAgentEnv=MultiAgentEnv)
        report["smac"] = "stub"
    if not _have("matplotlib"):
        m = _stub("matplotlib")
        m.__path__ = []
        m.pyplot = _stub("matplotlib.pyplot")
        report["matplotlib"] = "stub"
    return report


def find_reference_pytsc():
    """Locate an importable reference ``pytsc`` (never required by the product):
    an installed package, ``baseline/_ref`` in this repo, or ``/root/reference``
    (present only in the build container).  Returns the path added to
    ``sys.path`` or ``None``."""
    if _have("pytsc"):
        return "installed"
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("PYTSC_REFERENCE"),
                 os.path.join(here, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "pytsc")):
            sys.path.insert(0, cand)
            return cand
    return None
