"""pytsc_b200 -- a B200-native `gpu` simulator backend for rbokade/pytsc.

Two ways in:

* ``register()`` adds the backend to an importable pytsc, after which
  ``pytsc.TrafficSignalNetwork(scenario, "gpu", ...)``, ``run_controllers``,
  ``Evaluate`` and the MARL wrappers work with it unchanged;
* ``BatchedTrafficSignalNetwork`` steps B replicas at once with device tensors
  in and out (one CUDA launch per env-step).

Both drive the same C-ABI library (``include/tsc_b200.h``,
``pytsc_b200/csrc/tsc_b200.cu``).  There is no CPU fallback.
"""
from __future__ import annotations

__version__ = "0.1.0"

BACKEND_NAME = "gpu"


def register(install_import_stubs: bool = True):
    """Make ``"gpu"`` a pytsc simulator backend.

    pytsc looks its plugin classes up in ``pytsc.SIMULATOR_MODULES`` and checks
    the name against ``pytsc.SUPPORTED_SIMULATOR_BACKENDS`` at call time
    (``pytsc/__init__.py:9-14, 23-40``).  pytsc eagerly imports its CityFlow and
    SUMO backends; when those simulators are not installed, inert stand-ins are
    installed first (``compat.install_stubs``) so that the import succeeds."""
    from . import compat
    if install_import_stubs:
        compat.install_stubs()
    if compat.find_reference_pytsc() is None:
        raise ImportError("pytsc is not importable; install rbokade/pytsc or set PYTSC_REFERENCE")
    import pytsc
    from .backend import GPU_MODULES
    pytsc.SIMULATOR_MODULES[BACKEND_NAME] = GPU_MODULES
    if BACKEND_NAME not in pytsc.SUPPORTED_SIMULATOR_BACKENDS:
        pytsc.SUPPORTED_SIMULATOR_BACKENDS = tuple(pytsc.SUPPORTED_SIMULATOR_BACKENDS) + (BACKEND_NAME,)
    return pytsc


def __getattr__(name):
    if name in ("BatchedTrafficSignalNetwork", "BatchedEPyMARLTrafficSignalNetwork"):
        from . import env
        return getattr(env, name)
    raise AttributeError(name)
