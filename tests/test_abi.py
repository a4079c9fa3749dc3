"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, and exports
every function include/tsc_b200.h declares; struct layouts agree with ctypes."""
import ctypes as C
import os
import re
import subprocess

import pytest

from helpers import ROOT, build_scenario

HEADER = os.path.join(ROOT, "include", "tsc_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tsc_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(cuda_lib):
    from pytsc_b200 import binding
    names = declared_functions()
    assert len(names) >= 18
    assert set(names) == set(binding.SYMBOLS)
    for n in names:
        assert getattr(cuda_lib, n) is not None
    assert cuda_lib.tsc_abi_version() == 3


def test_library_is_sm100a_sass(cuda_lib):
    from pytsc_b200 import _build
    out = subprocess.run(["cuobjdump", "-lelf", _build.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_struct_layout_matches_header(tmp_path):
    """sizeof / offsetof of the two ABI structs, as gcc sees the header vs ctypes."""
    from pytsc_b200.binding import tsc_outputs_t
    from pytsc_b200.scenario import tsc_scenario_t
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "tsc_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(tsc_scenario_t), offsetof(tsc_scenario_t, tmpl), offsetof(tsc_scenario_t, interval),'
                   'sizeof(tsc_outputs_t), offsetof(tsc_outputs_t, metrics), offsetof(tsc_scenario_t, reward_type), offsetof(tsc_scenario_t, ctl_off));return 0;}')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    exp = [C.sizeof(tsc_scenario_t), tsc_scenario_t.tmpl.offset, tsc_scenario_t.interval.offset,
           C.sizeof(tsc_outputs_t), tsc_outputs_t.metrics.offset, tsc_scenario_t.reward_type.offset, tsc_scenario_t.ctl_off.offset]
    assert got == exp


def test_create_fails_loudly_without_gpu(cuda_lib):
    """No CPU fallback: on a machine without CUDA, tsc_create returns an error and
    the Python Engine refuses to construct."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from pytsc_b200.binding import Engine
    cfg, parser, cs = build_scenario("syn_1x1")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(cs, 1)
    s = cs.to_struct()
    h = C.c_void_p()
    rc = cuda_lib.tsc_create(C.byref(s), 1, 0, 0, C.byref(h))
    assert rc < 0 and cuda_lib.tsc_last_error()


def test_null_and_bad_arguments_return_einval(cuda_lib):
    """Error behaviour of the ABI without touching a device: every entry point rejects a null handle
    (or a null required pointer) with TSC_EINVAL and leaves a message; nothing crashes."""
    L = cuda_lib
    EINVAL = -1
    null = C.c_void_p()
    out = C.c_void_p()
    i = C.c_int32()
    assert L.tsc_create(None, 1, 0, 0, C.byref(out)) == EINVAL and b"null" in L.tsc_last_error()
    cfg, parser, cs = build_scenario("syn_1x1")
    s = cs.to_struct()
    assert L.tsc_create(C.byref(s), 0, 0, 0, C.byref(out)) == EINVAL and b"n_replicas" in L.tsc_last_error()
    s.abi_version = 1
    assert L.tsc_create(C.byref(s), 1, 0, 0, C.byref(out)) == EINVAL and b"abi_version" in L.tsc_last_error()
    s = cs.to_struct()
    s.interval = 0.5
    assert L.tsc_create(C.byref(s), 1, 0, 0, C.byref(out)) == EINVAL and b"interval" in L.tsc_last_error()
    for call in (lambda: L.tsc_reset(null, None), lambda: L.tsc_step(null, 1, None), lambda: L.tsc_set_phase(null, None, None),
                 lambda: L.tsc_init_program(null, 0, None), lambda: L.tsc_retrieve(null, None, None),
                 lambda: L.tsc_env_step(null, None, 0, 0, 5, None, None),
                 lambda: L.tsc_env_step_host(null, None, 0, 0, 5, None, None, None, None),
                 lambda: L.tsc_snapshot(null, 0, 0, None, None, None, None, None, None),
                 lambda: L.tsc_load_snapshot(null, 0, 0, None, None, None, None, None),
                 lambda: L.tsc_check(null, C.byref(i)), lambda: L.tsc_counters(null, None, None, None, None),
                 lambda: L.tsc_kernel_info(null, None, None, None, None), lambda: L.tsc_kernel_variant(null, None, None, None),
                 lambda: L.tsc_debug_timing(null, 0, None, 0), lambda: L.tsc_controller_act(null, 3, 0, None, None, None),
                 lambda: L.tsc_reset_replicas(null, None, 0, None), lambda: L.tsc_save_state(null, None, 0, None),
                 lambda: L.tsc_load_state(null, None, 0, None),
                 lambda: L.tsc_get_dims(null, None, None, None, None, None, None, None, None, None)):
        assert call() == EINVAL
        assert L.tsc_last_error()
    assert L.tsc_launch_count(null) == 0 and L.tsc_state_bytes(null) == 0
    L.tsc_destroy(null)      # no-op
