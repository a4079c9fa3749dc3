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
    assert cuda_lib.tsc_abi_version() == 2


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
