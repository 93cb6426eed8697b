"""CPU suite, part 2: the C-ABI library builds for sm_100a, loads, exports every symbol that
include/ffb.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

from funscript_flow_b200 import _native, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    return build.build()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ffb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ffb_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ffb.h but not exported"
    assert sorted(_native.load(lib_path)._ffb_symbols) == names   # the ctypes binding covers the whole header


def test_library_is_sm100a(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_version_and_level_plan(lib_path):
    lib = _native.load(lib_path)
    assert lib.ffb_version() == 100
    plan = _native.level_plan(1920, 1080, lib_path)
    assert [(p["w"], p["h"], p["ksize"]) for p in plan] == [(240, 135, 19), (480, 270, 9), (960, 540, 3), (1920, 1080, 3)]
    assert [(p["w"], p["h"]) for p in _native.level_plan(517, 389, lib_path)] == [(65, 49), (129, 97), (258, 194), (517, 389)]


def test_no_cpu_fallback(lib_path):
    if _native.device_count(lib_path) > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_native.FFBError) as ei:
        _native.FlowContext(0, lib_path)
    assert ei.value.code == -4 and "no CPU fallback" in str(ei.value)
    import funscript_flow_b200 as ffb
    import numpy as np
    from funscript_flow_b200 import api
    saved = dict(api._contexts)
    api._contexts.clear()          # other test modules may have injected the emulated context
    try:
        with pytest.raises(Exception):
            ffb.max_divergence(np.zeros((32, 32, 2), np.float32))
    finally:
        api._contexts.update(saved)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "funscript_flow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libffb_emu" not in src and "cuda_emu.h" not in src.replace('#include "cuda_emu.h"', ""), f


def test_gpu_info_strings_without_a_gpu():
    """get_gpu_info mirrors F:64-99; on a box without a device it says what the reference says ("CPU only") and
    get_available_backends reports CUDA False -- nothing pretends to be runnable."""
    from funscript_flow_b200 import api
    if _native.device_count() == 0:
        assert api.get_gpu_info() == "CPU only"
        assert api.get_available_backends() == {"CPU": False, "CUDA": False, "OpenCL": False, "DNN": False}
        assert _native.device_name(0) is None and _native.device_pci_bus_id(0) is None
    else:
        assert api.get_gpu_info().startswith("CUDA: ")
