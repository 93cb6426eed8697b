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
    assert lib.ffb_version() == 200
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


def test_header_is_plain_c_and_links(tmp_path):
    """include/ffb.h compiles as C11 with -Wall -Werror -pedantic (no C++ leaking into the boundary) and a C program
    that calls the device-independent entry points links against libffb.so and runs without a GPU."""
    import shutil
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = build.build()
    src = tmp_path / "use_ffb.c"
    src.write_text(r'''
#include <stdio.h>
#include "ffb.h"
int main(void) {
    int n = -1;
    ffb_ctx* ctx = NULL;
    int w[4], h[4], ks[4];
    double sg[4];
    int levels = 0;
    if (ffb_version() != FFB_VERSION) return 1;
    if (ffb_device_count(&n) != FFB_OK || n < 0) return 2;
    if (ffb_level_plan(1920, 1080, &levels, w, h, ks, sg) != FFB_OK || levels != 4 || w[0] != 240 || w[3] != 1920) return 3;
    if (n == 0 && ffb_create(0, &ctx) != FFB_E_NODEVICE) return 4;      /* no CPU fallback */
    if (n == 0 && ffb_last_error(NULL)[0] == 0) return 5;
    if (ctx) ffb_destroy(ctx);
    printf("ok %d devices, %s\n", n, ffb_kernel_name(FFB_K_FLOW_ITER));
    return 0;
}
''')
    exe = tmp_path / "use_ffb"
    cmd = [gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
           lib, "-Wl,-rpath," + os.path.dirname(lib)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and run.stdout.startswith("ok "), (run.returncode, run.stdout, run.stderr)
