"""TEST-ONLY: compile the kernel sources with g++ against tests/emu/cuda_emu.h into
tests/emu/libffb_emu.so so the CPU test-suite can execute the kernels' tiling logic without a GPU.
The product never loads this library."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "funscript_flow_b200", "csrc")
LIB = os.path.join(HERE, "libffb_emu.so")
DEPS = [os.path.join(CSRC, f) for f in ("ffb_api.cu", "ffb_kernels.cuh", "ffb_common.h")] + [
    os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "ffb.h")]


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS):
        return LIB
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-pthread", "-DFFB_EMU", "-mfma", "-ffp-contract=fast",
           "-Wall", "-Wno-unused-function", "-Wno-unused-variable", "-Wno-unknown-pragmas",
           "-I", HERE, "-I", CSRC, "-I", os.path.join(ROOT, "include"),
           "-o", LIB, os.path.join(CSRC, "ffb_api.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emu build failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
