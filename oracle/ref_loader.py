"""CPU ORACLE helper (test infrastructure) -- load the reference's *own* hot-path functions out of
``FunscriptFlow.pyw`` without importing it (the file pulls in PySide6/matplotlib at import time).

The .pyw is parsed with ``ast``; only whitelisted top-level definitions are kept and executed
in a namespace seeded with the stdlib / NumPy / cv2 names they use.  Nothing is copied into the
repo's history: the source is read where it lies (``/root/reference``, which exists only in the build
container) or from the git-ignored staging copy ``baseline/_ref/FunscriptFlow.pyw`` that travels to the GPU
box for the reference arm of bench.py.  What the tests need from it is stored as golden vectors under
``tests/golden/`` by ``tests/golden/make_golden.py``.
"""
from __future__ import annotations

import ast
import os
import sys
import types

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# Where the unmodified FunscriptFlow.pyw may lie: the read-only reference checkout of the build container, or the
# git-ignored staging copy baseline/_ref/ that __graft_entry__.build() makes from it so that the file travels to
# the GPU box with the repo snapshot (it is never committed: .gitignore lists baseline/_ref/).
STAGED_PYW = os.path.join(_ROOT, "baseline", "_ref", "FunscriptFlow.pyw")
_CANDIDATES = [os.environ.get("FFB_REFERENCE_PYW", ""), "/root/reference/FunscriptFlow.pyw", STAGED_PYW]
REFERENCE_PYW = next((p for p in _CANDIDATES if p and os.path.isfile(p)), "/root/reference/FunscriptFlow.pyw")


def stage(source: str = "/root/reference/FunscriptFlow.pyw") -> bool:
    """Copy the reference file, byte for byte, into baseline/_ref/ (git-ignored).  Returns True when the staged
    copy exists afterwards."""
    import shutil
    if os.path.isfile(source):
        os.makedirs(os.path.dirname(STAGED_PYW), exist_ok=True)
        if not os.path.isfile(STAGED_PYW) or open(source, "rb").read() != open(STAGED_PYW, "rb").read():
            shutil.copyfile(source, STAGED_PYW)
    return os.path.isfile(STAGED_PYW)

_KEEP = {
    "max_divergence", "radial_motion_weighted", "precompute_flow_info", "precompute_flow_info_gpu",
    "precompute_flow_info_opencl", "precompute_flow_info_dnn", "precompute_wrapper",
    "fetch_frames_optimized", "AsyncVideoReader", "VideoReaderCV", "load_strings", "process_video",
    "run_headless",
}
_KEEP_ASSIGN = {"SUPPORTED_VIDEO_EXTENSIONS", "STRINGS"}

_PRELUDE = """
import gc, os, sys, math, json, time, glob, threading, concurrent.futures
import numpy as np
import cv2
from multiprocessing import Pool
from queue import Queue, Empty
from typing import List, Dict, Optional, Tuple, Any
"""


def available() -> bool:
    return os.path.isfile(REFERENCE_PYW)


def load(module_name: str = "ffref", serial_pools: bool = False) -> types.ModuleType:
    """Return a module object holding the reference's functions.

    serial_pools=True replaces multiprocessing.Pool / ProcessPoolExecutor by in-process serial
    stand-ins (deterministic, no fork) -- used when driving ``process_video`` from tests.
    """
    if not available():
        raise FileNotFoundError(REFERENCE_PYW)
    with open(REFERENCE_PYW, "r", encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=REFERENCE_PYW)
    body = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in _KEEP:
            body.append(node)
        elif isinstance(node, ast.Assign) and any(
                isinstance(t, ast.Name) and t.id in _KEEP_ASSIGN for t in node.targets):
            body.append(node)
    mod = types.ModuleType(module_name)
    mod.__file__ = REFERENCE_PYW
    exec(_PRELUDE, mod.__dict__)
    if serial_pools:
        mod.Pool = _SerialPool
        import concurrent.futures as _cf
        mod.concurrent = types.SimpleNamespace(futures=types.SimpleNamespace(
            ProcessPoolExecutor=_SerialExecutor, ThreadPoolExecutor=_cf.ThreadPoolExecutor,
            as_completed=_cf.as_completed))
    exec(compile(ast.Module(body=body, type_ignores=[]), REFERENCE_PYW, "exec"), mod.__dict__)
    sys.modules[module_name] = mod   # lets multiprocessing pickle the functions by name
    return mod


def postproc_function(mod: types.ModuleType):
    """Wrap the statements of process_video with 1266 <= lineno < 1391 (integration .. actions)
    as ``f(final_flow_list, fps, effective_fps, params, time_stamps_unused=None) -> actions``."""
    with open(REFERENCE_PYW, "r", encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=REFERENCE_PYW)
    pv = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "process_video")
    stmts = [s for s in pv.body if 1266 <= s.lineno < 1391]
    args = ast.arguments(posonlyargs=[], args=[ast.arg("final_flow_list"), ast.arg("fps"),
                                               ast.arg("effective_fps"), ast.arg("params"),
                                               ast.arg("log_func")],
                         kwonlyargs=[], kw_defaults=[], defaults=[])
    fn = ast.FunctionDef(name="ref_postproc", args=args,
                         body=stmts + [ast.Return(ast.Name("actions", ast.Load()))],
                         decorator_list=[], lineno=1, col_offset=0)
    if sys.version_info >= (3, 12):
        fn.type_params = []
    module = ast.fix_missing_locations(ast.Module(body=[fn], type_ignores=[]))
    ns = mod.__dict__
    exec(compile(module, REFERENCE_PYW, "exec"), ns)
    return ns["ref_postproc"]


class _SerialPool:
    def __init__(self, processes=None):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def starmap(self, fn, args):
        return [fn(*a) for a in args]


class _Future:
    def __init__(self, v):
        self._v = v

    def result(self):
        return self._v


class _SerialExecutor:
    def __init__(self, max_workers=None):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def submit(self, fn, *a, **k):
        return _Future(fn(*a, **k))
