"""CPU ORACLE (test / baseline infrastructure, NOT product code) -- the reference's bracket loop
(F:1188-1242) driven with its own parallel structure: a multiprocessing.Pool over pairs for the flow
phase, the single-threaded +-6 centre mean, and a process pool for the radial pass.  The heavy
arithmetic is the very cv2 / NumPy calls the reference makes (oracle/motion_np.py cites the lines).
Used by bench.py as the `cpu_baseline` / `--impl reference` arm ("kind": "port": the reference is a
.pyw that cannot travel to the GPU box, its dependency cv2 can)."""
from __future__ import annotations

import os
import time
from multiprocessing import get_context
from typing import Dict, Sequence

import numpy as np

from . import motion_np as mo


def _phase1(args):
    p0, p1, cfg = args
    import cv2
    cv2.setNumThreads(1)           # parallelism comes from the pool, as in the reference
    return mo.precompute_flow_info(p0, p1, cfg)


def _warm(i):
    import cv2  # noqa: F401
    return i


def _phase2(args):
    flow, center, cut, pov = args
    return mo.radial_motion_weighted(flow, center, cut, pov)


def usable_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_bracket(frames: Sequence[np.ndarray], config: Dict, processes: int):
    """Returns (scalars, cuts, seconds_total, seconds_flow_phase)."""
    # "spawn", not "fork": the parent has usually run cv2 already (its worker threads do not survive a
    # fork and the first cv2 call in a forked child can deadlock)
    ctx = get_context("spawn")
    pairs = [(a, b, config) for a, b in zip(frames[:-1], frames[1:])]
    with ctx.Pool(processes=processes) as pool:
        pool.map(_warm, range(processes * 2), chunksize=1)   # worker start-up (spawn + imports) is not timed:
        t0 = time.perf_counter()                              # the reference forks, which is near free
        infos = pool.map(_phase1, pairs, chunksize=1)
        t1 = time.perf_counter()
        centers = mo.smooth_centers([i["pos_center"] for i in infos])
        pov = bool(config.get("pov_mode", False))
        vals = pool.map(_phase2, [(i["flow"], centers[j], i["cut"], pov) for j, i in enumerate(infos)], chunksize=1)
    t2 = time.perf_counter()
    return np.asarray(vals), np.array([i["cut"] for i in infos]), t2 - t0, t1 - t0
