"""REFERENCE ARM (baseline infrastructure, NOT product code) -- the reference's own functions, loaded unmodified
from FunscriptFlow.pyw (oracle/ref_loader.py), driven the way its bracket loop drives them:

    F:1190-1191   Pool(processes=threads).starmap(precompute_wrapper, [(p, params) for p in pairs])
    F:1201-1214   single-threaded +-6 centre mean (np.mean of the neighbours' pos_center, truncated at the ends)
    F:1232-1236   ProcessPoolExecutor(max_workers=threads).submit(radial_motion_weighted, flow, centre, cut, pov)

Only the orchestration lines above are restated here; every function that computes is the reference's.  The
worker processes are spawned (not forked: the parent has usually run cv2, whose threads do not survive a fork)
and load the reference module in their initializer, so the functions pickle by name exactly as under fork.
Worker start-up is excluded from the timing (the reference forks, which is nearly free)."""
from __future__ import annotations

import concurrent.futures
import time
from multiprocessing import get_context
from typing import Dict, Sequence

import numpy as np

from . import ref_loader

MODULE = "ffref"


def available() -> bool:
    return ref_loader.available()


def _init_worker():
    ref_loader.load(MODULE)


def _warm(i):
    import cv2  # noqa: F401
    return i


def run_bracket(frames: Sequence[np.ndarray], params: Dict, threads: int):
    """One bracket through the reference's functions.  Returns (scalars, cuts, seconds_total, seconds_flow_phase)."""
    ref = ref_loader.load(MODULE)
    prm = dict(params)
    prm.setdefault("threads", threads)
    ctx = get_context("spawn")
    pairs = list(zip(frames[:-1], frames[1:]))                                        # F:1188
    with ctx.Pool(processes=threads, initializer=_init_worker) as pool, \
            concurrent.futures.ProcessPoolExecutor(max_workers=threads, mp_context=ctx, initializer=_init_worker) as ex:
        pool.map(_warm, range(threads * 2), chunksize=1)
        list(ex.map(_warm, range(threads * 2)))
        t0 = time.perf_counter()
        precomputed = pool.starmap(ref.precompute_wrapper, [(p, prm) for p in pairs])  # F:1190-1191
        t1 = time.perf_counter()
        final_centers = []
        for j, info in enumerate(precomputed):                                         # F:1203-1214
            center_list = [info["pos_center"]]
            for i in range(1, 7):
                if j - i >= 0:
                    center_list.append(precomputed[j - i]["pos_center"])
                if j + i < len(precomputed):
                    center_list.append(precomputed[j + i]["pos_center"])
            final_centers.append(np.mean(np.array(center_list), axis=0))
        futures = [ex.submit(ref.radial_motion_weighted, info["flow"], final_centers[j], info["cut"], prm.get("pov_mode", False))
                   for j, info in enumerate(precomputed)]                               # F:1232-1236
        vals = [f.result() for f in futures]
        t2 = time.perf_counter()
    return np.asarray(vals, np.float64), np.array([bool(i["cut"]) for i in precomputed]), t2 - t0, t1 - t0
