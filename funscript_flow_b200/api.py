"""Drop-in processing functions with the reference's names, signatures and return shapes
(FunscriptFlow.pyw, "F:n" = line n), backed by the sm_100a kernels through the C ABI.

    precompute_flow_info(p0, p1, config)            F:843-907   (8-key dict, F:898-907)
    precompute_flow_info_gpu(p0, p1, cut_threshold) F:982-1017  (the slot this build fills)
    precompute_wrapper(p, params)                   F:1019-1021
    max_divergence(flow)                            F:748-758
    radial_motion_weighted(flow, center, is_cut, pov_mode=False)   F:761-785
    get_available_backends()                        F:32-63

plus the batched entry the GPU runner uses, because per-pair calls from forked pool workers
(F:1190-1191) cannot drive a GPU:

    process_bracket(frames, params) -> dict of per-pair arrays   (F:1188-1242 in one call)

There is no CPU path here: every function raises if libffb.so or a CUDA device is missing.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Sequence

import numpy as np

from . import _native

DEFAULT_CUT_THRESHOLD = 7          # F:876
DEFAULT_BATCH_FRAMES = 64

_contexts: Dict[int, _native.FlowContext] = {}


def default_device() -> int:
    """FFB_DEVICE, else the torchrun LOCAL_RANK folded onto the visible devices (ranks may share a GPU), else 0."""
    if os.environ.get("FFB_DEVICE", "") != "":
        return int(os.environ["FFB_DEVICE"])
    if os.environ.get("LOCAL_RANK", "") != "":
        try:
            n = _native.device_count()
        except Exception:
            n = 0
        return int(os.environ["LOCAL_RANK"]) % max(1, n)
    return 0


def get_context(device: Optional[int] = None) -> _native.FlowContext:
    """Process-wide context per device (created lazily; raises without a GPU)."""
    dev = default_device() if device is None else int(device)
    ctx = _contexts.get(dev)
    if ctx is None:
        ctx = _native.FlowContext(dev)
        _contexts[dev] = ctx
    return ctx


_aux_contexts: Dict[int, _native.FlowContext] = {}


def get_aux_context(primary: _native.FlowContext) -> _native.FlowContext:
    """A second context on the device (and from the library) of `primary`, for work that runs beside it:
    the other eye of a side-by-side VR frame (runner, vr_eye="both")."""
    key = id(primary)
    ctx = _aux_contexts.get(key)
    if ctx is None:
        ctx = _native.FlowContext(primary.device, lib_path=primary.lib_path)
        _aux_contexts.clear()          # one auxiliary context at a time
        _aux_contexts[key] = ctx
    return ctx


def set_context(ctx: _native.FlowContext, device: int = 0) -> None:
    """Install an externally created context (tests use this to inject the emulated library)."""
    _contexts[int(device)] = ctx


def get_available_backends() -> Dict[str, bool]:
    """Shape of F:32-63's result.  Only the CUDA entry can be True: this build ships no CPU,
    OpenCL or DIS path."""
    try:
        n = _native.device_count()
    except Exception:
        n = 0
    return {"CPU": False, "CUDA": n > 0, "OpenCL": False, "DNN": False}


def get_gpu_info() -> str:
    """F:64-99: "CUDA: <device name>" for one GPU, "CUDA: <n> devices" for several, "CPU only" when there is none
    (the reference's wording for 'no accelerator'; this build cannot run in that case)."""
    try:
        n = _native.device_count()
    except Exception:
        n = 0
    if n <= 0:
        return "CPU only"
    if n == 1:
        return f"CUDA: {_native.device_name(0) or 'Device 0'}"
    return f"CUDA: {n} devices"


def _as_frames(frames) -> np.ndarray:
    if isinstance(frames, np.ndarray) and frames.ndim == 3:
        arr = frames
    else:
        arr = np.stack([np.asarray(f) for f in frames])
    if arr.dtype != np.uint8:
        raise TypeError("frames must be uint8 grayscale")
    return np.ascontiguousarray(arr)


def process_bracket(frames, params: Optional[Dict] = None, *, ctx: Optional[_native.FlowContext] = None,
                    batch_frames: int = DEFAULT_BATCH_FRAMES, return_flows: bool = False) -> Dict:
    """One bracket of N consecutive sampled frames -> N-1 pairs (F:1188-1242 as one GPU pass).

    Returns a dict with per-pair arrays: scalar (f64), cut (bool), cx / cy (int32), val (f32),
    mean_mag (f32), centers (f64 [N-1, 2], the +-6 smoothed centre) and n_pairs.  With
    return_flows=True also `flows` (f32 [k, H, W, 2]) for the last k <= ring pairs (tests).
    """
    params = params or {}
    ctx = ctx or get_context()
    arr = _as_frames(frames)
    n, h, w = arr.shape
    ctx.configure(w, h, max(1, min(batch_frames, n)), max(n - 1, 1))
    ctx.bracket_begin(bool(params.get("pov_mode", False)), float(params.get("cut_threshold", DEFAULT_CUT_THRESHOLD)))
    try:
        ctx.bracket_push(arr)
    except BaseException:
        ctx.bracket_abort()      # leave the context usable
        raise
    res = ctx.bracket_finish()
    if return_flows:
        k = res["n_pairs"]
        first = max(0, k - ctx.flow_ring_size)
        res["flow_first"] = first
        res["flows"] = np.stack([ctx.get_flow(p) for p in range(first, k)]) if k else np.empty((0, h, w, 2), np.float32)
    return res


def shard_bounds(n_pairs: int, parts: int):
    """Contiguous, balanced pair ranges [a, b) of a bracket of n_pairs pairs for `parts` GPUs: shard r
    needs frames a .. b (one frame of overlap with shard r+1).  Ranges may be empty when parts > n_pairs."""
    return [(n_pairs * r // parts, n_pairs * (r + 1) // parts) for r in range(parts)]


def shard_phase1(ctx: _native.FlowContext, frames, a: int, b: int, params: Optional[Dict] = None,
                 batch_frames: int = DEFAULT_BATCH_FRAMES) -> Dict:
    """Phase 1 of pairs [a, b) of the bracket `frames` on `ctx` (flows, raw centres, cut test; F:1188-1199).
    The shard's final flows stay resident in `ctx` until shard_phase2."""
    params = params or {}
    arr = _as_frames(frames[a:b + 1])
    n, h, w = arr.shape
    ctx.configure(w, h, max(1, min(batch_frames, n)), max(n - 1, 1))
    ctx.bracket_begin_shard(n - 1, bool(params.get("pov_mode", False)), float(params.get("cut_threshold", DEFAULT_CUT_THRESHOLD)))
    try:
        ctx.bracket_push(arr)
        return ctx.bracket_phase1_finish()
    except BaseException:
        ctx.bracket_abort()
        raise


def shard_phase2(ctx: _native.FlowContext, cx_all: np.ndarray, cy_all: np.ndarray, a: int, b: int):
    """Phase 2 of pairs [a, b): cx_all / cy_all are the raw centres of ALL pairs of the bracket (gathered from the
    shards); the +-6 window (F:1201-1214) reads [a-6, b+6) clipped to the bracket.  Returns (scalar, centers)."""
    lo, hi = max(0, a - 6), min(len(cx_all), b + 6)
    return ctx.bracket_radial(cx_all[lo:hi], cy_all[lo:hi], a - lo, b - a)


def process_bracket_on_contexts(frames, params: Optional[Dict], ctxs: Sequence[_native.FlowContext],
                                batch_frames: int = DEFAULT_BATCH_FRAMES) -> Dict:
    """One bracket cut into len(ctxs) frame-range shards, shard r on ctxs[r] (several GPUs driven by one process, or
    several contexts of one GPU).  Same dict as process_bracket, bit for bit."""
    arr = _as_frames(frames)
    n_pairs = len(arr) - 1
    bounds = shard_bounds(n_pairs, len(ctxs))
    p1 = [shard_phase1(c, arr, a, b, params, batch_frames) if b > a else None for c, (a, b) in zip(ctxs, bounds)]
    cat = {k: np.concatenate([p[k] for p in p1 if p is not None]) for k in ("cx", "cy", "val", "mean_mag", "cut")}
    parts = [shard_phase2(c, cat["cx"], cat["cy"], a, b) for c, (a, b) in zip(ctxs, bounds) if b > a]
    cat["scalar"] = np.concatenate([p[0] for p in parts])
    cat["centers"] = np.concatenate([p[1] for p in parts])
    cat["n_pairs"] = n_pairs
    return cat


class BracketPipeline:
    """Consecutive brackets (of one video, or the clips of a library) pipelined over two contexts of one GPU: bracket
    i+1 is pushed -- its frames start uploading -- before the results of bracket i are fetched, so the upload of the
    first batch and the drain of the last one no longer sit between brackets.  The kernels of the two contexts are
    chained (ffb_chain_after), only uploads overlap.  Results are those of process_bracket, bit for bit.

        pipe = BracketPipeline(ctx)
        for frames in brackets:
            done = pipe.submit(frames, params)      # returns the result of the bracket submitted before, or None
        last = pipe.flush()

    Host frames (pinned or pageable) must stay valid until their bracket's result has been returned."""

    def __init__(self, ctx: Optional[_native.FlowContext] = None, batch_frames: int = DEFAULT_BATCH_FRAMES):
        primary = ctx or get_context()
        self.ctxs = [primary, get_aux_context(primary)]
        self.batch_frames = batch_frames
        self._turn = 0
        self._pending = None        # (context, keep-alive array) of the bracket whose result has not been fetched

    def submit(self, frames, params: Optional[Dict] = None):
        params = params or {}
        arr = _as_frames(frames)
        n, h, w = arr.shape
        ctx = self.ctxs[self._turn]
        other = self.ctxs[self._turn ^ 1]
        ctx.configure(w, h, max(1, min(self.batch_frames, n)), max(n - 1, 1))
        if self._pending is not None:
            ctx.chain_after(other)
        ctx.bracket_begin(bool(params.get("pov_mode", False)), float(params.get("cut_threshold", DEFAULT_CUT_THRESHOLD)))
        try:
            ctx.bracket_push(arr)
        except BaseException:
            ctx.bracket_abort()
            raise
        done = self.flush()
        self._pending = (ctx, arr)
        self._turn ^= 1
        return done

    def flush(self):
        """Result of the bracket submitted last (None when there is none)."""
        if self._pending is None:
            return None
        ctx, _ = self._pending
        self._pending = None
        return ctx.bracket_finish()


def precompute_flow_info(p0: np.ndarray, p1: np.ndarray, config: Dict) -> Dict:
    """F:843-907.  `config["backend"]` is ignored (there is one backend); `pov_mode` and the
    hidden `cut_threshold` key are honoured like the CPU branch (F:876-894)."""
    ctx = get_context()
    p0 = np.ascontiguousarray(p0, dtype=np.uint8)
    p1 = np.ascontiguousarray(p1, dtype=np.uint8)
    if p0.shape != p1.shape or p0.ndim != 2:
        raise ValueError("p0 and p1 must be equal-sized 2-D uint8 arrays")
    h, w = p0.shape
    pov = bool(config.get("pov_mode"))
    ctx.configure(w, h, 2, 1)       # incremental: a context already configured for this frame size is left alone
    ctx.bracket_begin(pov, float(config.get("cut_threshold", DEFAULT_CUT_THRESHOLD)))
    try:
        ctx.bracket_push(p0)
        ctx.bracket_push(p1)
    except BaseException:
        ctx.bracket_abort()
        raise
    r = ctx.bracket_finish()
    flow = ctx.get_flow(0)
    if pov:   # F:882: plain Python ints and the int 0
        pos_center, val = (w // 2, h - 1), 0
    else:     # F:757-758: NumPy integer / float32 scalars
        pos_center, val = (np.int64(r["cx"][0]), np.int64(r["cy"][0])), np.float32(r["val"][0])
    return {
        "flow": flow,
        "pos_center": pos_center,
        "neg_center": pos_center,
        "val_pos": val,
        "val_neg": val,
        "cut": bool(r["cut"][0]),
        "cut_center": pos_center[0],
        "mean_mag": np.float32(r["mean_mag"][0]),
    }


def precompute_flow_info_gpu(p0: np.ndarray, p1: np.ndarray, cut_threshold) -> Dict:
    """F:982-1017: same dict, threshold passed positionally, no config (hence no POV mode)."""
    return precompute_flow_info(p0, p1, {"cut_threshold": cut_threshold})


def precompute_wrapper(p, params):
    """F:1019-1021."""
    return precompute_flow_info(p[0], p[1], params)


def max_divergence(flow: np.ndarray):
    """F:748-758: (x, y, value) of the first maximum of |d flow[...,0]/d row + d flow[...,1]/d col|."""
    x, y, v = get_context().max_divergence(flow)
    return np.int64(x), np.int64(y), v


def radial_motion_weighted(flow: np.ndarray, center, is_cut, pov_mode: bool = False):
    """F:761-785."""
    if is_cut:
        return 0.0
    return np.float64(get_context().radial_motion(flow, center, False, bool(pov_mode)))


def smooth_centers(centers: Sequence, radius: int = 6) -> np.ndarray:
    """F:1201-1214 on the host (the runner does this on the device; kept for API parity)."""
    c = np.asarray(centers, dtype=np.int64).reshape(-1, 2)
    n = len(c)
    csum = np.concatenate([np.zeros((1, 2), np.int64), np.cumsum(c, axis=0)])
    j = np.arange(n)
    lo = np.maximum(j - radius, 0)
    hi = np.minimum(j + radius + 1, n)
    return (csum[hi] - csum[lo]) / (hi - lo)[:, None].astype(np.float64)
