"""CPU ORACLE (test infrastructure, NOT product code) -- NumPy restatement of the reference's
per-pair motion functions and of the 1-D post-processing that turns the per-pair scalars into
funscript actions.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import it.

Each function cites the reference lines it follows (F:n = /root/reference/FunscriptFlow.pyw
line n).  The functions are pinned two ways (tests/test_oracle_golden.py):
  * against known answers recorded from the *reference's own functions* (AST-loaded from the
    .pyw in the build container by tests/golden/make_golden.py -> tests/golden/*.npz/json);
  * live against the AST-loaded reference whenever /root/reference is present.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np

DEFAULT_CUT_THRESHOLD = 7  # F:876


# ----------------------------------------------------------------------------- A3
def max_divergence(flow: np.ndarray) -> Tuple[int, int, np.float32]:
    """F:748-758.  The quantity is d(flow[...,0])/d(row) + d(flow[...,1])/d(col) -- i.e. with the
    axes swapped relative to a true divergence; np.gradient rules (central /2 inside, one-sided
    at the edges, float32); first maximum of |.| in C order.  Returns (x, y, value)."""
    u = np.asarray(flow[..., 0], dtype=np.float32)
    v = np.asarray(flow[..., 1], dtype=np.float32)
    h, w = u.shape
    du = np.empty_like(u)
    du[1:-1] = (u[2:] - u[:-2]) / np.float32(2.0)
    du[0] = u[1] - u[0]
    du[-1] = u[-1] - u[-2]
    dv = np.empty_like(v)
    dv[:, 1:-1] = (v[:, 2:] - v[:, :-2]) / np.float32(2.0)
    dv[:, 0] = v[:, 1] - v[:, 0]
    dv[:, -1] = v[:, -1] - v[:, -2]
    d = du + dv
    flat = int(np.argmax(np.abs(d)))
    y, x = divmod(flat, w)
    return x, y, d[y, x]


def divergence_field(flow: np.ndarray) -> np.ndarray:
    """The swapped-axis 'divergence' field itself (for margin guards in tests)."""
    return (np.gradient(flow[..., 0], axis=0) + np.gradient(flow[..., 1], axis=1)).astype(np.float32)


# ----------------------------------------------------------------------------- A4
def mean_magnitude(flow: np.ndarray) -> np.float32:
    """F:889-890: float32 per-pixel magnitude, float32 mean."""
    u = flow[..., 0].astype(np.float32)
    v = flow[..., 1].astype(np.float32)
    return np.mean(np.sqrt(u * u + v * v, dtype=np.float32), dtype=np.float32)


def is_cut(mean_mag, cut_threshold=DEFAULT_CUT_THRESHOLD) -> bool:
    """F:891-894: strict greater-than."""
    return bool(mean_mag > cut_threshold)


# ----------------------------------------------------------------------------- A6
def radial_motion_weighted(flow: np.ndarray, center, cut: bool, pov_mode: bool = False) -> float:
    """F:761-785.  0.0 on a cut; otherwise the mean over all pixels of
    (u*(x-cx) + v*(y-cy)) * wx(x) * wy(y) with wx = (W-x)/W right of the centre, x/W at or left
    of it (same for y); in POV mode the unweighted mean.  float64 like the reference."""
    if cut:
        return 0.0
    h, w = flow.shape[:2]
    xs = np.arange(w, dtype=np.int64)[None, :]
    ys = np.arange(h, dtype=np.int64)[:, None]
    cx, cy = float(center[0]), float(center[1])
    dot = flow[..., 0] * (xs - cx) + flow[..., 1] * (ys - cy)
    if pov_mode:
        return float(np.mean(dot))
    acc = np.where(xs > cx, dot * (w - xs) / w, dot * xs / w)
    acc = np.where(ys > cy, acc * (h - ys) / h, acc * ys / h)
    return float(np.mean(acc))


# ----------------------------------------------------------------------------- A5
def smooth_centers(centers: Sequence[Tuple[int, int]], radius: int = 6) -> np.ndarray:
    """F:1201-1214: mean of each pair's centre with up to `radius` neighbours on both sides,
    truncated at the bracket ends (no outlier rejection).  int -> float64 [N, 2]."""
    c = np.asarray(centers, dtype=np.int64).reshape(-1, 2)
    n = len(c)
    out = np.empty((n, 2), dtype=np.float64)
    for j in range(n):
        lo, hi = max(0, j - radius), min(n, j + radius + 1)
        out[j] = c[lo:hi].sum(axis=0) / float(hi - lo)
    return out


# ----------------------------------------------------------------------------- A0
def precompute_flow_info(p0: np.ndarray, p1: np.ndarray, config: Dict, flow_fn=None) -> Dict:
    """F:875-907 (the CPU branch).  `flow_fn(p0, p1) -> f32[H,W,2]`; default is cv2's Farneback
    with the reference's parameters (F:878-879) -- the very call the reference makes."""
    if flow_fn is None:
        import cv2
        flow = cv2.calcOpticalFlowFarneback(p0, p1, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    else:
        flow = flow_fn(p0, p1)
    if config.get("pov_mode"):
        x, y, val = p0.shape[1] // 2, p0.shape[0] - 1, 0       # F:880-882
    else:
        x, y, val = max_divergence(flow)                        # F:884
    mm = mean_magnitude(flow)
    cut = is_cut(mm, config.get("cut_threshold", DEFAULT_CUT_THRESHOLD))
    return {"flow": flow, "pos_center": (x, y), "neg_center": (x, y), "val_pos": val, "val_neg": val,
            "cut": cut, "cut_center": x, "mean_mag": mm}


def process_bracket(frames: Sequence[np.ndarray], config: Dict, flow_fn=None):
    """F:1188-1242 for one bracket, serially: phase 1 per pair, centre smoothing, radial pass.
    Returns (scalars f64[N-1], cuts bool[N-1], infos)."""
    infos = [precompute_flow_info(a, b, config, flow_fn) for a, b in zip(frames[:-1], frames[1:])]
    centers = smooth_centers([i["pos_center"] for i in infos])
    vals = np.array([radial_motion_weighted(i["flow"], centers[j], i["cut"], bool(config.get("pov_mode", False)))
                     for j, i in enumerate(infos)], dtype=np.float64)
    cuts = np.array([i["cut"] for i in infos], dtype=bool)
    return vals, cuts, infos


# ----------------------------------------------------------------------------- N1
def postprocess(final_flow_list: List[Tuple[float, bool, int]], fps: float, params: Dict):
    """F:1266-1386: integrate (reset at cuts), half-step shift, windowed linear detrend with Hann
    blending, 5-tap binomial smooth, rolling min/max normalisation, optional slope-inversion
    keyframes.  Returns the list of {"at", "pos"} actions."""
    step = max(1, int(math.ceil(fps / 30.0)))                    # F:1127
    eff_fps = fps / step                                         # F:1128
    n = len(final_flow_list)
    vals = [float(v) for v, _, _ in final_flow_list]
    cuts = [bool(c) for _, c, _ in final_flow_list]
    stamps = [t for _, _, t in final_flow_list]

    cum = [0.0]                                                  # F:1267-1281
    for i in range(1, n):
        cum.append(0.0 if cuts[i] else cum[-1] + (vals[i - 1] + vals[i]) / 2)
    cum = [cum[0]] + [(cum[i] + cum[i - 1]) / 2 for i in range(1, n)]   # F:1284
    cum = np.asarray(cum, dtype=np.float64)

    win = int(params["detrend_window"] * eff_fps)                # F:1287
    det = np.zeros(n)
    wsum = np.zeros(n)
    jumps = np.nonzero(np.abs(np.diff(cum)) > 1000)[0] + 1       # F:1288-1294
    bounds = [0, *[int(j) for j in jumps], n]
    half = win // 2
    for s0, s1 in zip(bounds[:-1], bounds[1:]):                  # F:1300-1328
        length = s1 - s0
        if length < 5:
            det[s0:s1] = cum[s0:s1] - np.mean(cum[s0:s1])
            continue
        if length <= win:
            spans = [(s0, s1)]
        else:
            spans = [(a, min(a + win, s1)) for a in range(s0, s1 - half, half)]
        for a, b in spans:
            seg = cum[a:b]
            t = np.arange(b - a)
            fit = np.polyfit(t, seg, 1)
            wts = np.hanning(b - a)
            det[a:b] += (seg - np.polyval(fit, t)) * wts
            wsum[a:b] += wts
    det = det / np.maximum(wsum, 1e-6)                           # F:1331

    sm = np.convolve(det, [1 / 16, 1 / 4, 3 / 8, 1 / 4, 1 / 16], mode="same")   # F:1333
    # np.convolve(mode="same") returns max(n, 5) samples: for a series shorter than the smoother the reference
    # carries on with 5 samples (F:1340, F:1369) and drops the actions whose index has no time stamp (the
    # IndexError is caught at F:1383-1385) -- restated as is
    m = len(sm)
    nwin = int(params["norm_window"] * eff_fps)                  # F:1335-1349
    if nwin % 2 == 0:
        nwin += 1
    hn = nwin // 2
    norm = np.empty(m)
    for i in range(m):
        loc = sm[max(0, i - hn): min(m, i + hn + 1)]
        lo, hi = loc.min(), loc.max()
        norm[i] = 50 if hi - lo == 0 else (sm[i] - lo) / (hi - lo) * 100

    if params["keyframe_reduction"]:                             # F:1366-1376
        keys = [0] + [i for i in range(1, m - 1)
                      if ((norm[i] - norm[i - 1]) < 0) != ((norm[i + 1] - norm[i]) < 0)] + [m - 1]
    else:
        keys = list(range(m))
    keys = [k for k in keys if k < n]                            # F:1379-1385
    return [{"at": int((stamps[k] / fps) * 1000), "pos": 100 - int(round(norm[k]))} for k in keys]   # F:1378-1382
