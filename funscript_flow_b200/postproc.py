"""Host-side 1-D post-processing: per-pair scalars -> funscript actions.

Mirrors FunscriptFlow.pyw F:1266-1395 (integration with resets at cuts, half-step shift,
Hann-blended windowed linear detrend, 5-tap binomial smoothing, rolling min/max normalisation,
slope-inversion keyframes, JSON writer).  O(N) NumPy on the host: there is nothing here worth a
kernel, but keyframe parity with the reference needs every rounding rule reproduced.
"""
from __future__ import annotations

import json
import math
from typing import Optional, Dict, List, Sequence

import numpy as np

DISCONTINUITY = 1000.0          # F:1288 (hard-coded in the reference)
SMOOTH_TAPS = np.array([1 / 16, 1 / 4, 3 / 8, 1 / 4, 1 / 16])   # F:1333


def sampling_step(fps: float) -> int:
    """F:1127: frames are sub-sampled to at most ~30 fps."""
    return max(1, int(math.ceil(fps / 30.0)))


def integrate(values: Sequence[float], cuts: Sequence[bool]) -> np.ndarray:
    """F:1267-1284: trapezoid running sum that restarts at 0 on a cut, then the half-step shift."""
    v = np.asarray(values, dtype=np.float64)
    n = len(v)
    cum = np.zeros(n, dtype=np.float64)
    for i in range(1, n):   # sequential on purpose: same summation order as the reference
        cum[i] = 0.0 if cuts[i] else cum[i - 1] + (v[i - 1] + v[i]) / 2
    out = cum.copy()
    out[1:] = (cum[1:] + cum[:-1]) / 2
    return out


def detrend(cum: np.ndarray, window: int) -> np.ndarray:
    """F:1290-1331: per continuous segment, overlapping windows of `window` samples (hop
    window//2), least-squares line removed, Hann-weighted overlap-add."""
    n = len(cum)
    acc = np.zeros(n)
    wsum = np.zeros(n)
    breaks = (np.flatnonzero(np.abs(np.diff(cum)) > DISCONTINUITY) + 1).tolist()
    edges = [0] + breaks + [n]
    hop = window // 2
    for a, b in zip(edges[:-1], edges[1:]):
        if b - a < 5:                      # F:1306-1308
            acc[a:b] = cum[a:b] - np.mean(cum[a:b])
            continue
        starts = [a] if b - a <= window else range(a, b - hop, hop)
        for s in starts:
            e = b if b - a <= window else min(s + window, b)
            seg = cum[s:e]
            t = np.arange(e - s)
            line = np.polyval(np.polyfit(t, seg, 1), t)
            hann = np.hanning(e - s)
            acc[s:e] += (seg - line) * hann
            wsum[s:e] += hann
    return acc / np.maximum(wsum, 1e-6)


def normalise(sig: np.ndarray, window: int) -> np.ndarray:
    """F:1335-1349: rolling min/max over an odd window -> 0..100 (50 where the window is flat)."""
    if window % 2 == 0:
        window += 1
    half = window // 2
    n = len(sig)
    out = np.empty(n)
    for i in range(n):
        loc = sig[max(0, i - half): min(n, i + half + 1)]
        lo, hi = loc.min(), loc.max()
        out[i] = 50 if hi - lo == 0 else (sig[i] - lo) / (hi - lo) * 100
    return out


def keyframes(norm: np.ndarray, reduce: bool) -> List[int]:
    """F:1366-1376: first, last and every slope inversion."""
    n = len(norm)
    if not reduce:
        return list(range(n))
    d = np.diff(norm) < 0
    inner = (np.flatnonzero(d[:-1] != d[1:]) + 1).tolist() if n > 2 else []
    return [0] + inner + [n - 1]


def scalars_to_actions(values: Sequence[float], cuts: Sequence[bool], frame_indices: Sequence[int], fps: float,
                       params: Dict, errors: Optional[List[str]] = None) -> List[Dict[str, int]]:
    """The whole of F:1266-1386 for one video."""
    eff_fps = fps / sampling_step(fps)
    cum = integrate(values, cuts)
    det = detrend(cum, int(params["detrend_window"] * eff_fps))
    smooth = np.convolve(det, SMOOTH_TAPS, mode="same")
    norm = normalise(smooth, int(params["norm_window"] * eff_fps))
    # np.convolve(mode="same") hands back max(n, 5) samples: for a series shorter than the 5-tap smoother the
    # reference normalises and picks keyframes over those 5 samples and drops the indices without a time stamp
    # (F:1379-1385, where it also flags the video as failed) -- kept as is; `errors` collects the dropped indices
    acts = []
    for k in keyframes(norm, bool(params["keyframe_reduction"])):
        if k >= len(frame_indices):
            if errors is not None:
                errors.append(f"Error computing action at segment index {k}: list index out of range")
            continue
        acts.append({"at": int((frame_indices[k] / fps) * 1000), "pos": 100 - int(round(norm[k]))})   # F:1380-1382
    return acts


def write_funscript(path: str, actions: List[Dict[str, int]]) -> None:
    """F:1391-1394."""
    with open(path, "w") as fh:
        json.dump({"version": "1.0", "actions": actions}, fh, indent=2)
