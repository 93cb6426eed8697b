"""Deterministic synthetic clips: a seeded low-pass random texture warped
by a Gaussian-windowed radial "breathing" blob, optional global pan and hard scene cuts.

Host-side data generation only (NumPy + cv2.remap); it is the input of the hot path, not part
of it.  Used by tests/, bench.py and the golden-vector script so that every arm sees the same
frames.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np

try:  # cv2 is only needed to *generate* frames
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def base_texture(width: int, height: int, seed: int, sigma: Optional[float] = None) -> np.ndarray:
    """T_seed = 127 + 60 * N / max|N| with N = GaussianBlur(standard_normal, sigma)."""
    if sigma is None:
        sigma = 3.0 if min(width, height) >= 1080 else 2.0
    rng = np.random.default_rng(seed)
    n = rng.standard_normal((height, width)).astype(np.float32)
    n = cv2.GaussianBlur(n, (0, 0), sigma)
    return (127.0 + 60.0 * n / np.abs(n).max()).astype(np.float32)


@dataclass
class ClipSpec:
    width: int
    height: int
    n_frames: int
    seed: int = 0
    amplitude: float = 0.15          # A in s(t) = A sin(2 pi t / period)
    period: float = 30.0             # frames
    pan: Tuple[float, float] = (0.0, 0.0)   # px per frame
    cuts: Sequence[int] = field(default_factory=tuple)   # frame indices where a new scene starts
    center: Tuple[float, float] = (0.55, 0.45)           # blob centre as a fraction of (W, H)
    sigma_b: float = 0.12            # blob radius as a fraction of min(W, H)
    stereo: bool = False             # side-by-side: right half repeats the left half pattern

    def scene_of(self, t: int) -> int:
        return int(sum(1 for c in self.cuts if t >= c))


class ClipGenerator:
    """Iterates uint8 [H, W] frames of a ClipSpec; frames are generated on demand."""

    def __init__(self, spec: ClipSpec):
        if cv2 is None:
            raise RuntimeError("cv2 is required to generate synthetic clips")
        self.spec = spec
        self._tex = {}
        h, w = spec.height, spec.width
        self._ys, self._xs = np.mgrid[0:h, 0:w].astype(np.float32)

    def _scene(self, scene: int):
        s = self.spec
        if scene not in self._tex:
            w = s.width // 2 if s.stereo else s.width
            tex = base_texture(w, s.height, s.seed + 1000 * scene)
            if scene % 2 == 1:
                tex = 254.0 - tex  # contrast inversion makes a cut unmistakable
            if s.stereo:
                tex = np.concatenate([tex, np.roll(tex, 4, axis=1)], axis=1)
            cx = (s.center[0] + 0.07 * scene) % 1.0
            cy = (s.center[1] + 0.05 * scene) % 1.0
            self._tex = {scene: (tex, cx, cy)}  # keep only the live scene (memory)
        return self._tex[scene]

    def frame(self, t: int) -> np.ndarray:
        s = self.spec
        tex, cfx, cfy = self._scene(s.scene_of(t))
        cx, cy = cfx * s.width, cfy * s.height
        if s.stereo:
            cx = cfx * s.width / 2
        sb = s.sigma_b * min(s.width // 2 if s.stereo else s.width, s.height)
        st = s.amplitude * math.sin(2.0 * math.pi * t / s.period)
        px = self._xs - cx
        py = self._ys - cy
        g = np.exp(-(px * px + py * py) / (2.0 * sb * sb)).astype(np.float32)
        if s.stereo:
            px2 = self._xs - (cx + s.width / 2)
            g2 = np.exp(-(px2 * px2 + py * py) / (2.0 * sb * sb)).astype(np.float32)
            dx = st * (px * g + px2 * g2)
            dy = st * py * (g + g2)
        else:
            dx = st * px * g
            dy = st * py * g
        dx = dx + np.float32(s.pan[0] * t)
        dy = dy + np.float32(s.pan[1] * t)
        out = cv2.remap(tex, (self._xs - dx).astype(np.float32), (self._ys - dy).astype(np.float32),
                        cv2.INTER_CUBIC, borderMode=cv2.BORDER_REFLECT_101)
        return np.clip(out, 0, 255).astype(np.uint8)

    def frames(self, start: int = 0, stop: Optional[int] = None) -> Iterator[np.ndarray]:
        stop = self.spec.n_frames if stop is None else stop
        for t in range(start, stop):
            yield self.frame(t)

    def stack(self, start: int = 0, stop: Optional[int] = None) -> np.ndarray:
        return np.stack(list(self.frames(start, stop)))


# The five BASELINE.json configs as concrete specs (frame counts are the full clips; tests and
# the bench take windows of them).
def config_spec(name: str) -> ClipSpec:
    name = name.upper()
    if name == "C1":   # 640x360, 300 frames, 30 fps, radial blob, no pan, no cuts
        return ClipSpec(640, 360, 300, seed=0, amplitude=0.15, period=30.0)
    if name == "C2":   # 1920x1080, 30 fps, 10 min
        return ClipSpec(1920, 1080, 18000, seed=0, amplitude=0.15, period=30.0)
    if name == "C3":   # 3840x2160, 60 fps, pan + hard cuts every ~10 s
        return ClipSpec(3840, 2160, 3600, seed=3, amplitude=0.15, period=60.0, pan=(1.5, 0.0),
                        cuts=tuple(range(600, 3600, 600)))
    if name == "C4":   # VR side-by-side 5760x2880
        return ClipSpec(5760, 2880, 600, seed=4, amplitude=0.15, period=30.0, stereo=True)
    if name == "C5":   # one of 64 1080p videos (seed selects the video)
        return ClipSpec(1920, 1080, 1800, seed=5, amplitude=0.15, period=30.0)
    if name == "P256":  # what the product GUI really feeds: 256x256
        return ClipSpec(256, 256, 300, seed=7, amplitude=0.15, period=30.0)
    raise KeyError(name)


def make_clip(width: int, height: int, n_frames: int, **kw) -> np.ndarray:
    return ClipGenerator(ClipSpec(width, height, n_frames, **kw)).stack()
