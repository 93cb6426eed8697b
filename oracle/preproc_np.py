"""CPU ORACLE (test infrastructure, NOT product code) -- frame pre-processing of the reference
(SURVEY.md row N2): decoded BGR frame -> the 256x256 grayscale frame the flow consumes.

Reference lines (F:n = FunscriptFlow.pyw):  BGR->RGB + resize to (256, 256) in the reader (F:173-189,
F:1057-1065); VR mode: no resize in the reader, then cv2.resize(f, (512, 512)) and the crop
[256:, :256] (F:1074-1079); cv2.cvtColor(..., COLOR_RGB2GRAY) (F:1079, F:1082).

The arithmetic is OpenCV's 8-bit fixed-point path (un-vendored dependency, restated from its published
algorithm and pinned bit-exactly against the installed cv2 by tests/test_oracle.py):
  resize INTER_LINEAR, uint8: 11-bit coefficients  round(w * 2048)  per axis, horizontal pass in int,
      vertical  dst = (((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2
  RGB2GRAY, uint8:  (9798 R + 19235 G + 3735 B + 16384) >> 15
"""
from __future__ import annotations

import numpy as np

OUT = 256          # F:1057
VR_SIZE = 512      # F:1076


def resize_tables(dst_n: int, src_n: int, reset_at_borders: bool):
    """cv::resize INTER_LINEAR tables for uint8: the two source indices and the two 11-bit weights.
    Along x OpenCV resets the weights to (1, 0) where the footprint leaves the image; along y it only
    clips the row indices and keeps the fractional weights (visible when up-scaling: the first and
    last rows are  b0*row + b1*row  with b0 + b1 = 2048, rounded separately)."""
    scale = src_n / dst_n
    d = np.arange(dst_n, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if reset_at_borders:
        lo = s < 0
        s[lo] = 0
        f[lo] = 0
        hi = s >= src_n - 1
        s[hi] = src_n - 1
        f[hi] = 0
    a1 = np.rint(f * np.float32(2048)).astype(np.int64)
    a0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int64)
    return np.clip(s, 0, src_n - 1), np.clip(s + 1, 0, src_n - 1), a0, a1


def resize_u8_linear(src: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """cv2.resize(src, (dst_w, dst_h)) for uint8 images [H, W] or [H, W, C]."""
    sh, sw = src.shape[:2]
    if (sw, sh) == (dst_w, dst_h):
        return src.copy()
    x0, x1, xa0, xa1 = resize_tables(dst_w, sw, True)
    y0, y1, ya0, ya1 = resize_tables(dst_h, sh, False)
    S = src.astype(np.int64)
    ex = (None, slice(None), None) if S.ndim == 3 else (None, slice(None))
    ey = (slice(None), None, None) if S.ndim == 3 else (slice(None), None)
    H = S[:, x0] * xa0[ex] + S[:, x1] * xa1[ex]
    out = (((ya0[ey] * (H[y0] >> 4)) >> 16) + ((ya1[ey] * (H[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def rgb_to_gray(rgb: np.ndarray) -> np.ndarray:
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    return ((9798 * r + 19235 * g + 3735 * b + 16384) >> 15).astype(np.uint8)


def frame_to_gray(bgr: np.ndarray, vr_mode: bool = False) -> np.ndarray:
    """Decoded BGR frame -> uint8 [256, 256] exactly as fetch_frames_optimized feeds the flow."""
    rgb = bgr[..., ::-1]
    if vr_mode:
        return rgb_to_gray(resize_u8_linear(rgb, VR_SIZE, VR_SIZE)[VR_SIZE // 2:, :VR_SIZE // 2])
    return rgb_to_gray(resize_u8_linear(rgb, OUT, OUT))


def frame_window_to_gray(bgr: np.ndarray, target, window) -> np.ndarray:
    """Row N4 generalisation: cv2.resize(rgb, target)[y:y+h, x:x+w] -> gray, with target=(w, h) and
    window=(x, y, w, h).  frame_to_gray is target=(256,256), window=(0,0,256,256) and, in VR mode,
    target=(512,512), window=(0,256,256,256)."""
    x, y, w, h = window
    rgb = resize_u8_linear(np.ascontiguousarray(bgr[..., ::-1]), int(target[0]), int(target[1]))
    return rgb_to_gray(rgb[y:y + h, x:x + w])
