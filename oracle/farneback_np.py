"""CPU ORACLE (test infrastructure, NOT product code) -- stage-wise NumPy restatement of
``cv2.calcOpticalFlowFarneback(prev, next, None, 0.5, 3, 15, 3, 5, 1.2, 0)``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path (``funscript_flow_b200``)
never does.

Where the arithmetic lives
--------------------------
The reference (``/root/reference/FunscriptFlow.pyw``) calls OpenCV at F:878-879.  OpenCV is an
un-vendored third-party dependency (``opencv-python 4.11.0.86`` in ``uv.lock:238-239``; this
image ships ``opencv-python-headless 4.13.0``), so its source is not under ``/root/reference``.
This file restates the *published* algorithm of ``modules/video/src/optflowgf.cpp`` (Farneback
polynomial expansion) plus the two imgproc primitives it leans on (``GaussianBlur`` with
``BORDER_REFLECT_101`` and ``resize(INTER_LINEAR)``), written from the specification in
SURVEY.md section 3.4, and is *pinned behaviourally*: ``tests/test_oracle_vs_cv2.py`` asserts
it against the installed ``cv2`` on every test clip, and ``tests/golden/*.npz`` hold cv2 /
reference outputs generated in the build container by ``tests/golden/make_golden.py``.

Every stage is exposed on its own so that each CUDA kernel has a per-stage checker:

    pyramid_level   (A1a)   u8 image        -> f32 level image
    poly_exp        (A1b)   f32 level image -> f32 [h, w, 5] polynomial coefficients
    update_matrices (A1c)   R0, R1, flow    -> f32 [h, w, 5]
    blur_solve      (A1d)   M               -> f32 [h, w, 2] flow
    upsample_flow   (A1e)   coarse flow     -> f32 [h, w, 2] initial flow (x2)
    farneback               u8, u8          -> f32 [H, W, 2]

``acc`` selects the accumulator type of the polyexp horizontal pass and of the box sums:
``np.float64`` mirrors the CPU implementation of OpenCV (double accumulators), ``np.float32``
mirrors what an all-fp32 GPU kernel computes (used to budget tolerances, never as the checker
of record).
"""
from __future__ import annotations

import math

import numpy as np

# Parameters the reference passes at F:878-879.
PYR_SCALE = 0.5
LEVELS = 3
WINSIZE = 15
ITERATIONS = 3
POLY_N = 5
POLY_SIGMA = 1.2
MIN_SIZE = 32  # optflowgf.cpp: coarsest level must be at least 32 px in both dimensions

BORDER_ATTEN = np.array([0.14, 0.14, 0.4472, 0.4472, 0.4472], dtype=np.float32)


def cv_round(x: float) -> int:
    """cvRound: round half to even (SSE cvtsd2si semantics)."""
    return int(np.rint(x))


# --------------------------------------------------------------------------------------
# level geometry
# --------------------------------------------------------------------------------------
def level_plan(width: int, height: int, levels: int = LEVELS, pyr_scale: float = PYR_SCALE):
    """List of dicts (coarsest first) with k, scale, sigma, ksize, w, h  (SURVEY 3.4 steps 1-2)."""
    k = 0
    scale = 1.0
    while k < levels:
        scale *= pyr_scale
        if width * scale < MIN_SIZE or height * scale < MIN_SIZE:
            break
        k += 1
    top = k
    plan = []
    for k in range(top, -1, -1):
        scale = 1.0
        for _ in range(k):
            scale *= pyr_scale
        sigma = (1.0 / scale - 1.0) * 0.5
        ksize = max(cv_round(sigma * 5) | 1, 3)
        plan.append(dict(k=k, scale=scale, sigma=sigma, ksize=ksize,
                         w=cv_round(width * scale), h=cv_round(height * scale)))
    return plan


# --------------------------------------------------------------------------------------
# imgproc primitives
# --------------------------------------------------------------------------------------
def gaussian_kernel(ksize: int, sigma: float) -> np.ndarray:
    """cv::getGaussianKernel(ksize, sigma, CV_32F): fixed table for sigma<=0 and small odd
    sizes, otherwise exp(-x^2 / 2 sigma^2) normalised in double and rounded to float32."""
    small = {1: [1.0], 3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
             7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125]}
    if sigma <= 0 and ksize in small:
        return np.asarray(small[ksize], dtype=np.float32)
    if sigma <= 0:
        sigma = ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    t = np.exp(-0.5 / (sigma * sigma) * x * x)
    return (t / t.sum()).astype(np.float32)


def _reflect101(idx: np.ndarray, n: int) -> np.ndarray:
    if n == 1:
        return np.zeros_like(idx)
    period = 2 * (n - 1)
    m = np.mod(idx, period)
    return np.where(m >= n, period - m, m)


def gaussian_blur_f32(img: np.ndarray, ksize: int, sigma: float) -> np.ndarray:
    """cv::GaussianBlur on a float32 image, BORDER_REFLECT_101, separable, float32 arithmetic
    (horizontal pass first, then vertical; symmetric taps are paired like OpenCV's
    symmetric row/column filters)."""
    kern = gaussian_kernel(ksize, sigma)
    r = ksize // 2
    h, w = img.shape
    src = img.astype(np.float32)
    xs = np.arange(w)
    tmp = kern[r] * src
    for k in range(1, r + 1):
        tmp = tmp + kern[r + k] * (src[:, _reflect101(xs - k, w)] + src[:, _reflect101(xs + k, w)])
    tmp = tmp.astype(np.float32)
    ys = np.arange(h)
    out = kern[r] * tmp
    for k in range(1, r + 1):
        out = out + kern[r + k] * (tmp[_reflect101(ys - k, h), :] + tmp[_reflect101(ys + k, h), :])
    return out.astype(np.float32)


def linear_coeffs(dst_n: int, src_n: int):
    """cv::resize INTER_LINEAR per-axis tables: source index i0 and float32 weight of i0+1
    (half-pixel centres, clamped at both ends)."""
    scale = src_n / dst_n
    d = np.arange(dst_n, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    i0 = np.floor(f).astype(np.int64)
    a = (f - i0.astype(np.float32)).astype(np.float32)
    lo = i0 < 0
    i0[lo] = 0
    a[lo] = 0.0
    hi = i0 >= src_n - 1
    i0[hi] = src_n - 1
    a[hi] = 0.0
    return i0, a


def resize_linear_f32(img: np.ndarray, dst_w: int, dst_h: int) -> np.ndarray:
    """cv::resize(src, dst, Size(dst_w, dst_h), 0, 0, INTER_LINEAR) for float32 images with 1 or
    more channels (horizontal interpolation first, then vertical, all float32)."""
    src = img.astype(np.float32)
    sh, sw = src.shape[:2]
    if (sw, sh) == (dst_w, dst_h):
        return src.copy()
    xi, xa = linear_coeffs(dst_w, sw)
    yi, ya = linear_coeffs(dst_h, sh)
    xi1 = np.minimum(xi + 1, sw - 1)
    yi1 = np.minimum(yi + 1, sh - 1)
    if src.ndim == 3:
        xa_b = xa[None, :, None]
        ya_b = ya[:, None, None]
    else:
        xa_b = xa[None, :]
        ya_b = ya[:, None]
    one = np.float32(1.0)
    rows = (src[:, xi] * (one - xa_b) + src[:, xi1] * xa_b).astype(np.float32)
    out = rows[yi] * (one - ya_b) + rows[yi1] * ya_b
    return out.astype(np.float32)


# --------------------------------------------------------------------------------------
# A1a: pyramid level
# --------------------------------------------------------------------------------------
def pyramid_level(img_u8: np.ndarray, lvl: dict) -> np.ndarray:
    """u8 frame -> f32 level image: convertTo(f32), GaussianBlur(ksize, sigma) at full
    resolution, bilinear resize to (w, h)   (SURVEY 3.4 step 4)."""
    f = img_u8.astype(np.float32)
    b = gaussian_blur_f32(f, lvl["ksize"], lvl["sigma"])
    return resize_linear_f32(b, lvl["w"], lvl["h"])


# --------------------------------------------------------------------------------------
# A1b: polynomial expansion
# --------------------------------------------------------------------------------------
def poly_exp_constants(n: int = POLY_N, sigma: float = POLY_SIGMA):
    """FarnebackPrepareGaussian: float32 taps g, xg, xxg (index 0..n, symmetric /
    antisymmetric) and the four entries of inv(G) used by the expansion, in double."""
    if sigma < np.finfo(np.float32).eps:
        sigma = n * 0.3
    xs = np.arange(-n, n + 1)
    g = np.exp(-(xs * xs) / (2.0 * sigma * sigma)).astype(np.float32)
    s = 1.0 / float(np.sum(g.astype(np.float64)))
    g = (g.astype(np.float64) * s).astype(np.float32)
    xg = (xs.astype(np.float32) * g).astype(np.float32)
    xxg = ((xs * xs).astype(np.float32) * g).astype(np.float32)
    G = np.zeros((6, 6), dtype=np.float64)
    for y in xs:
        for x in xs:
            gg = np.float32(g[y + n] * g[x + n])
            G[0, 0] += float(gg)
            G[1, 1] += float(np.float32(gg * np.float32(x * x)))
            G[3, 3] += float(np.float32(gg * np.float32(x * x * x * x)))
            G[5, 5] += float(np.float32(gg * np.float32(x * x * y * y)))
    G[2, 2] = G[0, 3] = G[0, 4] = G[3, 0] = G[4, 0] = G[1, 1]
    G[4, 4] = G[3, 3]
    G[3, 4] = G[4, 3] = G[5, 5]
    inv = np.linalg.inv(G)
    return dict(g=g[n:].copy(), xg=xg[n:].copy(), xxg=xxg[n:].copy(),
                ig11=float(inv[1, 1]), ig03=float(inv[0, 3]),
                ig33=float(inv[3, 3]), ig55=float(inv[5, 5]))


def poly_exp(img: np.ndarray, n: int = POLY_N, sigma: float = POLY_SIGMA, acc=np.float64) -> np.ndarray:
    """FarnebackPolyExp: separable (2n+1)-tap correlation with replicate border.  Vertical pass
    in float32 (3 rows: g, xg, xxg along y), horizontal pass with ``acc`` accumulators.
    Returns f32 [h, w, 5] = (d/dy, d/dx, yy, xx, xy)   (SURVEY 3.4 step 5)."""
    c = poly_exp_constants(n, sigma)
    g, xg, xxg = c["g"], c["xg"], c["xxg"]
    src = img.astype(np.float32)
    h, w = src.shape
    ys = np.arange(h)
    r0 = src * g[0]
    r1 = np.zeros_like(src)
    r2 = np.zeros_like(src)
    for k in range(1, n + 1):
        up = src[np.maximum(ys - k, 0)]
        dn = src[np.minimum(ys + k, h - 1)]
        p = up + dn
        r0 = (r0 + g[k] * p).astype(np.float32)
        r1 = (r1 + xg[k] * (dn - up)).astype(np.float32)
        r2 = (r2 + xxg[k] * p).astype(np.float32)
    xs = np.arange(w)
    a0, a1, a2 = r0.astype(acc), r1.astype(acc), r2.astype(acc)
    b1 = a0 * acc(g[0])
    b3 = a1 * acc(g[0])
    b5 = a2 * acc(g[0])
    b2 = np.zeros_like(b1)
    b4 = np.zeros_like(b1)
    b6 = np.zeros_like(b1)
    for k in range(1, n + 1):
        xl = np.maximum(xs - k, 0)
        xr = np.minimum(xs + k, w - 1)
        tg = a0[:, xr] + a0[:, xl]
        b1 = b1 + tg * acc(g[k])
        b4 = b4 + tg * acc(xxg[k])
        b2 = b2 + (a0[:, xr] - a0[:, xl]) * acc(xg[k])
        b3 = b3 + (a1[:, xr] + a1[:, xl]) * acc(g[k])
        b6 = b6 + (a1[:, xr] - a1[:, xl]) * acc(xg[k])
        b5 = b5 + (a2[:, xr] + a2[:, xl]) * acc(g[k])
    ig11, ig03, ig33, ig55 = (acc(c[k]) for k in ("ig11", "ig03", "ig33", "ig55"))
    out = np.empty((h, w, 5), dtype=np.float32)
    out[..., 0] = b3 * ig11
    out[..., 1] = b2 * ig11
    out[..., 2] = b1 * ig03 + b5 * ig33
    out[..., 3] = b1 * ig03 + b4 * ig33
    out[..., 4] = b6 * ig55
    return out


# --------------------------------------------------------------------------------------
# A1c: update matrices
# --------------------------------------------------------------------------------------
def border_scale(w: int, h: int) -> np.ndarray:
    """Per-pixel attenuation: product of BORDER_ATTEN[distance to each edge] for distances < 5."""
    def axis(n):
        s = np.ones(n, dtype=np.float32)
        i = np.arange(n)
        near = i < 5
        s[near] = s[near] * BORDER_ATTEN[i[near]]
        far = i >= n - 5
        s[far] = s[far] * BORDER_ATTEN[n - i[far] - 1]
        return s
    return (axis(w)[None, :] * np.ones((h, 1), np.float32)) * axis(h)[:, None]


def update_matrices(R0: np.ndarray, R1: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """FarnebackUpdateMatrices: bilinear gather of R1 at (x+dx, y+dy) when the 2x2 footprint is
    inside the image, else the R0-only fallback; 5 px border attenuation; returns
    f32 [h, w, 5] = (G11, G12, G22, h1, h2)   (SURVEY 3.4 step 6).  float32 throughout."""
    f32 = np.float32
    h, w = flow.shape[:2]
    ys, xs = np.mgrid[0:h, 0:w]
    dx = flow[..., 0].astype(f32)
    dy = flow[..., 1].astype(f32)
    fx = (xs.astype(f32) + dx).astype(f32)
    fy = (ys.astype(f32) + dy).astype(f32)
    x1 = np.floor(fx).astype(np.int64)
    y1 = np.floor(fy).astype(np.int64)
    fx = (fx - x1.astype(f32)).astype(f32)
    fy = (fy - y1.astype(f32)).astype(f32)
    inside = (x1 >= 0) & (x1 < w - 1) & (y1 >= 0) & (y1 < h - 1)
    xc = np.clip(x1, 0, w - 2) if w > 1 else np.zeros_like(x1)
    yc = np.clip(y1, 0, h - 2) if h > 1 else np.zeros_like(y1)
    one = f32(1.0)
    a00 = ((one - fx) * (one - fy)).astype(f32)
    a01 = (fx * (one - fy)).astype(f32)
    a10 = ((one - fx) * fy).astype(f32)
    a11 = (fx * fy).astype(f32)
    x2 = np.minimum(xc + 1, w - 1)
    y2 = np.minimum(yc + 1, h - 1)

    def samp(c):
        p = R1[..., c]
        return (a00 * p[yc, xc] + a01 * p[yc, x2] + a10 * p[y2, xc] + a11 * p[y2, x2]).astype(f32)

    r2 = np.where(inside, samp(0), f32(0)).astype(f32)
    r3 = np.where(inside, samp(1), f32(0)).astype(f32)
    r4 = np.where(inside, (R0[..., 2] + samp(2)) * f32(0.5), R0[..., 2]).astype(f32)
    r5 = np.where(inside, (R0[..., 3] + samp(3)) * f32(0.5), R0[..., 3]).astype(f32)
    r6 = np.where(inside, (R0[..., 4] + samp(4)) * f32(0.25), R0[..., 4] * f32(0.5)).astype(f32)
    r2 = ((R0[..., 0] - r2) * f32(0.5)).astype(f32)
    r3 = ((R0[..., 1] - r3) * f32(0.5)).astype(f32)
    r2 = (r2 + (r4 * dy + r6 * dx)).astype(f32)
    r3 = (r3 + (r6 * dy + r5 * dx)).astype(f32)
    sc = border_scale(w, h)
    r2, r3, r4, r5, r6 = (np.asarray(v * sc, dtype=f32) for v in (r2, r3, r4, r5, r6))
    M = np.empty((h, w, 5), dtype=f32)
    M[..., 0] = r4 * r4 + r6 * r6
    M[..., 1] = (r4 + r5) * r6
    M[..., 2] = r5 * r5 + r6 * r6
    M[..., 3] = r4 * r2 + r6 * r3
    M[..., 4] = r6 * r2 + r5 * r3
    return M


# --------------------------------------------------------------------------------------
# A1d: box blur + 2x2 solve
# --------------------------------------------------------------------------------------
def box_sum(M: np.ndarray, win: int = WINSIZE, acc=np.float64) -> np.ndarray:
    """win x win box *sum* with replicate border (vertical then horizontal), ``acc`` sums."""
    m = win // 2
    h, w = M.shape[:2]
    a = M.astype(acc)
    ys = np.arange(h)
    v = np.zeros_like(a)
    for d in range(-m, m + 1):
        v = v + a[np.clip(ys + d, 0, h - 1)]
    xs = np.arange(w)
    s = np.zeros_like(a)
    for d in range(-m, m + 1):
        s = s + v[:, np.clip(xs + d, 0, w - 1)]
    return s


def blur_solve(M: np.ndarray, win: int = WINSIZE, acc=np.float64) -> np.ndarray:
    """FarnebackUpdateFlow_Blur without the interleaved matrix refresh: box mean of M, then
    the regularised 2x2 solve   (SURVEY 3.4 step 7)."""
    s = box_sum(M, win, acc)
    scale = acc(1.0 / (win * win))
    g11 = s[..., 0] * scale
    g12 = s[..., 1] * scale
    g22 = s[..., 2] * scale
    h1 = s[..., 3] * scale
    h2 = s[..., 4] * scale
    idet = acc(1.0) / (g11 * g22 - g12 * g12 + acc(1e-3))
    flow = np.empty(M.shape[:2] + (2,), dtype=np.float32)
    flow[..., 0] = (g11 * h2 - g12 * h1) * idet
    flow[..., 1] = (g22 * h1 - g12 * h2) * idet
    return flow


# --------------------------------------------------------------------------------------
# A1e: flow up-sampling between levels
# --------------------------------------------------------------------------------------
def upsample_flow(flow: np.ndarray, w: int, h: int, pyr_scale: float = PYR_SCALE) -> np.ndarray:
    """resize(prevFlow, (w, h), INTER_LINEAR) * (1 / pyr_scale)   (SURVEY 3.4 step 3)."""
    return (resize_linear_f32(flow, w, h) * np.float32(1.0 / pyr_scale)).astype(np.float32)


# --------------------------------------------------------------------------------------
# full algorithm
# --------------------------------------------------------------------------------------
def frame_expansion(img_u8: np.ndarray, acc=np.float64):
    """Per-frame half of the algorithm: list (coarsest first) of (level dict, I_k, R_k)."""
    h, w = img_u8.shape
    out = []
    for lvl in level_plan(w, h):
        I = pyramid_level(img_u8, lvl)
        out.append((lvl, I, poly_exp(I, acc=acc)))
    return out


def farneback(prev_u8: np.ndarray, next_u8: np.ndarray, acc=np.float64, iterations: int = ITERATIONS,
              return_levels: bool = False):
    """Dense flow f32 [H, W, 2] (channel 0 = x displacement u, channel 1 = y displacement v)."""
    assert prev_u8.shape == next_u8.shape and prev_u8.ndim == 2
    e0 = frame_expansion(prev_u8, acc)
    e1 = frame_expansion(next_u8, acc)
    flow = None
    per_level = []
    for (lvl, _, R0), (_, _, R1) in zip(e0, e1):
        if flow is None:
            flow = np.zeros((lvl["h"], lvl["w"], 2), dtype=np.float32)
        else:
            flow = upsample_flow(flow, lvl["w"], lvl["h"])
        for _ in range(iterations):
            M = update_matrices(R0, R1, flow)
            flow = blur_solve(M, WINSIZE, acc)
        per_level.append(flow)
    return (flow, per_level) if return_levels else flow
