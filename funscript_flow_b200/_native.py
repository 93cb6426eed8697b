"""ctypes binding of libffb.so (the C ABI in include/ffb.h).

There is deliberately no fallback: if the CUDA library is missing or no device is usable the
import of a context raises.  (`load(path)` accepts an explicit path only so the CPU test-suite
can point it at the g++-built emulation of the same sources under tests/emu/.)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libffb.so")

K_NAMES = ["pyramid", "polyexp", "upsample", "flow_iter", "divmag", "radial", "small", "preprocess"]
K_FLOW_ITER = 3


class FFBError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ffb error {code}: {msg}")
        self.code = code


_libs = {}


def load(path: Optional[str] = None) -> C.CDLL:
    path = os.path.abspath(path or DEFAULT_LIB)
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise FileNotFoundError(
            f"{path} not found: build it with `python -m funscript_flow_b200.build` (needs nvcc). "
            "There is no CPU fallback.")
    lib = C.CDLL(path)
    vp, i32, u8p, f32p, f64p, i32p, sz = (C.c_void_p, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_float),
                                          C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_size_t)
    sig = {
        "ffb_version": (i32, []),
        "ffb_device_count": (i32, [i32p]),
        "ffb_device_pci_bus_id": (i32, [i32, C.c_char_p, i32]),
        "ffb_device_name": (i32, [i32, C.c_char_p, i32]),
        "ffb_last_error": (C.c_char_p, [vp]),
        "ffb_create": (i32, [i32, C.POINTER(vp)]),
        "ffb_destroy": (None, [vp]),
        "ffb_host_alloc": (i32, [C.POINTER(vp), sz]),
        "ffb_host_free": (i32, [vp]),
        "ffb_configure": (i32, [vp, i32, i32, i32, i32]),
        "ffb_get_geometry": (i32, [vp, i32p, i32p, i32p, i32p]),
        "ffb_alloc_counts": (i32, [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
        "ffb_bracket_begin": (i32, [vp, i32, C.c_double]),
        "ffb_bracket_begin_shard": (i32, [vp, i32, C.c_double, i32]),
        "ffb_bracket_phase1_finish": (i32, [vp, i32p, i32p, i32p, f32p, f32p, u8p]),
        "ffb_bracket_radial": (i32, [vp, i32p, i32p, i32, i32, f64p, f64p]),
        "ffb_bracket_push": (i32, [vp, vp, i32, sz, sz]),
        "ffb_bracket_finish": (i32, [vp, i32p, f64p, u8p, i32p, i32p, f32p, f32p, f64p]),
        "ffb_sync": (i32, [vp]),
        "ffb_chain_after": (i32, [vp, vp]),
        "ffb_bracket_abort": (i32, [vp]),
        "ffb_bracket_get_flow": (i32, [vp, i32, f32p]),
        "ffb_flow_ring_size": (i32, [vp]),
        "ffb_farneback": (i32, [vp, u8p, u8p, i32, i32, sz, f32p]),
        "ffb_max_divergence": (i32, [vp, f32p, i32, i32, i32p, i32p, f32p]),
        "ffb_mean_magnitude": (i32, [vp, f32p, i32, i32, f32p]),
        "ffb_radial_motion": (i32, [vp, f32p, i32, i32, C.c_double, C.c_double, i32, i32, f64p]),
        "ffb_level_plan": (i32, [i32, i32, i32p, i32p, i32p, i32p, f64p]),
        "ffb_stage_pyramid": (i32, [vp, u8p, i32, i32, sz, i32, f32p]),
        "ffb_stage_polyexp": (i32, [vp, f32p, i32, i32, f32p]),
        "ffb_stage_update_matrices": (i32, [vp, f32p, f32p, f32p, i32, i32, f32p]),
        "ffb_stage_flow_iter": (i32, [vp, f32p, f32p, f32p, i32, i32, f32p]),
        "ffb_stage_upsample_flow": (i32, [vp, f32p, i32, i32, i32, i32, f32p]),
        "ffb_preprocess_configure": (i32, [vp, i32, i32, i32]),
        "ffb_bracket_push_bgr": (i32, [vp, vp, i32, sz, sz]),
        "ffb_stage_preprocess": (i32, [vp, u8p, i32, i32, sz, i32, u8p]),
        "ffb_preprocess_configure_window": (i32, [vp] + [i32] * 8),
        "ffb_stage_preprocess_window": (i32, [vp, u8p, i32, i32, sz] + [i32] * 6 + [u8p]),
        "ffb_profile": (i32, [vp, i32]),
        "ffb_profile_reset": (i32, [vp]),
        "ffb_kernel_stats": (i32, [vp, i32, C.POINTER(C.c_int64), f64p, f64p]),
        "ffb_flow_iter_level_stats": (i32, [vp, i32, C.POINTER(C.c_int64), f64p, f64p]),
        "ffb_launch_count": (C.c_int64, [vp]),
        "ffb_kernel_name": (C.c_char_p, [i32]),
        "ffb_timer_mark": (i32, [vp, i32]),
        "ffb_timer_elapsed": (i32, [vp, i32, i32, f64p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)   # AttributeError here = the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    lib._ffb_symbols = sorted(sig)
    _libs[path] = lib
    return lib


def device_count(lib_path: Optional[str] = None) -> int:
    n = C.c_int32(0)
    load(lib_path).ffb_device_count(C.byref(n))
    return int(n.value)


def device_pci_bus_id(device: int = 0, lib_path: Optional[str] = None) -> Optional[str]:
    buf = C.create_string_buffer(32)
    if load(lib_path).ffb_device_pci_bus_id(int(device), buf, 32) != 0:
        return None
    return buf.value.decode().lower()


def device_name(device: int = 0, lib_path: Optional[str] = None) -> Optional[str]:
    buf = C.create_string_buffer(256)
    if load(lib_path).ffb_device_name(int(device), buf, 256) != 0:
        return None
    return buf.value.decode()


def level_plan(width: int, height: int, lib_path: Optional[str] = None):
    lib = load(lib_path)
    n = C.c_int32(0)
    w = (C.c_int32 * 4)()
    h = (C.c_int32 * 4)()
    ks = (C.c_int32 * 4)()
    sg = (C.c_double * 4)()
    rc = lib.ffb_level_plan(width, height, C.byref(n), w, h, ks, sg)
    if rc:
        raise FFBError(rc, "ffb_level_plan")
    return [dict(w=w[i], h=h[i], ksize=ks[i], sigma=sg[i]) for i in range(n.value)]


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class PinnedBuffer:
    """Page-locked host memory exposed as a NumPy uint8 array (frames DMA straight out of it)."""

    def __init__(self, shape, lib_path: Optional[str] = None):
        self._lib = load(lib_path)
        self.nbytes = int(np.prod(shape))
        p = C.c_void_p()
        rc = self._lib.ffb_host_alloc(C.byref(p), self.nbytes)
        if rc:
            raise FFBError(rc, "ffb_host_alloc")
        self._ptr = p
        self.array = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(self.nbytes,)).reshape(shape)

    def free(self):
        if self._ptr:
            self.array = None
            self._lib.ffb_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class FlowContext:
    """One GPU, one host thread.  Thin, allocation-free wrappers over the C ABI."""

    def __init__(self, device: int = 0, lib_path: Optional[str] = None):
        self._lib = load(lib_path)
        self.lib_path = lib_path
        h = C.c_void_p()
        rc = self._lib.ffb_create(int(device), C.byref(h))
        if rc:
            raise FFBError(rc, (self._lib.ffb_last_error(None) or b"").decode())
        self._h = h
        self.device = device

    # -- plumbing
    def _ck(self, rc: int):
        if rc:
            raise FFBError(rc, (self._lib.ffb_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ffb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- geometry / bracket API
    def configure(self, width: int, height: int, batch_frames: int = 16, max_bracket_pairs: int = 4096):
        """Incremental on the C side: unchanged or smaller requests touch nothing on the device."""
        self._ck(self._lib.ffb_configure(self._h, int(width), int(height), int(batch_frames), int(max_bracket_pairs)))

    @property
    def geometry(self):
        """(width, height, batch_frames, max_bracket_pairs) of the last successful configure, None before it."""
        v = [C.c_int32(0) for _ in range(4)]
        self._ck(self._lib.ffb_get_geometry(self._h, *(C.byref(x) for x in v)))
        geo = tuple(int(x.value) for x in v)
        return geo if geo[0] > 0 else None

    def alloc_counts(self):
        """(cudaMalloc calls, cudaHostAlloc calls) made for this context so far."""
        d, h = C.c_int64(0), C.c_int64(0)
        self._ck(self._lib.ffb_alloc_counts(self._h, C.byref(d), C.byref(h)))
        return int(d.value), int(h.value)

    def bracket_begin(self, pov_mode: bool = False, cut_threshold: float = 7.0):
        self._ck(self._lib.ffb_bracket_begin(self._h, int(bool(pov_mode)), float(cut_threshold)))

    # -- frame-range shard of a bracket: phase 1 (flows, raw centres, cut test), exchange of the raw centres
    #    with the neighbouring shards on the host, phase 2 (radial scalars)
    def bracket_begin_shard(self, shard_pairs: int, pov_mode: bool = False, cut_threshold: float = 7.0):
        self._ck(self._lib.ffb_bracket_begin_shard(self._h, int(bool(pov_mode)), float(cut_threshold), int(shard_pairs)))

    def bracket_phase1_finish(self):
        m = self.geometry[3]
        out = dict(cx=np.empty(m, np.int32), cy=np.empty(m, np.int32), val=np.empty(m, np.float32),
                   mean_mag=np.empty(m, np.float32), cut=np.empty(m, np.uint8))
        n = C.c_int32(0)
        i32p = C.POINTER(C.c_int32)
        self._ck(self._lib.ffb_bracket_phase1_finish(self._h, C.byref(n), out["cx"].ctypes.data_as(i32p),
                                                     out["cy"].ctypes.data_as(i32p), _f32(out["val"]), _f32(out["mean_mag"]),
                                                     _u8(out["cut"])))
        k = n.value
        res = {key: v[:k].copy() for key, v in out.items()}
        res["cut"] = res["cut"].astype(bool)
        res["n_pairs"] = k
        return res

    def bracket_radial(self, cx_ext, cy_ext, first: int, n_pairs: int):
        """cx_ext / cy_ext: raw centres of the shard's pairs and of up to 6 neighbours on each side (clipped to the
        bracket), the shard's first pair at index `first`.  Returns (scalar f64[n], centers f64[n, 2])."""
        cx_ext = np.ascontiguousarray(cx_ext, np.int32)
        cy_ext = np.ascontiguousarray(cy_ext, np.int32)
        assert cx_ext.shape == cy_ext.shape and cx_ext.ndim == 1
        scalar = np.empty(max(n_pairs, 1), np.float64)
        centers = np.empty((max(n_pairs, 1), 2), np.float64)
        i32p, f64p = C.POINTER(C.c_int32), C.POINTER(C.c_double)
        self._ck(self._lib.ffb_bracket_radial(self._h, cx_ext.ctypes.data_as(i32p), cy_ext.ctypes.data_as(i32p), int(cx_ext.size),
                                              int(first), scalar.ctypes.data_as(f64p), centers.ctypes.data_as(f64p)))
        return scalar[:n_pairs].copy(), centers[:n_pairs].copy()

    def bracket_push(self, frames: np.ndarray):
        """frames: uint8 [n, H, W] (or [H, W]); the last axis must be contiguous."""
        if frames.ndim == 2:
            frames = frames[None]
        assert frames.dtype == np.uint8 and frames.ndim == 3 and frames.strides[2] == 1
        n = frames.shape[0]
        stride = frames.strides[0] if n > 1 else frames.strides[1] * frames.shape[1]
        self._ck(self._lib.ffb_bracket_push(self._h, frames.ctypes.data, n, frames.strides[1], stride))

    def bracket_push_ptr(self, ptr: int, n: int, pitch: int, frame_stride: int):
        """Raw pointer variant (device pointers, e.g. torch_tensor.data_ptr())."""
        self._ck(self._lib.ffb_bracket_push(self._h, C.c_void_p(ptr), int(n), int(pitch), int(frame_stride)))

    def preprocess_configure(self, src_width: int, src_height: int, vr_mode: bool = False):
        """Source geometry of decoded BGR frames for bracket_push_bgr (output is always 256x256)."""
        self._ck(self._lib.ffb_preprocess_configure(self._h, int(src_width), int(src_height), int(bool(vr_mode))))

    def preprocess_configure_window(self, src_width: int, src_height: int, target, window):
        """General form (row N4): resize to target=(w, h), keep window=(x, y, w, h) of the resized frame."""
        self._ck(self._lib.ffb_preprocess_configure_window(self._h, int(src_width), int(src_height), int(target[0]), int(target[1]),
                                                           *(int(v) for v in window)))

    def bracket_push_bgr(self, frames: np.ndarray):
        """frames: uint8 [n, H, W, 3] (or [H, W, 3]) BGR as decoded; resized + gray-converted on the GPU."""
        if frames.ndim == 3:
            frames = frames[None]
        assert frames.dtype == np.uint8 and frames.ndim == 4 and frames.shape[3] == 3
        assert frames.strides[3] == 1 and frames.strides[2] == 3
        n = frames.shape[0]
        stride = frames.strides[0] if n > 1 else frames.strides[1] * frames.shape[1]
        self._ck(self._lib.ffb_bracket_push_bgr(self._h, frames.ctypes.data, n, frames.strides[1], stride))

    def stage_preprocess(self, bgr: np.ndarray, vr_mode: bool = False) -> np.ndarray:
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
        h, w = bgr.shape[:2]
        out = np.empty((256, 256), np.uint8)
        self._ck(self._lib.ffb_stage_preprocess(self._h, _u8(bgr), w, h, w * 3, int(bool(vr_mode)), _u8(out)))
        return out

    def stage_preprocess_window(self, bgr: np.ndarray, target, window) -> np.ndarray:
        bgr = np.ascontiguousarray(bgr, dtype=np.uint8)
        h, w = bgr.shape[:2]
        out = np.empty((int(window[3]), int(window[2])), np.uint8)
        self._ck(self._lib.ffb_stage_preprocess_window(self._h, _u8(bgr), w, h, w * 3, int(target[0]), int(target[1]),
                                                       *(int(v) for v in window), _u8(out)))
        return out

    def bracket_abort(self):
        """Drop the open bracket (error path / cancel); the context can be configured again afterwards."""
        self._ck(self._lib.ffb_bracket_abort(self._h))

    def bracket_finish(self):
        m = self.geometry[3]
        out = dict(scalar=np.empty(m, np.float64), cut=np.empty(m, np.uint8), cx=np.empty(m, np.int32),
                   cy=np.empty(m, np.int32), val=np.empty(m, np.float32), mean_mag=np.empty(m, np.float32),
                   centers=np.empty((m, 2), np.float64))
        n = C.c_int32(0)
        self._ck(self._lib.ffb_bracket_finish(
            self._h, C.byref(n), out["scalar"].ctypes.data_as(C.POINTER(C.c_double)), _u8(out["cut"]),
            out["cx"].ctypes.data_as(C.POINTER(C.c_int32)), out["cy"].ctypes.data_as(C.POINTER(C.c_int32)),
            _f32(out["val"]), _f32(out["mean_mag"]), out["centers"].ctypes.data_as(C.POINTER(C.c_double))))
        k = n.value
        res = {key: v[:k].copy() for key, v in out.items()}
        res["cut"] = res["cut"].astype(bool)
        res["n_pairs"] = k
        return res

    def chain_after(self, earlier: "FlowContext"):
        """Kernels queued on this context from now on start after everything `earlier` (same device) has queued."""
        self._ck(self._lib.ffb_chain_after(self._h, earlier._h))

    def sync(self):
        self._ck(self._lib.ffb_sync(self._h))

    def get_flow(self, pair: int) -> np.ndarray:
        w, h = self.geometry[:2]
        out = np.empty((h, w, 2), np.float32)
        self._ck(self._lib.ffb_bracket_get_flow(self._h, int(pair), _f32(out)))
        return out

    @property
    def flow_ring_size(self) -> int:
        return int(self._lib.ffb_flow_ring_size(self._h))

    # -- per-call functions
    def farneback(self, prev: np.ndarray, nxt: np.ndarray) -> np.ndarray:
        prev = np.ascontiguousarray(prev, dtype=np.uint8)
        nxt = np.ascontiguousarray(nxt, dtype=np.uint8)
        assert prev.shape == nxt.shape and prev.ndim == 2
        h, w = prev.shape
        out = np.empty((h, w, 2), np.float32)
        self._ck(self._lib.ffb_farneback(self._h, _u8(prev), _u8(nxt), w, h, w, _f32(out)))
        return out

    def max_divergence(self, flow: np.ndarray):
        flow = np.ascontiguousarray(flow, dtype=np.float32)
        h, w = flow.shape[:2]
        x, y, v = C.c_int32(), C.c_int32(), C.c_float()
        self._ck(self._lib.ffb_max_divergence(self._h, _f32(flow), w, h, C.byref(x), C.byref(y), C.byref(v)))
        return x.value, y.value, np.float32(v.value)

    def mean_magnitude(self, flow: np.ndarray) -> np.float32:
        flow = np.ascontiguousarray(flow, dtype=np.float32)
        h, w = flow.shape[:2]
        v = C.c_float()
        self._ck(self._lib.ffb_mean_magnitude(self._h, _f32(flow), w, h, C.byref(v)))
        return np.float32(v.value)

    def radial_motion(self, flow: np.ndarray, center, is_cut: bool, pov_mode: bool = False) -> float:
        flow = np.ascontiguousarray(flow, dtype=np.float32)
        h, w = flow.shape[:2]
        v = C.c_double()
        self._ck(self._lib.ffb_radial_motion(self._h, _f32(flow), w, h, float(center[0]), float(center[1]),
                                             int(bool(is_cut)), int(bool(pov_mode)), C.byref(v)))
        return float(v.value)

    # -- stage hooks
    def stage_pyramid(self, img: np.ndarray, level_k: int) -> np.ndarray:
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape
        plan = level_plan(w, h, self._lib._name)
        lv = plan[len(plan) - 1 - level_k]
        out = np.empty((lv["h"], lv["w"]), np.float32)
        self._ck(self._lib.ffb_stage_pyramid(self._h, _u8(img), w, h, w, int(level_k), _f32(out)))
        return out

    def stage_polyexp(self, img: np.ndarray) -> np.ndarray:
        """f32 [h, w] -> f32 [h, w, 5] (converted from the kernel's 5 planes)."""
        img = np.ascontiguousarray(img, dtype=np.float32)
        h, w = img.shape
        out = np.empty((5, h, w), np.float32)
        self._ck(self._lib.ffb_stage_polyexp(self._h, _f32(img), w, h, _f32(out)))
        return np.ascontiguousarray(out.transpose(1, 2, 0))

    @staticmethod
    def _planes(R: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(np.asarray(R, np.float32).transpose(2, 0, 1))

    def stage_update_matrices(self, R0, R1, flow) -> np.ndarray:
        h, w = R0.shape[:2]
        p0, p1 = self._planes(R0), self._planes(R1)
        fl = None if flow is None else np.ascontiguousarray(flow, np.float32)
        out = np.empty((5, h, w), np.float32)
        self._ck(self._lib.ffb_stage_update_matrices(self._h, _f32(p0), _f32(p1), None if fl is None else _f32(fl),
                                                     w, h, _f32(out)))
        return np.ascontiguousarray(out.transpose(1, 2, 0))

    def stage_flow_iter(self, R0, R1, flow_in) -> np.ndarray:
        h, w = R0.shape[:2]
        p0, p1 = self._planes(R0), self._planes(R1)
        fl = None if flow_in is None else np.ascontiguousarray(flow_in, np.float32)
        out = np.empty((h, w, 2), np.float32)
        self._ck(self._lib.ffb_stage_flow_iter(self._h, _f32(p0), _f32(p1), None if fl is None else _f32(fl),
                                               w, h, _f32(out)))
        return out

    def stage_upsample_flow(self, flow_c: np.ndarray, w: int, h: int) -> np.ndarray:
        fc = np.ascontiguousarray(flow_c, np.float32)
        hc, wc = fc.shape[:2]
        out = np.empty((h, w, 2), np.float32)
        self._ck(self._lib.ffb_stage_upsample_flow(self._h, _f32(fc), wc, hc, w, h, _f32(out)))
        return out

    # -- instrumentation
    def profile(self, enable: bool):
        self._ck(self._lib.ffb_profile(self._h, int(enable)))

    def profile_reset(self):
        self._ck(self._lib.ffb_profile_reset(self._h))

    def kernel_stats(self):
        out = {}
        for kid, name in enumerate(K_NAMES):
            n, ms, by = C.c_int64(), C.c_double(), C.c_double()
            self._ck(self._lib.ffb_kernel_stats(self._h, kid, C.byref(n), C.byref(ms), C.byref(by)))
            out[name] = dict(launches=int(n.value), ms=float(ms.value), alg_bytes=float(by.value))
        return out

    def flow_iter_level_stats(self):
        """k_flow_iter launches / device ms / algorithmic bytes per pyramid level k (0 = full resolution)."""
        out = {}
        for k in range(4):
            n, ms, by = C.c_int64(), C.c_double(), C.c_double()
            self._ck(self._lib.ffb_flow_iter_level_stats(self._h, k, C.byref(n), C.byref(ms), C.byref(by)))
            if n.value:
                out[k] = dict(launches=int(n.value), ms=float(ms.value), alg_bytes=float(by.value))
        return out

    def timer_mark(self, slot: int):
        self._ck(self._lib.ffb_timer_mark(self._h, int(slot)))

    def timer_elapsed_ms(self, slot_from: int, slot_to: int) -> float:
        ms = C.c_double()
        self._ck(self._lib.ffb_timer_elapsed(self._h, int(slot_from), int(slot_to), C.byref(ms)))
        return float(ms.value)

    @property
    def launch_count(self) -> int:
        return int(self._lib.ffb_launch_count(self._h))
