"""Generate the golden vectors under tests/golden/ from the REFERENCE ITSELF.

Run in the build container only (needs /root/reference and cv2):

    python tests/golden/make_golden.py

The reference's functions are AST-loaded from /root/reference/FunscriptFlow.pyw (oracle/ref_loader.py)
and executed unchanged; cv2 is the installed opencv-python-headless (4.13.0 here; the reference's
lock file pins 4.11.0.86 -- version skew that cannot be checked offline).
Nothing from the reference is copied: only inputs and the outputs it produced are stored.

Files written
  kat_motion.json       known answers of max_divergence / radial_motion_weighted on seeded random flow
  pairs.npz             small frame pairs + the reference precompute_flow_info() outputs (incl. cv2 flow)
  bracket.npz           a 28-frame 128x96 clip with one hard cut + the reference bracket loop's outputs
  postproc.json         per-pair scalar series + the actions the reference's post-processing emits
  video_c1.json         actions of the reference's process_video() on the C1 clip (640x360, 300 frames, FFV1)
  video_multibracket.json   the same for a clip cut into three brackets (batch_size=30), prefetch race neutralised
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

from funscript_flow_b200.synth import ClipGenerator, ClipSpec, make_clip  # noqa: E402
from oracle import ref_loader  # noqa: E402


def write_video(path, frames, fps):
    h, w = frames[0].shape
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), fps, (w, h), True)
    assert vw.isOpened()
    for f in frames:
        vw.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
    vw.release()


def main():
    ref = ref_loader.load(serial_pools=True)
    meta = {"cv2": cv2.__version__, "numpy": np.__version__}

    # ---- 1. known answers on seeded random flow fields
    kat = {"meta": meta, "cases": []}
    for seed, shape in [(0, (8, 10)), (1, (33, 47)), (2, (64, 64))]:
        flow = np.random.default_rng(seed).standard_normal(shape + (2,)).astype(np.float32)
        x, y, v = ref.max_divergence(flow)
        case = {"seed": seed, "shape": list(shape), "max_divergence": [int(x), int(y), float(v)], "radial": []}
        h, w = shape
        for cx, cy in [(w * 0.45, h * 0.4), (float(w // 2), float(h // 2)), (0.0, 0.0), (w - 1.0, h - 1.0)]:
            for pov in (False, True):
                case["radial"].append({"center": [cx, cy], "pov": pov,
                                       "value": float(ref.radial_motion_weighted(flow, [cx, cy], False, pov))})
        case["radial_cut"] = float(ref.radial_motion_weighted(flow, [1.0, 1.0], True, False))
        kat["cases"].append(case)
    # analytic pure expansion
    ys, xs = np.mgrid[0:360, 0:640].astype(np.float32)
    flow = np.stack([0.01 * (xs - 352), 0.01 * (ys - 162)], axis=-1).astype(np.float32)
    x, y, v = ref.max_divergence(flow)
    kat["expansion_640x360"] = {"radial": float(ref.radial_motion_weighted(flow, [352, 162], False, False)),
                                "max_divergence": [int(x), int(y), float(v)]}
    json.dump(kat, open(os.path.join(HERE, "kat_motion.json"), "w"), indent=1)

    # ---- 2. small frame pairs through the reference's precompute_flow_info (CPU backend)
    out = {}
    specs = [("a", 256, 256, 5, 12.0, 0.3), ("b", 200, 136, 6, 10.0, 0.25), ("c", 96, 64, 7, 8.0, 0.2)]
    for name, w, h, seed, period, amp in specs:
        clip = make_clip(w, h, 6, seed=seed, period=period, amplitude=amp)
        p0, p1 = clip[2], clip[3]
        for pov in (False, True):
            info = ref.precompute_flow_info(p0, p1, {"backend": "CPU", "pov_mode": pov})
            tag = f"{name}_pov" if pov else name
            out[f"{tag}_center"] = np.array(info["pos_center"], dtype=np.int64)
            out[f"{tag}_val"] = np.float32(info["val_pos"])
            out[f"{tag}_cut"] = np.bool_(info["cut"])
            out[f"{tag}_mean_mag"] = np.float32(info["mean_mag"])
        out[f"{name}_p0"], out[f"{name}_p1"] = p0, p1
        out[f"{name}_flow"] = info["flow"]
    np.savez_compressed(os.path.join(HERE, "pairs.npz"), **out)

    # ---- 3. one bracket with a hard cut through the reference's bracket-loop arithmetic
    spec = ClipSpec(128, 96, 28, seed=11, amplitude=0.3, period=9.0, cuts=(15,))
    frames = ClipGenerator(spec).stack()
    params = {"backend": "CPU", "threads": 1, "pov_mode": False, "cut_threshold": 2.0}
    infos = [ref.precompute_wrapper(p, params) for p in zip(frames[:-1], frames[1:])]
    centers = []
    for j in range(len(infos)):   # the reference's own smoothing arithmetic, driven like F:1203-1214
        lst = [infos[j]["pos_center"]]
        for i in range(1, 7):
            if j - i >= 0:
                lst.append(infos[j - i]["pos_center"])
            if j + i < len(infos):
                lst.append(infos[j + i]["pos_center"])
        centers.append(np.mean(np.array(lst), axis=0))
    scal = [ref.radial_motion_weighted(i["flow"], centers[j], i["cut"], False) for j, i in enumerate(infos)]
    # The clip is only a valid golden for the CENTRES when every implementation must find the same argmax: the
    # reference's own top-1 / top-2 gap of |div| has to be far above the ~1e-6 px by which flow fields of different
    # (correct) implementations differ, and the NumPy restatement has to agree with cv2 on every centre.  The tests
    # then assert the centres, the smoothed centres and the scalars unconditionally.
    from oracle import farneback_np as fb, motion_np as mo
    gaps = []
    for j, i in enumerate(infos):
        d = np.abs(mo.divergence_field(i["flow"]))
        flat = int(np.argmax(d))
        top1 = float(d.flat[flat])
        d.flat[flat] = -1.0
        gaps.append(top1 - float(d.max()))
        ox, oy, _ = mo.max_divergence(fb.farneback(frames[j], frames[j + 1]))
        assert (int(ox), int(oy)) == tuple(int(v) for v in i["pos_center"]), ("oracle and cv2 disagree on the centre of pair", j)
    assert min(gaps) >= 3e-4, ("bracket clip lost its argmax margin", min(gaps))
    np.savez_compressed(
        os.path.join(HERE, "bracket.npz"), frames=frames, cut_threshold=np.float64(2.0),
        centers_raw=np.array([i["pos_center"] for i in infos], dtype=np.int64), centers=np.array(centers),
        val=np.array([i["val_pos"] for i in infos], dtype=np.float32),
        mean_mag=np.array([i["mean_mag"] for i in infos], dtype=np.float32),
        cut=np.array([i["cut"] for i in infos], dtype=bool), scalar=np.array(scal, dtype=np.float64),
        argmax_gap=np.array(gaps, dtype=np.float64))

    # ---- 4. post-processing known answers (reference statements F:1266-1390)
    pp = ref_loader.postproc_function(ref)
    cases = []
    rng = np.random.default_rng(5)
    for n, fps, kf, cut_at in [(400, 60.0, True, (150,)), (400, 60.0, False, (150,)), (90, 29.97, True, ()),
                               (7, 30.0, True, (3,)), (240, 25.0, True, (60, 61, 200))]:
        step = max(1, int(np.ceil(fps / 30.0)))
        vals = (np.sin(np.arange(n) * 0.21) * 4 + rng.standard_normal(n) * 0.3).tolist()
        cuts = [i in cut_at for i in range(n)]
        ffl = [(vals[i], cuts[i], i * step) for i in range(n)]
        prm = {"detrend_window": 2.0, "norm_window": 3.0, "keyframe_reduction": kf}
        acts = pp(ffl, fps, fps / step, prm, lambda *_: None)
        cases.append({"fps": fps, "params": prm, "values": vals, "cuts": cuts, "frame_indices": [i * step for i in range(n)],
                      "actions": acts})
    json.dump({"meta": meta, "cases": cases}, open(os.path.join(HERE, "postproc.json"), "w"))

    # ---- 5. the reference's process_video end to end on a C1-style clip (single bracket => no prefetch race)
    spec = ClipSpec(640, 360, 300, seed=0, amplitude=0.15, period=30.0)      # config C1 as BASELINE.json states it: 300 frames
    clip = ClipGenerator(spec).stack()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "c1.avi")
        write_video(path, clip, 30.0)
        settings = {"threads": 1, "detrend_window": 2.0, "norm_window": 3.0, "batch_size": 3000, "overwrite": True,
                    "vr_mode": False, "pov_mode": False, "keyframe_reduction": True, "backend": "CPU"}
        logs = []
        err = ref.process_video(path, settings, logs.append)
        assert not err, logs
        acts = json.load(open(os.path.join(td, "c1.funscript")))["actions"]
        # the per-pair series the reference computed on the decoded 256x256 frames (for diagnosis)
        frames = ref.fetch_frames_optimized(path, list(range(300)), settings)
        infos = [ref.precompute_wrapper(p, settings) for p in zip(frames[:-1], frames[1:])]
    json.dump({"meta": meta, "spec": {"width": 640, "height": 360, "n_frames": 300, "seed": 0, "amplitude": 0.15,
                                      "period": 30.0, "fps": 30.0},
               "settings": settings, "actions": acts,
               "centers_raw": [[int(i["pos_center"][0]), int(i["pos_center"][1])] for i in infos],
               "mean_mag": [float(i["mean_mag"]) for i in infos]},
              open(os.path.join(HERE, "video_c1.json"), "w"))
    # ---- 6. several brackets through the reference's process_video with its prefetch race neutralised (SURVEY Q3,
    # F:1155-1185): the real threading.Thread is replaced by a stand-in whose start() runs the prefetch synchronously
    # and whose is_alive() answers True, so that F:1157-1161 pops exactly the frames the prefetch just put -- the
    # semantics the code intends (every bracket computed on its own frames).  Nothing else of the reference changes.
    import types

    class _SyncThread:
        def __init__(self, target=None, args=(), kwargs=None, **_):
            self._target, self._args, self._kwargs = target, args, kwargs or {}

        def start(self):
            self._target(*self._args, **self._kwargs)

        def is_alive(self):
            return True

        def join(self, timeout=None):
            return None
    real_threading = ref.threading
    ref.threading = types.SimpleNamespace(Thread=_SyncThread, Lock=real_threading.Lock, Event=real_threading.Event)
    try:
        spec = ClipSpec(640, 360, 75, seed=4, amplitude=0.2, period=24.0)
        clip = ClipGenerator(spec).stack()
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "mb.avi")
            write_video(path, clip, 30.0)
            settings = {"threads": 1, "detrend_window": 2.0, "norm_window": 3.0, "batch_size": 30, "overwrite": True,
                        "vr_mode": False, "pov_mode": False, "keyframe_reduction": True, "backend": "CPU"}
            logs = []
            err = ref.process_video(path, settings, logs.append)
            assert not err, logs
            acts = json.load(open(os.path.join(td, "mb.funscript")))["actions"]
    finally:
        ref.threading = real_threading
    json.dump({"meta": meta, "spec": {"width": 640, "height": 360, "n_frames": 75, "seed": 4, "amplitude": 0.2, "period": 24.0,
                                      "fps": 30.0},
               "settings": settings, "actions": acts, "brackets": [[0, 30], [30, 60], [60, 75]],
               "note": "reference process_video with batch_size=30 (brackets of 30, 30 and 15 frames: 29 + 29 + 14 pairs, none "
                       "across a bracket boundary), prefetch thread run synchronously (race-neutralised, SURVEY Q3)"},
              open(os.path.join(HERE, "video_multibracket.json"), "w"))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
