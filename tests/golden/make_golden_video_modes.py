"""Regenerates tests/golden/video_modes.json: the reference's process_video() (AST-loaded from
/root/reference, build container only) on the C1-style clip with its other option combinations --
VR mode (512x512 resize + bottom-left quadrant, F:1074-1079), POV mode (F:880-882, F:770-775) and a 60 fps
container (step = 2 sub-sampling, F:1127).  Single bracket each, so the reference's prefetch race (SURVEY
5.3) cannot occur.  Run:  python tests/golden/make_golden_video_modes.py

Clip choice: the centre of motion is an argmax, so a clip is only a fair funscript-level fixture when that
argmax is well conditioned.  The generator parameters below were picked so that on every pair the top-1 /
top-2 gap of |divergence| is >= 1e-3 and cv2's flow and the NumPy restatement agree on its location (the
script asserts both and records the smallest gap).  A clip that violates this is kept as a documented
limitation in profiles/r1_parity_sensitivity.txt (seed 3, period 24: the reference's maximum sits on a
border pixel whose cv2 flow differs from every other implementation's by 0.05 px).
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402
import numpy as np  # noqa: E402

from funscript_flow_b200.synth import ClipGenerator, ClipSpec  # noqa: E402
from oracle import ref_loader  # noqa: E402

CASES = [
    {"name": "vr", "fps": 30.0, "n_frames": 90, "settings": {"vr_mode": True, "pov_mode": False, "keyframe_reduction": True}},
    {"name": "pov_60fps", "fps": 60.0, "n_frames": 150, "settings": {"vr_mode": False, "pov_mode": True, "keyframe_reduction": False}},
    {"name": "vr_pov", "fps": 30.0, "n_frames": 60, "settings": {"vr_mode": True, "pov_mode": True, "keyframe_reduction": True}},
    # 4 frames = 3 pairs: shorter than the 5-tap smoother, the reference keeps going on np.convolve's 5 samples,
    # drops the indices without a time stamp and returns error_occurred = True (F:1383-1385)
    {"name": "short_4_frames", "fps": 30.0, "n_frames": 4, "settings": {"vr_mode": False, "pov_mode": True, "keyframe_reduction": False}},
]


def main():
    ref = ref_loader.load(serial_pools=True)
    out = {"meta": {"cv2": cv2.__version__, "numpy": np.__version__}, "cases": []}
    for case in CASES:
        spec = dict(width=640, height=360, n_frames=case["n_frames"], seed=9, amplitude=0.3, period=20.0)
        clip = ClipGenerator(ClipSpec(spec["width"], spec["height"], spec["n_frames"], seed=spec["seed"],
                                      amplitude=spec["amplitude"], period=spec["period"])).stack()
        settings = {"threads": 1, "detrend_window": 2.0, "norm_window": 3.0, "batch_size": 3000, "overwrite": True, "backend": "CPU"}
        settings.update(case["settings"])
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "clip.avi")
            vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), case["fps"], (spec["width"], spec["height"]), True)
            assert vw.isOpened()
            for f in clip:
                vw.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
            vw.release()
            logs = []
            err = bool(ref.process_video(path, settings, logs.append))
            assert err == (case["n_frames"] < 6), logs
            acts = json.load(open(os.path.join(td, "clip.funscript")))["actions"]
        gap = None
        if not settings["pov_mode"]:      # the argmax only matters without the POV shortcut
            import math
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import parity_checks as pc
            from oracle import farneback_np as fb, preproc_np as pp
            step = max(1, int(math.ceil(case["fps"] / 30.0)))
            gray = [pp.frame_to_gray(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR), settings["vr_mode"]) for f in clip[::step]]
            gap = 1.0
            for a, b in zip(gray[:-1], gray[1:]):
                m1 = pc.argmax_margin(cv2.calcOpticalFlowFarneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0))
                m2 = pc.argmax_margin(fb.farneback(a, b))
                assert (m1[0], m1[1]) == (m2[0], m2[1]) and min(m1[3], m2[3]) >= 1e-3, (case["name"], m1, m2)
                gap = min(gap, m1[3], m2[3])
        out["cases"].append({"name": case["name"], "fps": case["fps"], "spec": spec, "settings": settings, "actions": acts,
                             "min_argmax_gap": gap, "error_occurred": err})
        print(case["name"], len(acts), "actions", "min argmax gap", gap)
    json.dump(out, open(os.path.join(HERE, "video_modes.json"), "w"))


if __name__ == "__main__":
    main()
