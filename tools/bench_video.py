"""End-to-end wall clock of process_video() on the C1-style clip (640x360, 300 frames, 30 fps) written as a lossless
FFV1 .avi and as MJPG -- the call a user of the reference makes (F:1094); decode on the host, everything else on the GPU.
The reference's own process_video needs 11.4 s for this clip on 8 cores (SURVEY appendix A, e2e.py probe)."""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
from funscript_flow_b200 import api, runner
from funscript_flow_b200.synth import ClipGenerator, ClipSpec

clip = ClipGenerator(ClipSpec(640, 360, 300, seed=0, amplitude=0.15, period=30.0)).stack()
prm = {"threads": 8, "detrend_window": 2.0, "norm_window": 3.0, "batch_size": 3000, "overwrite": True, "vr_mode": False,
       "pov_mode": False, "keyframe_reduction": False, "backend": "CUDA"}
out = {}
with tempfile.TemporaryDirectory() as td:
    for codec in ("FFV1", "MJPG"):
        path = os.path.join(td, f"c1_{codec}.avi")
        vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*codec), 30.0, (640, 360), True)
        if not vw.isOpened():
            continue
        for f in clip:
            vw.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
        vw.release()
        logs = []
        runner.process_video(path, prm, logs.append)                    # warm-up (context, buffers)
        t0 = time.perf_counter()
        err = runner.process_video(path, prm, logs.append)
        t_all = time.perf_counter() - t0
        t0 = time.perf_counter()
        n = sum(1 for _ in runner.iter_sampled_bgr(path, list(range(300))))
        t_dec = time.perf_counter() - t0
        ctx = api.get_context()
        ctx.profile(True); ctx.profile_reset()
        runner.process_video_series(path, prm)
        st = ctx.kernel_stats(); ctx.profile(False)
        out[codec] = {"process_video_s": round(t_all, 4), "decode_only_s": round(t_dec, 4), "frames": n, "error": bool(err),
                      "gpu_kernel_ms": round(sum(v["ms"] for v in st.values()), 3), "file_MB": round(os.path.getsize(path) / 1e6, 1)}
print(json.dumps({"clip": "C1 640x360 x300 @30fps", "reference_process_video_s_8cores": 11.4, **out}))
