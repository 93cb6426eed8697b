"""Row N2 / N4 measurement (GPU box): decoded BGR frames -> 256x256 gray on the device (default), or
with --native -> gray at the decoded 1920x1080 size and the flow at that size (row N4).
Prints frames/s end to end from pinned host colour frames (H2D inside) and kernel-only, next to cv2 on
one host core (the reference's per-frame resize + cvtColor, F:173-189 / F:1079-1082)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cv2
from funscript_flow_b200 import _native

W, H, N = 1920, 1080, 64
rng = np.random.default_rng(0)
frames = rng.integers(0, 256, (N, H, W, 3), dtype=np.uint8)
ctx = _native.FlowContext(0)
NATIVE = "--native" in sys.argv
if NATIVE:
    ctx.configure(W, H, 64, 4096)
    ctx.preprocess_configure_window(W, H, (W, H), (0, 0, W, H))
else:
    ctx.configure(256, 256, 64, 4096)
    ctx.preprocess_configure(W, H, False)
pin = _native.PinnedBuffer(frames.shape)
pin.array[...] = frames
def run(k):
    ctx.bracket_begin(False, 7.0)
    for _ in range(k):
        ctx.bracket_push_bgr(pin.array)
    return ctx.bracket_finish()
run(2)
ctx.profile(True); ctx.profile_reset()
t0 = time.perf_counter(); r = run(8); t = time.perf_counter() - t0
st = ctx.kernel_stats()
pre = st["preprocess"]
cv2.setNumThreads(1)
t1 = time.perf_counter()
for f in frames[:32]:
    rgb = cv2.cvtColor(f, cv2.COLOR_BGR2RGB)
    cv2.cvtColor(rgb if NATIVE else cv2.resize(rgb, (256, 256)), cv2.COLOR_RGB2GRAY)
tc = (time.perf_counter() - t1) / 32
print(json.dumps({"row": ("N4 native-resolution" if NATIVE else "N2") + " preprocess+flow from 1080p BGR frames", "frames": 8 * N, "e2e_frames_per_s": 8 * N / t,
                  "h2d_gb_s": 8 * N * W * H * 3 / t / 1e9,
                  "k_preprocess_us_per_frame": 1000 * pre["ms"] / (8 * N), "k_preprocess_launches": pre["launches"],
                  "cv2_one_core_us_per_frame": tc * 1e6, "pairs": int(r["n_pairs"]),
                  "kernel_ms": {k: round(v["ms"], 3) for k, v in st.items()}}))
