"""Diagnostic (GPU box): golden bracket through the CUDA path; prints per-pair deviations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cv2
from funscript_flow_b200 import _native, api
g = np.load("tests/golden/bracket.npz")
ctx = _native.FlowContext(0)
r = api.process_bracket(g["frames"], {"cut_threshold": 2.0}, ctx=ctx, batch_frames=5, return_flows=True)
rel = (r["mean_mag"] - g["mean_mag"]) / g["mean_mag"]
print("generic" if os.environ.get("FFB_PYR_GENERIC") else "fast", "mean_mag rel dev max", np.abs(rel).max())
print(np.array2string(rel, precision=2))
for j in range(r["flow_first"], r["n_pairs"]):
    ref = cv2.calcOpticalFlowFarneback(g["frames"][j], g["frames"][j + 1], None, 0.5, 3, 15, 3, 5, 1.2, 0)
    d = np.abs(r["flows"][j - r["flow_first"]] - ref)
    print(j, "med %.2e p99 %.2e max %.2e n>1e-3 %d n>0.05 %d" % (np.median(d), np.percentile(d, 99), d.max(), (d > 1e-3).sum(), (d > 0.05).sum()))
