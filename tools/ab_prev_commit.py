"""A/B on the GPU box: a previous commit's kernels (scratch/libffb_prev.so, see tools/ab_prev_commit.sh) vs the working tree, in one process:
per-kernel CUDA-event times on one flow stream, bit-equality of a small bracket, wall time per 64-pair 1080p bracket on two streams."""
import json, os, sys, time
root = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, root)
import numpy as np
from funscript_flow_b200 import _native, api
from funscript_flow_b200.synth import make_clip
prev = os.path.join(root, "scratch/libffb_prev.so")
out = {}
clip = make_clip(1920, 1080, 65, seed=2)
small = make_clip(328, 200, 9, seed=3)
os.environ["FFB_FLOW_STREAMS"] = "1"
new1, old1 = _native.FlowContext(0), _native.FlowContext(0, lib_path=prev)
ra = api.process_bracket(small, {}, ctx=new1, batch_frames=4, return_flows=True)
rb = api.process_bracket(small, {}, ctx=old1, batch_frames=4, return_flows=True)
out["small_flows_equal"] = bool(np.array_equal(ra["flows"], rb["flows"]))
for name, ctx in (("old", old1), ("new", new1), ("old2", old1), ("new2", new1)):
    api.process_bracket(clip, {}, ctx=ctx, batch_frames=64)
    ctx.profile(True); ctx.profile_reset()
    for _ in range(3):
        r = api.process_bracket(clip, {}, ctx=ctx, batch_frames=64)
    st = ctx.kernel_stats()
    ctx.profile(False)
    out[name] = {k: round(v["ms"] / 3, 4) for k, v in st.items()}
    out[name + "_scalar_sum"] = repr(float(np.sum(r["scalar"])))
del os.environ["FFB_FLOW_STREAMS"]
new2, old2 = _native.FlowContext(0), _native.FlowContext(0, lib_path=prev)
for name, ctx in (("old", old2), ("new", new2), ("old2", old2), ("new2", new2)):
    api.process_bracket(clip, {}, ctx=ctx, batch_frames=64)
    t0 = time.perf_counter()
    for _ in range(8):
        api.process_bracket(clip, {}, ctx=ctx, batch_frames=64)
    out[name + "_ms_per_bracket_2streams"] = round((time.perf_counter() - t0) / 8 * 1e3, 3)
print(json.dumps(out, indent=1))
