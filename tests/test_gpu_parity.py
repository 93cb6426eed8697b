"""`-m gpu`: the parity tests proper.  The sm_100a library is called through the C ABI and compared
with the oracle (oracle/), with cv2 (the dependency the reference calls at F:878) and with golden
vectors recorded from the reference's own functions (tests/golden/).  Nothing here reads
/root/reference."""
import json
import os

import numpy as np
import pytest

import parity_checks as pc
from funscript_flow_b200 import api, postproc, runner
from funscript_flow_b200.synth import ClipGenerator, ClipSpec, make_clip
from oracle import motion_np as mo

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("size", [(150, 101), (517, 389), (640, 360)])
def test_stages(gpu_ctx, size):
    pc.check_stages(gpu_ctx, *size)


@pytest.mark.parametrize("size", [(256, 256), (640, 360), (333, 217), (517, 389), (1920, 1080)])
def test_farneback_vs_cv2(gpu_ctx, size):
    print(pc.check_farneback_vs_cv2(gpu_ctx, *size))


@pytest.mark.parametrize("size", [(3840, 2160), (5760, 2880)])
def test_farneback_vs_cv2_large(gpu_ctx, size):
    """Configs C3 / C4 frame sizes (largest per-frame working set)."""
    print(pc.check_farneback_vs_cv2(gpu_ctx, *size, period=20.0, amplitude=0.1))


def test_reduction_known_answers(gpu_ctx, golden_dir):
    pc.check_reductions_kat(gpu_ctx, golden_dir)


def test_reductions_random_fields(gpu_ctx):
    pc.check_reductions_random(gpu_ctx, n_cases=40, max_side=300)


def test_golden_pairs(gpu_ctx, golden_dir):
    pc.check_golden_pairs(gpu_ctx, golden_dir)


def test_golden_bracket(gpu_ctx, golden_dir):
    pc.check_golden_bracket(gpu_ctx, golden_dir)
    pc.check_golden_bracket(gpu_ctx, golden_dir, batch_frames=64)


def test_edge_brackets(gpu_ctx):
    pc.check_edge_brackets(gpu_ctx)


def test_batch_independence_1080p(gpu_ctx):
    """Full-size property (BASELINE config C2 frame size): results are bit-identical however the
    frames are batched / pushed, which is also what makes bracket sharding across GPUs exact."""
    pc.check_batch_independence(gpu_ctx, 1920, 1080, n_frames=20)


def test_reduction_properties_at_full_size(gpu_ctx):
    """Size-independent properties at the largest BASELINE frame (5760x2880), where a CPU comparison of
    every pixel is still cheap for the reductions: linearity of the radial mean and of the mean magnitude,
    sign / offset invariances of the divergence argmax, and agreement with the oracle."""
    h, w = 2880, 5760
    rng = np.random.default_rng(21)
    base = np.zeros((h, w, 2), np.float32)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float32)
    base[..., 0] = 0.002 * (xs - 3000) + 0.3 * np.sin(ys / 97.0)
    base[..., 1] = 0.002 * (ys - 1300) + 0.3 * np.cos(xs / 131.0)
    base += rng.standard_normal(base.shape).astype(np.float32) * 0.01
    base[1717, 4001, 0] += 3.0                      # an unmistakable divergence peak (rows 1716 / 1718 see it)
    c = (3000.25, 1300.5)
    r1 = gpu_ctx.radial_motion(base, c, False)
    r2 = gpu_ctx.radial_motion(2.0 * base, c, False)
    assert abs(r2 - 2.0 * r1) <= 1e-9 * abs(r1)                         # exact scaling by 2 in every fp32 term
    ref = mo.radial_motion_weighted(base, c, False)
    assert abs(r1 - ref) <= 1e-6 * abs(ref)
    other = rng.standard_normal(base.shape).astype(np.float32) * 0.05
    rs = gpu_ctx.radial_motion(base + other, c, False)
    assert abs(rs - (r1 + gpu_ctx.radial_motion(other, c, False))) <= 1e-5 * abs(r1)
    assert gpu_ctx.radial_motion(base, c, True) == 0.0
    x, y, v = gpu_ctx.max_divergence(base)
    assert (x, y, v) == mo.max_divergence(base)                        # bit-exact on equal input
    x2, y2, v2 = gpu_ctx.max_divergence(-base)
    assert (x2, y2) == (x, y) and v2 == -v
    m1, m2 = gpu_ctx.mean_magnitude(base), gpu_ctx.mean_magnitude(2.0 * base)
    assert abs(m2 - 2.0 * m1) <= 2e-7 * m2 and abs(m1 - mo.mean_magnitude(base)) <= 2e-6 * m1


def test_streaming_paths_agree(gpu_ctx):
    """Stream / event plumbing under real asynchrony: a 150-frame bracket pushed (a) from pageable memory
    in 3-frame batches through the pinned double buffer, (b) from pinned memory in ragged pieces, (c) from
    device memory in one piece must give bit-identical per-pair results (staging buffers are reused ~50x)."""
    import torch
    from funscript_flow_b200 import _native
    clip = make_clip(320, 240, 150, seed=13, period=17.0, amplitude=0.3)
    ref = api.process_bracket(clip, {}, ctx=gpu_ctx, batch_frames=64)
    a = api.process_bracket(clip, {}, ctx=gpu_ctx, batch_frames=3)
    pin = _native.PinnedBuffer(clip.shape)
    pin.array[...] = clip
    gpu_ctx.configure(320, 240, 7, 149)
    gpu_ctx.bracket_begin(False, 7.0)
    for lo, hi in ((0, 1), (1, 30), (30, 31), (31, 100), (100, 150)):
        gpu_ctx.bracket_push(pin.array[lo:hi])
    b = gpu_ctx.bracket_finish()
    dev = torch.from_numpy(clip).cuda()
    gpu_ctx.configure(320, 240, 32, 149)
    gpu_ctx.bracket_begin(False, 7.0)
    gpu_ctx.bracket_push_ptr(dev.data_ptr(), 150, 320, 320 * 240)
    c = gpu_ctx.bracket_finish()
    for other in (a, b, c):
        assert other["n_pairs"] == 149
        for k in ("scalar", "cut", "cx", "cy", "val", "mean_mag", "centers"):
            assert np.array_equal(ref[k], other[k]), k


def test_bracket_vs_oracle_640x360(gpu_ctx):
    """Config C1 geometry: per-pair centres (margin-guarded, the count printed), cut flags, A5 and scalars vs the oracle."""
    clip = ClipGenerator(ClipSpec(640, 360, 24, seed=0, amplitude=0.15, period=30.0)).stack(3, 27)
    pc.check_bracket_vs_oracle(gpu_ctx, clip, batch_frames=8, what="C1 640x360")


def test_bracket_vs_oracle_1080p(gpu_ctx):
    """Config C2 geometry (the benchmarked size): centres (margin-guarded), cut flags, A5, scalars."""
    clip = ClipGenerator(ClipSpec(1920, 1080, 18000, seed=0, amplitude=0.15, period=30.0)).stack(11, 19)   # the fast half of the stroke
    pc.check_bracket_vs_oracle(gpu_ctx, clip, batch_frames=4, min_clear=4, what="C2 1080p")


def test_c3_bracket_4k_pan_and_hard_cuts(gpu_ctx):
    """Config C3 as BASELINE.json states it: 3840x2160, camera pan of 3 px per sampled frame, two hard scene cuts,
    the reference's default threshold 7 (F:876), 14 pairs: identical scene-cut indices against cv2 (every pair at
    least 0.05 px away from the threshold), cut pairs contribute exactly 0.0, centres / A5 / scalars as
    in the other bracket tests."""
    spec = ClipSpec(3840, 2160, 40, seed=3, amplitude=0.15, period=20.0, pan=(3.0, 0.0), cuts=(5, 10))
    clip = ClipGenerator(spec).stack(0, 15)
    r, vals, cuts = pc.check_bracket_vs_oracle(gpu_ctx, clip, {"cut_threshold": 7}, batch_frames=8, min_clear=4, what="C3 4K")
    ref_mm = np.array([mo.mean_magnitude(cv2.calcOpticalFlowFarneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0))
                       for a, b in zip(clip[:-1], clip[1:])])
    assert (np.abs(ref_mm - 7.0) > 0.05).all(), f"clip no longer keeps every pair clear of the threshold: {ref_mm}"
    assert np.flatnonzero(ref_mm > 7.0).tolist() == [4, 9], f"clip no longer has its two hard cuts: {ref_mm}"
    assert np.flatnonzero(r["cut"]).tolist() == [4, 9]
    assert np.allclose(r["mean_mag"], ref_mm, rtol=pc.MEAN_MAG_RTOL, atol=pc.MEAN_MAG_ATOL)
    assert (r["scalar"][[4, 9]] == 0.0).all() and (r["scalar"][[0, 1, 2, 3, 5, 6, 7, 8]] != 0.0).all()


def test_c4_bracket_vr_side_by_side(gpu_ctx):
    """Config C4: 5760x2880 side-by-side frames (the largest per-frame working set: 332 MB of expansion per frame),
    5 pairs through the bracket path: cut flags, centres, A5, scalars within 1e-3 of the oracle on cv2's flow."""
    spec = ClipSpec(5760, 2880, 12, seed=6, amplitude=0.15, period=16.0, stereo=True)
    clip = ClipGenerator(spec).stack(2, 8)
    pc.check_bracket_vs_oracle(gpu_ctx, clip, batch_frames=4, min_clear=2, what="C4 5760x2880")


def test_scene_cuts_and_pan(gpu_ctx):
    """Config C3 style (pan + hard cuts) at a CPU-checkable size: identical scene-cut indices."""
    spec = ClipSpec(960, 540, 40, seed=3, amplitude=0.15, period=20.0, pan=(1.5, 0.0), cuts=(13, 29))
    clip = ClipGenerator(spec).stack()
    thr = 3.0
    ref_mm = np.array([mo.mean_magnitude(cv2.calcOpticalFlowFarneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0))
                       for a, b in zip(clip[:-1], clip[1:])])
    r = api.process_bracket(clip, {"cut_threshold": thr}, ctx=gpu_ctx, batch_frames=16)
    assert np.allclose(r["mean_mag"], ref_mm, rtol=pc.MEAN_MAG_RTOL, atol=pc.MEAN_MAG_ATOL)
    clear = np.abs(ref_mm - thr) > 0.05          # margin guard
    assert clear[[12, 28]].all() and (ref_mm[[12, 28]] > thr).all(), "clip no longer has clear cuts"
    assert np.array_equal(r["cut"][clear], (ref_mm > thr)[clear])
    assert np.flatnonzero(r["cut"]).tolist() == np.flatnonzero(ref_mm > thr).tolist()
    assert (r["scalar"][r["cut"]] == 0.0).all()


def test_dropin_functions(gpu_ctx):
    """Module-level functions with the reference's signatures (F:748, F:761, F:843, F:982, F:1019)."""
    api.set_context(gpu_ctx)
    clip = make_clip(320, 240, 4, seed=12, period=9.0, amplitude=0.3)
    ref = mo.precompute_flow_info(clip[1], clip[2], {})
    for info in (api.precompute_flow_info(clip[1], clip[2], {"backend": "CUDA"}),
                 api.precompute_flow_info_gpu(clip[1], clip[2], 7),
                 api.precompute_wrapper((clip[1], clip[2]), {"threads": 8})):
        pc.assert_flow_close(info["flow"], ref["flow"], "drop-in")
        assert info["flow"].dtype == np.float32 and info["flow"].shape == (240, 320, 2)
        assert isinstance(info["cut"], bool) and info["cut"] == ref["cut"]
        assert isinstance(info["mean_mag"], np.float32) and isinstance(info["val_pos"], np.float32)
    x, y, v = api.max_divergence(ref["flow"])
    assert (int(x), int(y)) == tuple(ref["pos_center"]) and v == ref["val_pos"]      # bit-exact on equal input
    for c in ([100.5, 80.25], [160.0, 120.0]):
        for pov in (False, True):
            a = api.radial_motion_weighted(ref["flow"], c, False, pov)
            b = mo.radial_motion_weighted(ref["flow"], c, False, pov)
            assert abs(a - b) <= 1e-6 * abs(b) + pc.scalar_tol(ref["flow"], c)
    assert api.radial_motion_weighted(ref["flow"], [1, 1], True) == 0.0
    assert api.get_available_backends()["CUDA"] is True


def test_preprocess_bit_exact(gpu_ctx):
    """Row N2: device resize + gray conversion, bit-exact with cv2 (via the oracle), incl. 4K and VR sources."""
    pc.check_preprocess(gpu_ctx, sizes=((360, 640), (300, 200), (256, 256), (1080, 1920), (2160, 3840), (2880, 5760)))
    pc.check_bgr_push_equals_gray_push(gpu_ctx)
    pc.check_bgr_push_equals_gray_push(gpu_ctx, 1920, 1080, 40)


@pytest.mark.gpu
def test_preprocess_window_and_native_resolution(gpu_ctx):
    """Row N4: windowed pre-processing bit-exact at real sizes (1080p native, 4K side-by-side VR eyes);
    a native-resolution bracket fed as colour frames equals the host-converted gray bracket."""
    pc.check_preprocess_window(gpu_ctx)
    pc.check_preprocess_window(gpu_ctx, plans=((1080, 1920, (1920, 1080), (0, 0, 1920, 1080)),
                                               (1920, 3840, (3840, 1920), (0, 960, 1920, 960)),
                                               (1920, 3840, (3840, 1920), (1920, 960, 1920, 960)),
                                               (2160, 3840, (1280, 720), (100, 50, 1000, 600))))
    pc.check_native_resolution_bracket(gpu_ctx, 1920, 1080, 6)
    pc.check_native_resolution_bracket(gpu_ctx, 1280, 640, 5, vr=True, eye="right")



def test_process_video_matches_reference_funscript(gpu_ctx, golden_dir, tmp_path):
    """End to end on the C1-style clip: the .funscript written by our process_video() has the same
    keyframe timestamps as the one the reference's process_video() wrote (recorded in video_c1.json)."""
    api.set_context(gpu_ctx)
    g = json.load(open(os.path.join(golden_dir, "video_c1.json")))
    s = g["spec"]
    clip = ClipGenerator(ClipSpec(s["width"], s["height"], s["n_frames"], seed=s["seed"], amplitude=s["amplitude"],
                                  period=s["period"])).stack()
    path = str(tmp_path / "c1.avi")
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), s["fps"], (s["width"], s["height"]), True)
    if not vw.isOpened():
        pytest.skip("FFV1 writer unavailable on this box")
    for f in clip:
        vw.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
    vw.release()
    logs = []
    assert runner.process_video(path, g["settings"], logs.append) is False, logs
    acts = json.load(open(str(tmp_path / "c1.funscript")))["actions"]
    assert [a["at"] for a in acts] == [a["at"] for a in g["actions"]], (acts, g["actions"])
    assert max(abs(a["pos"] - b["pos"]) for a, b in zip(acts, g["actions"])) <= 1
    # a second call skips because the output exists (F:1105-1109)
    logs.clear()
    assert runner.process_video(path, dict(g["settings"], overwrite=False), logs.append) is False
    assert any("Skipping" in l for l in logs)


@pytest.mark.gpu
def test_process_video_modes_match_reference_funscripts(gpu_ctx, golden_dir, tmp_path):
    """VR mode, POV mode and a 60 fps container (step-2 sub-sampling) through process_video(): device
    pre-processing + hot path + host post-processing give the keyframe timestamps the reference wrote
    (tests/golden/video_modes.json), positions within one unit."""
    api.set_context(gpu_ctx)
    g = json.load(open(os.path.join(golden_dir, "video_modes.json")))
    for case in g["cases"]:
        sp = case["spec"]
        clip = ClipGenerator(ClipSpec(sp["width"], sp["height"], sp["n_frames"], seed=sp["seed"], amplitude=sp["amplitude"],
                                      period=sp["period"])).stack()
        path = str(tmp_path / (case["name"] + ".avi"))
        vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), case["fps"], (sp["width"], sp["height"]), True)
        if not vw.isOpened():
            pytest.skip("FFV1 writer unavailable on this box")
        for f in clip:
            vw.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
        vw.release()
        logs = []
        assert runner.process_video(path, case["settings"], logs.append) is case["error_occurred"], logs
        acts = json.load(open(str(tmp_path / (case["name"] + ".funscript"))))["actions"]
        assert [a["at"] for a in acts] == [a["at"] for a in case["actions"]], (case["name"], acts, case["actions"])
        assert max(abs(a["pos"] - b["pos"]) for a, b in zip(acts, case["actions"])) <= 1, case["name"]


@pytest.mark.gpu
def test_both_eyes_two_contexts_and_abort(gpu_ctx, tmp_path):
    """Row N4 on the device: vr_eye="both" drives two contexts on one GPU from one decode and equals the mean
    of the single-eye runs; an aborted bracket leaves no trace in the next one."""
    api.set_context(gpu_ctx)
    clip = make_clip(640, 320, 12, seed=41, period=9.0, amplitude=0.3)
    path = str(tmp_path / "sbs.avi")
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), 30.0, (640, 320), True)
    if not vw.isOpened():
        pytest.skip("FFV1 writer unavailable on this box")
    for f in clip:
        vw.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
    vw.release()
    for native in (False, True):
        prm = {"batch_size": 3000, "vr_mode": True, "pov_mode": False, "gpu_batch_frames": 5, "native_resolution": native}
        left = runner.process_video_series(path, dict(prm, vr_eye="left"), ctx=gpu_ctx)
        right = runner.process_video_series(path, dict(prm, vr_eye="right"), ctx=gpu_ctx)
        both = runner.process_video_series(path, dict(prm, vr_eye="both"), ctx=gpu_ctx)
        assert left[0] != right[0]
        assert both[0] == (0.5 * (np.asarray(left[0]) + np.asarray(right[0]))).tolist()
        assert both[1] == [a or b for a, b in zip(left[1], right[1])]
    gray = make_clip(320, 200, 9, seed=42)
    full = api.process_bracket(gray, {}, ctx=gpu_ctx, batch_frames=4)
    gpu_ctx.configure(320, 200, 4, 8)
    gpu_ctx.bracket_begin(False, 7.0)
    gpu_ctx.bracket_push(gray[:6])
    gpu_ctx.bracket_abort()
    again = api.process_bracket(gray, {}, ctx=gpu_ctx, batch_frames=4)
    for k in ("scalar", "cut", "cx", "cy", "val", "mean_mag"):
        assert np.array_equal(full[k], again[k]), k


def test_multi_bracket_video_matches_race_free_reference(gpu_ctx, golden_dir, tmp_path):
    """Several brackets end to end: the reference's process_video() with batch_size=30 on a 75-frame clip (brackets of
    30, 30 and 15 frames), driven with its prefetch race neutralised (SURVEY Q3, F:1155-1185; the generating script
    replaces threading.Thread by a synchronous stand-in), wrote tests/golden/video_multibracket.json; ours must write
    the same keyframe timestamps.  No pair spans a bracket boundary and the +-6 window is cut at bracket ends."""
    api.set_context(gpu_ctx)
    g = json.load(open(os.path.join(golden_dir, "video_multibracket.json")))
    s = g["spec"]
    clip = ClipGenerator(ClipSpec(s["width"], s["height"], s["n_frames"], seed=s["seed"], amplitude=s["amplitude"],
                                  period=s["period"])).stack()
    path = str(tmp_path / "mb.avi")
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), s["fps"], (s["width"], s["height"]), True)
    if not vw.isOpened():
        pytest.skip("FFV1 writer unavailable on this box")
    for f in clip:
        vw.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
    vw.release()
    logs = []
    assert runner.process_video(path, g["settings"], logs.append) is False, logs
    acts = json.load(open(str(tmp_path / "mb.funscript")))["actions"]
    assert [a["at"] for a in acts] == [a["at"] for a in g["actions"]], (acts, g["actions"])
    assert max(abs(a["pos"] - b["pos"]) for a, b in zip(acts, g["actions"])) <= 1
    res = runner.process_video_series(path, g["settings"], ctx=gpu_ctx)
    assert res[2] == [i for a, b in g["brackets"] for i in range(a, b - 1)]      # 29 + 29 + 14 pairs, none across a boundary


def test_integration_stub_from_the_docs_on_the_real_library(gpu_ctx):
    """The ctypes stub INTEGRATION.md tells a reference maintainer to paste (replacement body of
    precompute_flow_info_gpu, F:982-1017), executed as written against libffb.so on the GPU: the 8-key dict of
    F:1008-1017 with the reference's types, equal to the shipped Python mirror's."""
    import re
    from funscript_flow_b200 import _native
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(# --- in FunscriptFlow\.pyw.*?)```", text, re.S).group(1)
    code = code.replace("/path/to/funscript_flow_b200/libffb.so", _native.DEFAULT_LIB)
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    clip = make_clip(320, 240, 3, seed=77, period=9.0, amplitude=0.3)
    got = ns["precompute_flow_info_gpu"](clip[0], clip[1], 7)
    api.set_context(gpu_ctx)
    ours = api.precompute_flow_info_gpu(clip[0], clip[1], 7)
    assert set(got) == set(ours) == {"flow", "pos_center", "neg_center", "val_pos", "val_neg", "cut", "cut_center", "mean_mag"}
    assert np.array_equal(got["flow"], ours["flow"]) and got["flow"].dtype == np.float32
    assert tuple(int(v) for v in got["pos_center"]) == tuple(int(v) for v in ours["pos_center"])
    assert got["cut"] == ours["cut"] and isinstance(got["cut"], bool)
    assert np.float32(got["val_pos"]) == np.float32(ours["val_pos"]) and np.float32(got["mean_mag"]) == np.float32(ours["mean_mag"])
    ref = cv2.calcOpticalFlowFarneback(clip[0], clip[1], None, 0.5, 3, 15, 3, 5, 1.2, 0)
    pc.assert_flow_close(got["flow"], ref, "INTEGRATION.md stub")


def test_frame_range_shards_equal_the_single_bracket(gpu_ctx):
    """Frame ranges of ONE bracket on separate contexts (one frame of overlap, raw centres exchanged between the
    phases: ffb_bracket_begin_shard / phase1_finish / radial): every per-pair output equals the single-context
    bracket bit for bit, at 1080p, for 2, 3 and 5 shards.  The shards run on every visible GPU in turn (one context
    per shard), so on a multi-GPU box this also checks GPU-to-GPU reproducibility."""
    from funscript_flow_b200 import _native
    ngpu = _native.device_count()
    clip = ClipGenerator(ClipSpec(1920, 1080, 18000, seed=0, amplitude=0.15, period=30.0)).stack(7, 7 + 41)
    ref = api.process_bracket(clip, {}, ctx=gpu_ctx, batch_frames=16)
    for parts in (2, 3, 5):
        ctxs = [_native.FlowContext(i % ngpu) for i in range(parts)]
        try:
            got = api.process_bracket_on_contexts(clip, {}, ctxs, batch_frames=8)
        finally:
            for c in ctxs:
                c.close()
        for k in ("scalar", "cut", "cx", "cy", "val", "mean_mag", "centers"):
            assert np.array_equal(ref[k], got[k]), (parts, k)


def test_no_allocation_after_the_first_bracket(gpu_ctx, tmp_path):
    """a video is configured once; brackets after the first one (the shorter last bracket included),
    and per-pair drop-in calls of the same frame size in between, make no cudaMalloc / cudaHostAlloc."""
    api.set_context(gpu_ctx)
    clip = make_clip(640, 360, 50, seed=19, period=13.0, amplitude=0.3)
    path = str(tmp_path / "v.avi")
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), 30.0, (640, 360), True)
    if not vw.isOpened():
        pytest.skip("FFV1 writer unavailable on this box")
    for f in clip:
        vw.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
    vw.release()
    prm = {"batch_size": 20, "vr_mode": False, "pov_mode": False, "native_resolution": True, "gpu_batch_frames": 8}
    first = runner.process_video_series(path, prm, ctx=gpu_ctx)           # allocates for 640x360
    base = gpu_ctx.alloc_counts()
    again = runner.process_video_series(path, prm, ctx=gpu_ctx)           # brackets of 20, 20 and 10 frames: nothing new
    assert gpu_ctx.alloc_counts() == base and again[0] == first[0]
    info = api.precompute_flow_info(clip[0], clip[1], {})
    api.radial_motion_weighted(info["flow"], info["pos_center"], info["cut"])
    assert gpu_ctx.alloc_counts() == base
    third = runner.process_video_series(path, prm, ctx=gpu_ctx)
    assert gpu_ctx.alloc_counts() == base and third[0] == first[0]


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _torchrun(nproc, args, cwd, port=None, timeout=900):
    import subprocess
    import sys
    port = _free_port()          # a fixed port could collide with another job on the box
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + args
    res = subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=timeout)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return res


def test_run_headless_under_torchrun_matches_single_process(gpu_ctx, tmp_path):
    """Row N3 / config C5 in miniature on real GPUs: `torchrun -m funscript_flow_b200 FOLDER` with one process per
    GPU (two processes folded onto one GPU when only one is visible) deals the videos longest-first to the ranks
    (F:2606-2638 per rank, run.<rank>.log) and writes the same funscripts as a single process; a second run skips every
    file because the outputs exist (F:1105-1109)."""
    from funscript_flow_b200 import _native
    api.set_context(gpu_ctx)
    nproc = max(2, min(4, _native.device_count()))
    lengths = [40, 22, 31, 12]
    for k, n in enumerate(lengths):
        clip = make_clip(320, 240, n, seed=50 + k, period=9.0 + k, amplitude=0.3)
        vw = cv2.VideoWriter(str(tmp_path / f"v{k}.avi"), cv2.VideoWriter_fourcc(*"FFV1"), 30.0, (320, 240), True)
        if not vw.isOpened():
            pytest.skip("FFV1 writer unavailable on this box")
        for f in clip:
            vw.write(cv2.cvtColor(f, cv2.COLOR_GRAY2BGR))
        vw.release()
    prm = {"threads": 8, "detrend_window": 2.0, "norm_window": 3.0, "batch_size": 3000, "overwrite": True, "vr_mode": False,
           "pov_mode": False, "keyframe_reduction": True, "backend": "CUDA"}
    cwd = os.getcwd()
    os.chdir(str(tmp_path))
    try:
        assert runner.run_headless(str(tmp_path), prm, log_func=lambda m: None) == 0
    finally:
        os.chdir(cwd)
    single = {k: json.load(open(str(tmp_path / f"v{k}.funscript"))) for k in range(4)}
    for k in range(4):
        (tmp_path / f"v{k}.funscript").unlink()
    _torchrun(nproc, ["-m", "funscript_flow_b200", str(tmp_path), "--disable_keyframe_reduction", "--overwrite"], str(tmp_path), 29533)
    multi = {k: json.load(open(str(tmp_path / f"v{k}.funscript"))) for k in range(4)}
    assert multi == single
    logs = [open(str(tmp_path / f"run.{r}.log")).read() for r in range(nproc)]
    assert all("Batch processing complete." in l for l in logs)
    assert sum(l.count("Funscript saved") for l in logs) == 4 and "v0.avi" in logs[0]      # the longest video goes to rank 0
    res = _torchrun(nproc, ["-m", "funscript_flow_b200", str(tmp_path), "--disable_keyframe_reduction"], str(tmp_path), 29534)
    assert res.stdout.count("Skipping: output file exists") == 4


SHARD_WORKER = r"""
import json, os, sys
import numpy as np
from funscript_flow_b200 import _native, api, distributed
from funscript_flow_b200.synth import ClipGenerator, ClipSpec
backend = sys.argv[2]
rank, ws = distributed.init(backend)
ctx = _native.FlowContext(api.default_device())
api.set_context(ctx, api.default_device())
clip = ClipGenerator(ClipSpec(1280, 720, 18000, seed=2, amplitude=0.2, period=24.0)).stack(5, 5 + 34)
one = distributed.process_bracket_sharded(clip, {}, ctx=ctx, batch_frames=8)
prm = {"batch_size": 12, "detrend_window": 2.0, "norm_window": 3.0, "keyframe_reduction": True, "gpu_batch_frames": 8}
acts, series = distributed.process_frames_sharded(clip, 30.0, prm, ctx=ctx)
if rank == 0:
    json.dump({"one": {k: np.asarray(v).tolist() for k, v in one.items()}, "actions": acts, "values": series["values"].tolist(),
               "ws": ws, "device": api.default_device()}, open(sys.argv[1], "w"))
import torch.distributed as dist
if dist.is_initialized():
    dist.barrier(); dist.destroy_process_group()
ctx.close()
"""


def test_bracket_sharded_over_ranks_on_gpus(gpu_ctx, tmp_path):
    """SURVEY 8(e) on hardware: one 33-pair 720p bracket cut into frame ranges over the ranks of a torchrun job (NCCL
    all-gather of the raw centres and of the scalars when every rank has its own GPU, gloo when ranks share one):
    identical, bit for bit, to the single-process bracket; the same for a 3-bracket series with post-processing."""
    from funscript_flow_b200 import _native
    ngpu = _native.device_count()
    nproc = max(2, min(4, ngpu))
    backend = "nccl" if ngpu >= nproc else "gloo"
    script = str(tmp_path / "worker.py")
    open(script, "w").write(SHARD_WORKER)
    _torchrun(nproc, [script, str(tmp_path / "multi.json"), backend], str(tmp_path), 29535)
    multi = json.load(open(str(tmp_path / "multi.json")))
    clip = ClipGenerator(ClipSpec(1280, 720, 18000, seed=2, amplitude=0.2, period=24.0)).stack(5, 5 + 34)
    ref = api.process_bracket(clip, {}, ctx=gpu_ctx, batch_frames=16)
    assert multi["ws"] == nproc and multi["one"]["n_pairs"] == 33
    for k in ("scalar", "cut", "cx", "cy", "val", "mean_mag", "centers"):
        assert np.array_equal(np.asarray(multi["one"][k]), np.asarray(ref[k]).tolist()), k
    prm = {"batch_size": 12, "detrend_window": 2.0, "norm_window": 3.0, "keyframe_reduction": True, "gpu_batch_frames": 8}
    acts, series = runner.process_frames(clip, 30.0, prm, ctx=gpu_ctx, return_series=True)
    assert multi["values"] == series["values"].tolist() and multi["actions"] == acts
