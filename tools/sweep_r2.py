"""Round-2 variant sweep in ONE process on the GPU box (the kernel tuning knobs are read per launch):

    python tools/sweep_r2.py [--pairs 128] [--reps 6] [--size 1920x1080] "NAME:ENV=V,ENV=V" ...

Each configuration gets a fresh context (FFB_FLOW_STREAMS is read at creation), two warm-up brackets, then `reps`
brackets of device-resident frames timed with CUDA events on the compute stream.  Prints one JSON line per
configuration: pairs/s, per-kernel device ms per bracket, and the largest scalar difference against the first
configuration (variants that keep the summation order are bit-identical: 0.0)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

KNOBS = ("FFB_ITER_CFG", "FFB_ITER_CFG_K0", "FFB_ITER_CFG_K1", "FFB_ITER_CFG_K2", "FFB_ITER_CFG_K3", "FFB_ITER_OPT", "FFB_ITER_SWMAX", "FFB_ITER_SH", "FFB_ITER_MINSEG",
         "FFB_FLOW_STREAMS", "FFB_COPY_THREADS", "FFB_RADIAL_STREAM")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=128)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", default="1920x1080")
    ap.add_argument("configs", nargs="*")
    args = ap.parse_args()
    import torch
    from funscript_flow_b200 import _native
    from funscript_flow_b200.synth import ClipGenerator, ClipSpec
    W, H = (int(v) for v in args.size.split("x"))
    P = args.pairs
    gen = ClipGenerator(ClipSpec(W, H, 18000, seed=0, amplitude=0.15, period=30.0))
    frames = gen.stack(7, 7 + P + 1)
    d_frames = torch.from_numpy(frames).cuda()
    base = None
    for spec in args.configs or ["default:"]:
        name, _, envs = spec.partition(":")
        for k in KNOBS:
            os.environ.pop(k, None)
        batch = args.batch
        for kv in filter(None, envs.split(",")):
            k, v = kv.split("=")
            if k == "BATCH":          # pseudo-knob: frames per GPU batch of this configuration
                batch = int(v)
            else:
                os.environ[k] = v
        out = {"name": name, "env": envs}
        try:
            ctx = _native.FlowContext(0)
            ctx.configure(W, H, batch, P)

            def step():
                ctx.bracket_begin(False, 7.0)
                ctx.bracket_push_ptr(d_frames.data_ptr(), P + 1, W, W * H)
                return ctx.bracket_finish()
            for _ in range(2):
                r = step()
            ctx.sync()
            ctx.profile(True)
            ctx.profile_reset()
            ctx.timer_mark(0)
            for _ in range(args.reps):
                r = step()
            ctx.timer_mark(1)
            ctx.sync()
            ms = ctx.timer_elapsed_ms(0, 1) / args.reps
            st = ctx.kernel_stats()
            lv = ctx.flow_iter_level_stats()
            ctx.profile(False)
            out["pairs_per_s"] = round(P / ms * 1e3, 1)
            out["ms_per_bracket"] = round(ms, 3)
            out["kernel_ms"] = {k: round(v["ms"] / args.reps, 3) for k, v in st.items() if v["launches"]}
            it = st["flow_iter"]
            out["iter_frac"] = round(it["alg_bytes"] / (it["ms"] / 1e3) / 1e9 / 6455.6, 4) if it["ms"] > 0 else None
            if lv and all(v["ms"] > 0 for v in lv.values()):
                out["by_level_frac"] = {f"k{k}": round(v["alg_bytes"] / (v["ms"] / 1e3) / 1e9 / 6455.6, 3) for k, v in sorted(lv.items())}
            if base is None:
                base = r["scalar"].copy()
            out["max_scalar_diff_vs_first"] = float(np.max(np.abs(r["scalar"] - base)))
            out["cx_equal"] = True
            ctx.close()
        except Exception as exc:          # a variant that fails must not end the sweep
            out["error"] = repr(exc)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
