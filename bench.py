#!/usr/bin/env python
"""bench.py -- frame-pairs/sec of the hot path at 1080p (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

A step = one bracket of (pairs_per_step + 1) synthetic 1080p frames (config C2 of BASELINE.json:
"synthetic 1920x1080 30 fps 10-minute clip on 1 B200", a window of it) through the whole hot path:
pyramid + polynomial expansion per frame, 3 flow iterations on each of the 4 levels per pair,
divergence argmax + magnitude mean, +-6 centre smoothing, radial reduction, D2H of the per-pair
scalars.  `value` is measured with the frames already resident in HBM (CUDA events on the library's
compute stream); `e2e` through the public bracket API from pinned host frames (H2D inside the timed
region, wall clock between device synchronisations).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "frame_pairs_per_sec_1080p"
UNIT = "pairs/s"
# algorithmic bytes per pixel of the fused flow iteration (R0 20 + R1 20 + flow in 8 + out 8)
ITER_BYTES_PER_PX = 56.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--pairs-per-step", type=int, default=256,
                    help="pairs of one bracket = one step (the reference's brackets hold 3000 frames, F:2661)")
    ap.add_argument("--batch-frames", type=int, default=64)
    ap.add_argument("--cpu-sample-pairs", type=int, default=0, help="0 = 2 x host cores (bounded to 8..64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the process to the GPU's NUMA node")
    return ap.parse_args()


def workload_frames(width, height, n, rank):
    from funscript_flow_b200.synth import ClipGenerator, ClipSpec
    spec = ClipSpec(width, height, 18000, seed=0, amplitude=0.15, period=30.0)   # config C2 generator
    gen = ClipGenerator(spec)
    start = 7 + rank * n
    return gen.stack(start, start + n)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def cpu_reference(frames, cores, sample_pairs):
    from oracle import cpu_pipeline
    sub = list(frames[:sample_pairs + 1])
    _, _, sec, sec_flow = cpu_pipeline.run_bracket(sub, {}, cores)
    return sample_pairs / sec, sec, sample_pairs / sec_flow


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (cv2 + NumPy through the
    oracle port, all host cores), rank 0 only."""
    if rank != 0:
        return
    from oracle import cpu_pipeline
    cores = cpu_pipeline.usable_cores()
    sample = args.cpu_sample_pairs or int(min(64, max(8, 2 * cores)))
    frames = workload_frames(args.width, args.height, sample + 1, 0)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference(frames, cores, min(sample, cores))
    t = []
    for _ in range(args.steps):
        _, sec, _ = cpu_reference(frames, cores, sample)
        t.append(sec)
    total = sum(t)
    value = sample * args.steps / total
    import cv2
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C2: synthetic {args.width}x{args.height} 30 fps clip, window of {args.pairs_per_step + 1} frames per step per GPU",
                       "pairs_per_step": args.pairs_per_step, "sample_pairs_per_step": sample, "levels": 4, "iterations": 3,
                       "note": "each step is a bounded sample (the first sample_pairs_per_step pairs) of the GPU arm's step"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample} pairs/step x {args.steps} steps, cv2 {cv2.__version__} Farneback + NumPy via "
                                       f"multiprocessing.Pool({cores}) as F:1190-1236"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    from funscript_flow_b200 import _native, build
    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    # host side of the e2e path: staging buffers and upload calls on the NUMA node of the GPU's PCIe root
    all_cpus = os.sched_getaffinity(0)
    numa_node = None
    if not args.no_numa_bind:
        from funscript_flow_b200 import distributed as ffdist
        numa_node = ffdist.bind_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL announces its version on stdout when the communicator is created; stdout carries the ONE JSON
        # line of the contract, so that chatter goes to stderr
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group(backend="nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    W, H, P = args.width, args.height, args.pairs_per_step
    nf = P + 1
    frames = workload_frames(W, H, nf, rank)                     # uint8 [nf, H, W]; 65 x 2 MB > L2
    ctx = _native.FlowContext(local)
    ctx.configure(W, H, args.batch_frames, P)
    d_frames = torch.from_numpy(frames).cuda()
    pinned = _native.PinnedBuffer(frames.shape)
    pinned.array[...] = frames

    def step_resident():
        ctx.bracket_begin(False, 7.0)
        ctx.bracket_push_ptr(d_frames.data_ptr(), nf, W, W * H)
        return ctx.bracket_finish()

    def step_e2e():
        ctx.bracket_begin(False, 7.0)
        ctx.bracket_push(pinned.array)
        return ctx.bracket_finish()

    def fence():
        ctx.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm ("value") --------------------------------------------------------
    r0 = None
    for _ in range(args.warmup):
        r0 = step_resident()
    fence()
    sampler = ClockSampler(local)
    sampler.start()
    ctx.profile(True)
    ctx.profile_reset()
    l0 = ctx.launch_count
    ctx.timer_mark(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = step_resident()
    ctx.timer_mark(1)
    fence()
    wall_res = time.perf_counter() - t0
    dev_ms = ctx.timer_elapsed_ms(0, 1)
    launches = ctx.launch_count - l0
    stats = ctx.kernel_stats()
    by_level = ctx.flow_iter_level_stats()
    ctx.profile(False)
    # ---- end-to-end arm ("e2e") -----------------------------------------------------------------
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    fence()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r_e = step_e2e()
    fence()
    wall_e2e = time.perf_counter() - t0
    clocks = sampler.stop()
    # the device-resident and the end-to-end arm (and the warm-up) computed the same numbers
    assert r["n_pairs"] == P and np.array_equal(r["scalar"], r_e["scalar"])
    assert r0 is None or np.array_equal(r["scalar"], r0["scalar"])

    dev_ms = max_over_ranks(dev_ms)
    wall_e2e = max_over_ranks(wall_e2e)
    total_pairs = P * args.steps * world
    value = total_pairs / (dev_ms / 1000.0)
    e2e = total_pairs / wall_e2e

    if rank == 0:
        peak, peak_kind = measured_peak()
        it = stats["flow_iter"]
        achieved = it["alg_bytes"] / (it["ms"] / 1000.0) / 1e9 if it["ms"] > 0 else 0.0
        kernel_ms = {k: round(v["ms"] / args.steps, 4) for k, v in stats.items()}
        traffic, traffic_note = None, None
        try:   # DRAM bytes of the dominant launch shape from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            if (W, H) == (1920, 1080):
                traffic = tj["traffic"]
                traffic_note = (f"{tj['launch_shape']}: dram read+write {tj['traffic'] / 1e9:.3f} GB vs algorithmic "
                                f"{tj['alg_bytes_per_launch'] / 1e9:.3f} GB per launch ({tj['source']})")
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C2: synthetic {W}x{H} 30 fps clip, window of {nf} frames per step per GPU",
                       "pairs_per_step": P, "batch_frames": args.batch_frames, "levels": 4, "iterations": 3,
                       "l2": "inputs_exceed_l2 (per-step working set >> 126 MB)", "parallelism": f"brackets x{world}",
                       "numa_node": numa_node},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(nf * W * H), "d2h_bytes_per_step": int(P * 41),
                    "ms_per_step": 1000 * wall_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_flow_iter", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
                         "alg_bytes_per_launch": it["alg_bytes"] / max(1, it["launches"]),
                         "avg_launch_ms": it["ms"] / max(1, it["launches"]), "launches": it["launches"],
                         "timing": ("2 streams: the launch chains of the two half-batches overlap, so the time is the device "
                                    "time of the flow phases (CUDA events on the compute stream around fork..join) and "
                                    "achieved = bytes of all k_flow_iter launches / that time; FFB_FLOW_STREAMS=1 gives "
                                    "per-launch times and by_level")
                         if os.environ.get("FFB_FLOW_STREAMS", "2") != "1" else "per-launch CUDA events",
                         "by_level": {f"k{k}": {"launches": v["launches"], "ms_per_launch": v["ms"] / v["launches"],
                                                "achieved": v["alg_bytes"] / (v["ms"] / 1000.0) / 1e9,
                                                "frac": v["alg_bytes"] / (v["ms"] / 1000.0) / 1e9 / peak}
                                      for k, v in sorted(by_level.items()) if v["ms"] > 0},
                         "whole_path_bytes_per_pair": 269.7 * W * H,
                         "whole_path_frac": 269.7 * W * H * value / world / 1e9 / peak},
            "kernel_ms_per_step": kernel_ms,
            "kernel_ms_note": ("CUDA-event time per kernel class; with 2 flow streams k_flow_iter is the flow-phase time and "
                               "k_divmag (launched per slice, overlapping the other slice's flow tail) includes that overlap"),
            "wall_ms_per_step_resident": 1000 * wall_res / args.steps,
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_pipeline
            import cv2
            os.sched_setaffinity(0, all_cpus)      # the CPU baseline gets every host core again
            cores = cpu_pipeline.usable_cores()
            sample = args.cpu_sample_pairs or int(min(64, max(8, 2 * cores)))
            sample = min(sample, P)
            v, sec, vflow = cpu_reference(frames, cores, sample)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {sample} pairs of the same step, cv2 {cv2.__version__} Farneback + NumPy via "
                                              f"multiprocessing.Pool({cores}) as F:1190-1236, {sec:.1f} s",
                                    "flow_phase_only": vflow}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
