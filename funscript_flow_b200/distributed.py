"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for rendezvous, the barrier and the
exchange of 1-D per-pair results.  Whole videos and brackets are independent units; inside a
bracket, frame ranges carry one frame of overlap and the only exchange is an all-gather of the raw centres
(int32 x, y per pair) between the flow phase and the radial phase, plus the gather of the per-pair scalars --
kilobytes, never image data."""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import api, postproc


def world() -> (int, int):
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        elif part:
            cpus.append(int(part))
    return cpus


def bind_to_gpu_numa_node(device: Optional[int] = None, sysfs: str = "/sys") -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned staging buffers
    (first touch) and the upload threads are local to the GPU's PCIe root.  Call before the context is
    created.  Returns the node, or None when the topology is flat / unknown (nothing is changed then)."""
    from . import _native
    dev = api.default_device() if device is None else int(device)
    try:
        bus = _native.device_pci_bus_id(dev)
        if not bus:
            return None
        node = int(open(os.path.join(sysfs, "bus/pci/devices", bus, "numa_node")).read().strip())
        if node < 0:
            return None
        cpus = set(_parse_cpulist(open(os.path.join(sysfs, "devices/system/node", f"node{node}", "cpulist")).read()))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError):
        return None


def init(backend: Optional[str] = None):
    """Initialise torch.distributed from the torchrun environment (NCCL on GPU boxes, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    if ws > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend, rank=rank, world_size=ws)
    return rank, ws


def bracket_ranges(n_frames: int, bracket: int) -> List[tuple]:
    """Bracket [a, b) frame ranges of F:1145-1153 (brackets with < 2 frames are dropped)."""
    return [(a, min(a + bracket, n_frames)) for a in range(0, n_frames, bracket) if min(a + bracket, n_frames) - a >= 2]


def my_brackets(ranges: Sequence[tuple], rank: int, ws: int) -> List[int]:
    """Round-robin bracket ownership: bracket i belongs to rank i % world."""
    return [i for i in range(len(ranges)) if i % ws == rank]


def gather_objects(obj):
    """all_gather of a small picklable object (per-pair scalars only); identity when world == 1."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def all_gather_padded(arr: np.ndarray, max_rows: int) -> List[np.ndarray]:
    """all_gather of one small numeric array per rank (rows along axis 0, at most max_rows of them): the arrays are
    padded to max_rows, exchanged with ONE tensor collective (NCCL on GPU boxes, gloo on CPU) and cut back.
    This is the only exchange step of the path: O(pairs) integers / scalars, never image data."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [arr]
    ws = dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    # payload in float64 (exact for int32 and float32 values), first element = row count
    cols = int(np.prod(arr.shape[1:])) if arr.ndim > 1 else 1
    buf = np.zeros((max_rows * cols + 1,), np.float64)
    buf[0] = arr.shape[0]
    buf[1:1 + arr.shape[0] * cols] = np.asarray(arr, np.float64).reshape(-1)
    mine = torch.from_numpy(buf).to(dev)
    out = [torch.empty_like(mine) for _ in range(ws)]
    dist.all_gather(out, mine)
    res = []
    for t in out:
        v = t.cpu().numpy()
        n = int(v[0])
        res.append(v[1:1 + n * cols].reshape((n,) + tuple(arr.shape[1:])).astype(arr.dtype))
    return res


def process_bracket_sharded(frames: Sequence[np.ndarray], params: Dict, ctx=None, batch_frames: Optional[int] = None) -> Dict:
    """One bracket split into world-size frame ranges (F:1188, F:1203-1214): rank r runs pairs [a_r, b_r)
    on frames a_r .. b_r (one frame of overlap), the raw centres (int32 x, y per pair) are all-gathered after phase 1,
    every rank evaluates the +-6 window of its pairs on the gathered centres and the per-pair scalars are all-gathered.
    Every rank returns the whole bracket's arrays; they equal api.process_bracket's bit for bit."""
    rank, ws = world()
    ctx = ctx or api.get_context()
    n_pairs = len(frames) - 1
    bounds = api.shard_bounds(n_pairs, ws)
    a, b = bounds[rank]
    bf = int(batch_frames or params.get("gpu_batch_frames", api.DEFAULT_BATCH_FRAMES))
    most = max(hi - lo for lo, hi in bounds)
    if b > a:
        p1 = api.shard_phase1(ctx, frames, a, b, params, bf)
        local = np.stack([p1["cx"], p1["cy"], p1["cut"].astype(np.int32)], axis=1).astype(np.int32)
        extra = np.stack([p1["val"], p1["mean_mag"]], axis=1).astype(np.float32)
    else:
        local = np.zeros((0, 3), np.int32)
        extra = np.zeros((0, 2), np.float32)
    raw = np.concatenate(all_gather_padded(local, most))                 # [n_pairs, 3]: cx, cy, cut
    if b > a:
        scalar, centers = api.shard_phase2(ctx, raw[:, 0], raw[:, 1], a, b)
        mine = np.concatenate([scalar[:, None], centers, extra.astype(np.float64)], axis=1)
    else:
        mine = np.zeros((0, 5), np.float64)
    allv = np.concatenate(all_gather_padded(mine, most))                 # [n_pairs, 5]: scalar, centre x, y, val, mean_mag
    return dict(n_pairs=n_pairs, cx=raw[:, 0].copy(), cy=raw[:, 1].copy(), cut=raw[:, 2].astype(bool), scalar=allv[:, 0].copy(),
                centers=allv[:, 1:3].copy(), val=allv[:, 3].astype(np.float32), mean_mag=allv[:, 4].astype(np.float32))


def process_frames_sharded(frames: Sequence[np.ndarray], fps: float, params: Dict, frame_indices: Optional[Sequence[int]] = None,
                           ctx=None, mode: str = "frames"):
    """runner.process_frames with one video spread over the ranks.  Every rank returns the same (actions, series),
    identical to the single-GPU run: per-pair results do not depend on how frames are batched or sharded.

    mode "frames"   (default) every bracket is cut into world-size frame ranges (process_bracket_sharded), so a
                    single bracket -- any video up to 3000 sampled frames -- already occupies all GPUs;
    mode "brackets" whole brackets are dealt round-robin (no exchange at all, but at most one GPU per bracket)."""
    rank, ws = world()
    n = len(frames)
    idx = list(range(n)) if frame_indices is None else list(frame_indices)
    ranges = bracket_ranges(n, int(params.get("batch_size", 3000.0)))
    merged = {}
    if mode == "frames":
        for i, (a, b) in enumerate(ranges):
            r = process_bracket_sharded(frames[a:b], params, ctx=ctx)
            merged[i] = (r["scalar"], r["cut"], idx[a:b - 1])
    elif mode == "brackets":
        mine = {}
        for i in my_brackets(ranges, rank, ws):
            a, b = ranges[i]
            r = api.process_bracket(frames[a:b], params, ctx=ctx, batch_frames=int(params.get("gpu_batch_frames", api.DEFAULT_BATCH_FRAMES)))
            mine[i] = (r["scalar"], r["cut"], idx[a:b - 1])
        for part in gather_objects(mine):
            merged.update(part)
    else:
        raise ValueError("mode must be 'frames' or 'brackets'")
    values, cuts, stamps = [], [], []
    for i in range(len(ranges)):
        s, c, t = merged[i]
        values.extend(np.asarray(s).tolist())
        cuts.extend(np.asarray(c).tolist())
        stamps.extend(t)
    actions = postproc.scalars_to_actions(values, cuts, stamps, fps, params) if values else []
    return actions, dict(values=np.asarray(values), cuts=np.asarray(cuts, bool), frame_indices=np.asarray(stamps))
