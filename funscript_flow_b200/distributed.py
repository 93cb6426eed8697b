"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed only for rendezvous, the
barrier and a gather of the 1-D per-pair results.  The hot path has no exchange step: brackets (and
whole videos) are independent units, so there is no data-path collective -- each
rank computes its brackets and only O(pairs) scalars are gathered."""
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


def process_frames_sharded(frames: Sequence[np.ndarray], fps: float, params: Dict, frame_indices: Optional[Sequence[int]] = None,
                           ctx=None):
    """runner.process_frames with the brackets of one video sharded over the ranks.  Every rank
    returns the same (actions, series); results are identical to the single-GPU run because
    brackets are independent and per-pair reductions do not depend on the batch composition."""
    rank, ws = world()
    n = len(frames)
    idx = list(range(n)) if frame_indices is None else list(frame_indices)
    ranges = bracket_ranges(n, int(params.get("batch_size", 3000.0)))
    mine = {}
    for i in my_brackets(ranges, rank, ws):
        a, b = ranges[i]
        r = api.process_bracket(frames[a:b], params, ctx=ctx, batch_frames=int(params.get("gpu_batch_frames", api.DEFAULT_BATCH_FRAMES)))
        mine[i] = (r["scalar"], r["cut"], idx[a:b - 1])
    merged = {}
    for part in gather_objects(mine):
        merged.update(part)
    values, cuts, stamps = [], [], []
    for i in range(len(ranges)):
        s, c, t = merged[i]
        values.extend(np.asarray(s).tolist())
        cuts.extend(np.asarray(c).tolist())
        stamps.extend(t)
    actions = postproc.scalars_to_actions(values, cuts, stamps, fps, params) if values else []
    return actions, dict(values=np.asarray(values), cuts=np.asarray(cuts, bool), frame_indices=np.asarray(stamps))
