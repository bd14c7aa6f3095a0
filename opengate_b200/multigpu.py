"""Multi-GPU plumbing: one process per GPU (torchrun), histories sharded by global history id,
one NCCL sum-reduce of every accumulator grid at the end of a run.

The reference has no collective at all: its parallel runs are merged offline by summing images
and `number_of_samples` (opengate/actors/dataitems.py:1238-1247; thread split rule
opengate/sources/base.py:190-197).  Here the same merge is one `torch.distributed.reduce`
per grid over NVLink, on the device buffers the C-ABI exposes.
"""

from __future__ import annotations

import os

from . import abi


def world_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def split_primaries(n, world, rank):
    """n//T + (t < n%T), the reference's per-thread split (sources/base.py:190-197).
    Returns (first_history_offset, count) of this rank inside the run's id range."""
    base, rem = divmod(int(n), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


class _DeviceArray:
    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def grid_tensor(engine, actor, which):
    """torch view (no copy) of one fp64 accumulator grid living in the engine's device memory."""
    import torch
    ptr, n = engine.dose_grid_device_ptr(actor, which)
    return torch.as_tensor(_DeviceArray(ptr, n), device=torch.device("cuda", engine.device))


def active_grids(score_flags):
    f = score_flags | abi.SCORE_EDEP
    return [w for w, b in ((abi.GRID_EDEP, abi.SCORE_EDEP), (abi.GRID_EDEP_SQ, abi.SCORE_EDEP_SQ),
                           (abi.GRID_DOSE, abi.SCORE_DOSE), (abi.GRID_DOSE_SQ, abi.SCORE_DOSE_SQ),
                           (abi.GRID_COUNTS, abi.SCORE_COUNTS)) if f & b]


def reduce_dose_grids(engine, actor_ids, setup_or_flags, dst=0, group=None):
    """Sum every active grid of the given actors onto rank `dst` (in place, NCCL or gloo)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    descs = None
    if hasattr(setup_or_flags, "dose"):   # dose actors first, then the TLE dose actors (engine id order)
        descs = list(setup_or_flags.dose) + [t[0] for t in getattr(setup_or_flags, "tle", [])]
    for i, a in enumerate(actor_ids):
        flags = descs[i].score_flags if descs is not None else setup_or_flags[i]
        for w in active_grids(flags):
            dist.reduce(grid_tensor(engine, a, w), dst=dst, op=dist.ReduceOp.SUM, group=group)


def reduce_counters(stats: dict, device, group=None):
    """Sum the integer run statistics over ranks (histories, steps, ...); every rank gets the total."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return dict(stats)
    keys = [k for k in stats if k not in ("kernel_ms", "edep_total_MeV")]
    t = torch.tensor([int(stats[k]) for k in keys], dtype=torch.int64, device=device)
    dist.all_reduce(t, group=group)
    e = torch.tensor([stats["edep_total_MeV"], stats["kernel_ms"]], dtype=torch.float64, device=device)
    emax = e.clone()
    dist.all_reduce(e, group=group)
    dist.all_reduce(emax, op=dist.ReduceOp.MAX, group=group)
    out = {k: int(v) for k, v in zip(keys, t.tolist())}
    out["edep_total_MeV"] = float(e[0])
    out["kernel_ms"] = float(emax[1])
    return out
