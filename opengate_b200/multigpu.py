"""Multi-GPU plumbing: one process per GPU (torchrun), histories sharded by global history id,
one NCCL sum-reduce of the accumulator records at the end of a run.

The reference has no collective at all: its parallel runs are merged offline by summing images
and `number_of_samples` (opengate/actors/dataitems.py:1238-1247; thread split rule
opengate/sources/base.py:190-197).  Here the same merge is one `torch.distributed.reduce`
per dose actor over NVLink, on the device buffers the C-ABI exposes (`gate_b200_dose_buffers`).
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


def buffer_tensors(engine, actor):
    """torch views (no copy) of the accumulator buffers of one dose actor in the engine's device memory:
    the voxel records (the scored ones of edep, edep2, dose, dose2: stride * nvox doubles, see
    record_layout) and, when scored, the counts image."""
    import torch
    return [torch.as_tensor(_DeviceArray(ptr, n), device=torch.device("cuda", engine.device))
            for ptr, n in engine.dose_buffers(actor)]


def device_synchronize(device):
    """Host-side wait for everything queued on `device` (torch's streams, NCCL's and the engine's)."""
    import torch
    torch.cuda.synchronize(torch.device("cuda", int(device)))


def record_layout(score_flags):
    """(stride, {grid: offset}) of the voxel records of a dose actor: the scored ones of edep, edep2,
    dose, dose2 in that order, 1, 2 or 4 doubles per voxel (same rule as gate_b200_add_dose_actor)."""
    act = [w for w in active_grids(score_flags) if w != abi.GRID_COUNTS]
    stride = 1 if len(act) <= 1 else (2 if len(act) == 2 else 4)
    return stride, {w: i for i, w in enumerate(act)}


def active_grids(score_flags):
    f = score_flags | abi.SCORE_EDEP
    return [w for w, b in ((abi.GRID_EDEP, abi.SCORE_EDEP), (abi.GRID_EDEP_SQ, abi.SCORE_EDEP_SQ),
                           (abi.GRID_DOSE, abi.SCORE_DOSE), (abi.GRID_DOSE_SQ, abi.SCORE_DOSE_SQ),
                           (abi.GRID_COUNTS, abi.SCORE_COUNTS)) if f & b]


def reduce_dose_grids(engine, actor_ids, setup_or_flags=None, dst=0, group=None):
    """Sum the accumulator buffers of the given actors onto rank `dst` (in place, NCCL): one reduce of
    the voxel-record array per actor (+ one of the counts image when it is scored)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for a in actor_ids:
        for t in buffer_tensors(engine, a):
            dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM, group=group)


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


class SharedHostGrids:
    """Host arrays shared by the ranks of ONE node (POSIX shared memory), page-locked by every rank: after the
    all-reduce each rank holds the summed grids and copies a different slice of every grid into the shared array,
    so the device -> host read of the result runs over the host links of all GPUs at once instead of rank 0's alone
    (config 3 on 8 GPUs: 1.9 GB, 36 ms through one link).  Rank 0 owns the segment; `arrays[k]` are (z,y,x) float64
    views, complete on every rank after `gather()`."""

    def __init__(self, shapes, group=None, pin=True):
        import numpy as np
        import torch.distributed as dist
        from multiprocessing import shared_memory
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        sizes = [int(np.prod(s)) for s in shapes]
        total = 8 * sum(sizes)
        name = [None]
        if self.rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=max(total, 8))
            name[0] = self.shm.name
        dist.broadcast_object_list(name, src=0, group=group)
        if self.rank != 0:
            self.shm = shared_memory.SharedMemory(name=name[0])
        self.whole = np.ndarray((sum(sizes),), dtype=np.float64, buffer=self.shm.buf)
        self.pinned = False
        if pin:
            from . import core
            self.pinned = bool(core.host_register(self.whole))
        self.arrays, off = [], 0
        for s, n in zip(shapes, sizes):
            self.arrays.append(self.whole[off:off + n].reshape(s))
            off += n
        dist.barrier(group=group)

    def slice_of(self, k):
        """(first voxel, count) of grid k this rank copies: the split rule of the histories, over the voxels"""
        return split_primaries(self.arrays[k].size, self.world, self.rank)

    def gather(self, engine, actor, which_list, scales=None):
        """every rank copies its slice of each grid of `actor` from `engine` (which holds the all-reduced sums)
        into the shared arrays; barrier; the arrays are complete on every rank"""
        import torch.distributed as dist
        for k, w in enumerate(which_list):
            first, count = self.slice_of(k)
            engine.dose_grid_range(actor, w, first, count, self.arrays[k], scale=(scales[k] if scales else 1.0))
        dist.barrier(group=self.group)
        return self.arrays

    def close(self):
        import torch.distributed as dist
        if self.pinned:
            from . import core
            core.host_unregister(self.whole)
        self.arrays, self.whole = [], None
        dist.barrier(group=self.group)
        try:
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:
            pass


def all_reduce_dose_grids(engine, actor_ids, group=None):
    """Sum the accumulator buffers of the given actors over the ranks, in place, result on EVERY rank (NCCL)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for a in actor_ids:
        for t in buffer_tensors(engine, a):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
