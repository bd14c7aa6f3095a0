"""N>1 host logic on CPU: two gloo ranks shard the histories of one run by global history id
(split rule of the reference, sources/base.py:190-197), reduce their grids, and must reproduce the
single-rank result.  The transport itself is done by the CPU oracle (test infrastructure)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from opengate_b200 import abi, multigpu
from opengate_b200.sources import scale_counts_for_rank

HERE = os.path.dirname(os.path.abspath(__file__))


def test_split_rule_matches_reference_formula():
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            parts = [multigpu.split_primaries(n, world, r) for r in range(world)]
            assert sum(c for _, c in parts) == n
            assert [c for _, c in parts] == [int(scale_counts_for_rank(n, r, world)) for r in range(world)]
            # contiguous, ordered id ranges
            off = 0
            for f, c in parts:
                assert f == off
                off += c


def _worker(rank, world, port, n, q):
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import TransportOracle
    from small_setups import c3
    S = c3(n)
    o = TransportOracle(S.seed)
    sids, dids, _ = S.apply(o)
    o.begin_run()
    first, cnt = multigpu.split_primaries(n, world, rank)
    o.run_source(sids[0], first, cnt, 0)
    grids = {}
    for w in multigpu.active_grids(S.dose[0].score_flags):
        t = torch.from_numpy(np.array(o.dose_grid(0, w), dtype=np.float64, copy=True))
        dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
        grids[w] = t.numpy()
    st = multigpu.reduce_counters(o.stats(), "cpu")
    if rank == 0:
        q.put((grids, st))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_run_equals_single_rank():
    n = 6001
    sys.path.insert(0, HERE)
    from oracle.oracle import TransportOracle
    from small_setups import c3
    S = c3(n)
    ref = TransportOracle(S.seed)
    sids, _, _ = S.apply(ref)
    ref.begin_run()
    ref.run_source(sids[0], 0, n, 0)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    grids, st = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a = ref.stats()
    for k in ("histories", "tentative_collisions", "n_phot", "n_compt", "n_rayl", "deposits", "flushed_entries", "escaped"):
        assert st[k] == a[k], k
    for w, g in grids.items():
        want = np.array(ref.dose_grid(0, w), dtype=np.float64)
        if w == abi.GRID_COUNTS:
            assert np.array_equal(g, want)
        else:
            np.testing.assert_allclose(g, want, rtol=1e-12, atol=0)


def test_record_layout_follows_the_device_rule():
    """Voxel records of a dose actor: the scored ones of edep, edep2, dose, dose2 in that order, 1, 2 or 4
    doubles per voxel (gate_b200_add_dose_actor); the host side of the multi-GPU reduce relies on it."""
    from opengate_b200 import abi
    E, E2, D, D2, C = abi.SCORE_EDEP, abi.SCORE_EDEP_SQ, abi.SCORE_DOSE, abi.SCORE_DOSE_SQ, abi.SCORE_COUNTS
    assert multigpu.record_layout(0) == (1, {abi.GRID_EDEP: 0})                       # edep is always scored
    assert multigpu.record_layout(E | C) == (1, {abi.GRID_EDEP: 0})
    assert multigpu.record_layout(E | E2) == (2, {abi.GRID_EDEP: 0, abi.GRID_EDEP_SQ: 1})
    assert multigpu.record_layout(E | D) == (2, {abi.GRID_EDEP: 0, abi.GRID_DOSE: 1})
    assert multigpu.record_layout(E | E2 | D) == (4, {abi.GRID_EDEP: 0, abi.GRID_EDEP_SQ: 1, abi.GRID_DOSE: 2})
    assert multigpu.record_layout(E | E2 | D | D2 | C)[0] == 4
    assert multigpu.record_layout(E | D | D2)[1] == {abi.GRID_EDEP: 0, abi.GRID_DOSE: 1, abi.GRID_DOSE_SQ: 2}


# ------------------------------------------------------------------------------ the engine's own N>1 path

def _engine_worker(rank, world, port, tmp, q):
    """Simulation.run() under a 2-rank gloo group with random_seed='auto' and an activity source: the ranks
    must agree on ONE seed (rank 0's) and on the Poisson event count they split, and the merged result must be
    what a single process gives with that seed."""
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import random
    random.seed(1000 + rank)                      # 'auto' would draw different seeds on the two ranks
    import opengate_b200 as gate
    from oracle_backend import OracleBackend
    from test_host_api import ct_sim

    class Backend(OracleBackend):
        distributed = True

    import pathlib
    d = pathlib.Path(tmp) / f"rank{rank}"
    d.mkdir(exist_ok=True)
    sim, ct, src, dose, hu = ct_sim(d, n_primaries=0)
    sim.random_seed = "auto"
    src.number_of_primaries = 0
    src.activity = 5000 * gate.g4_units.Bq
    sim.run_timing_intervals = [[0, 1 * gate.g4_units.s]]
    dose.edep_uncertainty.active = True
    out = sim.run(_backend=Backend)
    q.put((rank, int(sim.current_random_seed), int(out.device_stats[0]["histories"]),
           dose.edep.image.array.copy() if rank == 0 else None, dose.edep_uncertainty.image.array.copy() if rank == 0 else None))
    dist.barrier()
    dist.destroy_process_group()


def test_engine_under_a_process_group_shares_seed_and_counts(tmp_path):
    sys.path.insert(0, HERE)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_engine_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, seed0, n0, edep, unc), (_, seed1, n1, _, _) = res
    assert seed0 == seed1                          # one seed for the whole job
    assert n0 == n1 and 300 < n0 < 550             # one Poisson draw (activity x spectrum-weight sum, GateGenericSource.cpp:622-626), split over the ranks
    # the single-process run with that seed gives the same images
    import opengate_b200 as gate
    from oracle_backend import OracleBackend
    from test_host_api import ct_sim
    (tmp_path / "single").mkdir()
    sim, ct, src, dose, hu = ct_sim(tmp_path / "single", n_primaries=0)
    sim.random_seed = seed0
    src.number_of_primaries = 0
    src.activity = 5000 * gate.g4_units.Bq
    sim.run_timing_intervals = [[0, 1 * gate.g4_units.s]]
    dose.edep_uncertainty.active = True
    out = sim.run(_backend=OracleBackend)
    assert out.device_stats[0]["histories"] == n0
    np.testing.assert_allclose(edep, dose.edep.image.array, rtol=1e-12, atol=0)
    np.testing.assert_allclose(unc, dose.edep_uncertainty.image.array, rtol=1e-9, atol=0)


# ------------------------------------------------------------------------------ shared host arrays of a node

def _shared_worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeEngine:
        """holds the 'all-reduced' grids; dose_grid_range as core.Engine implements it"""
        def __init__(self, grids):
            self.g = grids

        def dose_grid_range(self, actor, which, first, count, out, scale=1.0):
            out.reshape(-1)[first:first + count] = self.g[which].reshape(-1)[first:first + count] * scale
            return out

    shapes = [(5, 7, 3), (5, 7, 3)]
    grids = {0: np.arange(105, dtype=np.float64).reshape(5, 7, 3), 2: 1000.0 + np.arange(105, dtype=np.float64).reshape(5, 7, 3)}
    sh = multigpu.SharedHostGrids(shapes, pin=False)
    arrs = sh.gather(FakeEngine(grids), 0, [0, 2], scales=[1.0, 0.5])
    ok = np.array_equal(arrs[0], grids[0]) and np.array_equal(arrs[1], grids[2] * 0.5)
    q.put((rank, bool(ok), sh.slice_of(0)))
    sh.close()
    dist.destroy_process_group()


def test_shared_host_grids_assemble_the_slices_of_all_ranks():
    """multigpu.SharedHostGrids: three ranks each copy their slice of two grids into the POSIX shared arrays; every
    rank then sees the complete arrays; the slices partition the voxels."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_shared_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(3))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert [s for _, _, s in res] == [(0, 35), (35, 35), (70, 35)]
