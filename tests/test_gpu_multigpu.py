"""GPU tier, two devices (skipped on a one-GPU box): the N > 1 paths of GpuSimulationEngine on hardware.

* under a 2-rank NCCL process group (one process per GPU, the torchrun layout): histories split by the
  reference's rule, all_reduce of the voxel-record buffers, rank 0 holds the merged images -- they must equal the
  one-GPU run (results do not depend on the number of GPUs: same seed, global history ids) to the order of the
  fp64 additions;
* `sim.number_of_gpus = 2` inside one process (two engines, host-side sum).
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _two_gpus():
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.device_count() >= 2
    except Exception:
        return False


def _build(tmp, n, number_of_gpus=None):
    import opengate_b200 as gate
    from test_host_api import ct_sim
    sim, ct, src, dose, hu = ct_sim(tmp, n_primaries=n)
    sim.random_seed = 24681357
    if number_of_gpus:
        sim.number_of_gpus = number_of_gpus
    dose.dose.active = True
    dose.edep_uncertainty.active = True
    dose.counts.active = True
    plane = sim.add_volume("Box", "plane")
    plane.size = [600 * gate.g4_units.mm, 2 * gate.g4_units.mm, 600 * gate.g4_units.mm]
    plane.translation = [0, 300 * gate.g4_units.mm, 0]
    ph = sim.add_actor("PhaseSpaceActor", "phsp")
    ph.attached_to = plane
    ph.attributes = ["KineticEnergy", "PrePosition", "PreDirection", "Weight", "EventID", "TimeFromBeginOfEvent"]
    ph.write_to_disk = False
    stats = sim.add_actor("SimulationStatisticsActor", "stats")
    return sim, dose, ph, stats


def _results(out, dose, ph):
    r = {k: getattr(dose, k).image.array.copy() for k in ("edep", "edep_uncertainty", "dose", "counts")}
    r["histories"] = out.device_stats[0]["histories"]
    r["tentative"] = out.device_stats[0]["tentative_collisions"]
    r["rows"] = {k: v.copy() for k, v in ph.data.items()}
    return r


def _worker(rank, world, port, tmp, n, q):
    import pathlib
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    d = pathlib.Path(tmp) / f"rank{rank}"
    d.mkdir(exist_ok=True)
    sim, dose, ph, stats = _build(d, n)
    out = sim.run()
    if rank == 0:
        q.put(_results(out, dose, ph))
    dist.barrier()
    dist.destroy_process_group()


def _compare(a, b):
    assert a["histories"] == b["histories"] and a["tentative"] == b["tentative"]
    assert np.array_equal(a["counts"], b["counts"])
    for k in ("edep", "dose"):
        np.testing.assert_allclose(a[k], b[k], rtol=1e-12, atol=0, err_msg=k)
    np.testing.assert_allclose(a["edep_uncertainty"], b["edep_uncertainty"], rtol=1e-9, atol=0)
    assert set(a["rows"]) == set(b["rows"]) and len(a["rows"]["EventID"]) == len(b["rows"]["EventID"]) > 0
    for k in a["rows"]:
        assert np.array_equal(a["rows"][k], b["rows"][k]), k


@pytest.mark.skipif(not _two_gpus(), reason="needs two CUDA devices")
def test_two_rank_nccl_run_equals_one_gpu_run(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, HERE)
    n = 60001
    (tmp_path / "one").mkdir()
    sim, dose, ph, stats = _build(tmp_path / "one", n)
    one = _results(sim.run(), dose, ph)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), n, q)) for r in range(2)]
    for p in procs:
        p.start()
    two = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    _compare(one, two)


@pytest.mark.skipif(not _two_gpus(), reason="needs two CUDA devices")
def test_number_of_gpus_in_one_process_equals_one_gpu_run(tmp_path):
    sys.path.insert(0, HERE)
    n = 60001
    res = []
    for sub, g in (("one", None), ("two", 2)):
        (tmp_path / sub).mkdir()
        sim, dose, ph, stats = _build(tmp_path / sub, n, number_of_gpus=g)
        res.append(_results(sim.run(), dose, ph))
        if g:
            assert stats.counts["nb_gpus"] == 2
    _compare(*res)
