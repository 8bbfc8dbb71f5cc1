"""Multi-rank host logic of the sharded Sobol path on CPU: world_size 2 and 4, gloo backend, the
oracle-backed stand-in for the kernels (tests/fake_hotpath.py).  Checks that the union of the
ranks' pools / nfev equals the single-rank result -- single rank is the oracle for multi rank."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, init_file, out_dir, n, d):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    from _pytest.monkeypatch import MonkeyPatch
    import fake_hotpath
    mpatch = MonkeyPatch()
    fake_hotpath.install(mpatch)
    from oracle import oracle as orc
    from shgo_b200 import dist as sd
    from shgo_b200 import hotpath as hp
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    try:
        bounds = [(-5.12, 5.12)] * d
        tab = hp.SobolTables(bounds)
        rp, col = orc.hypercube_csr(n)
        m = n.bit_length() - 1
        row0, rows = sd.shard_range(n, world, rank)
        rp_l = torch.from_numpy(rp[row0:row0 + rows + 1].copy())
        col_l = torch.from_numpy(col[row0 * m:(row0 + rows) * m].copy())
        it = sd.ShardedSobolIteration(1, 0, tab, n, rp_l, col_l, exchange="allgather")
        it.step()
        it.step()                       # a second iteration reuses the buffers
        pool = it.global_pool()
        nfev = it.global_nfev()
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), pool=pool, nfev=nfev)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_iteration_matches_single_rank(world):
    from oracle import oracle as orc
    n, d = 1 << 12, 6
    with tempfile.TemporaryDirectory() as tmp:
        init_file = os.path.join(tmp, "init")
        mp.spawn(_worker, args=(world, init_file, tmp, n, d), nprocs=world, join=True)
        z = np.load(os.path.join(tmp, "out.npz"))
    b = np.array([(-5.12, 5.12)] * d)
    f, feas = orc.sobol_eval(orc.sobol_dirnums(d), "rastrigin", None, 0, n, b[:, 0], b[:, 1])
    rp, col = orc.hypercube_csr(n)
    ref = orc.compact_pool(orc.star_csr(rp, col, f))
    assert np.array_equal(z["pool"], ref)
    assert int(z["nfev"]) == n


def _lattice_worker(rank, world, init_file, out_dir, d, gen):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    from _pytest.monkeypatch import MonkeyPatch
    import fake_hotpath
    fake_hotpath.install(MonkeyPatch())
    from shgo_b200 import dist as sd
    from shgo_b200 import lattice
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    try:
        bounds = [(-5.12, 5.12)] * d
        tab = torch.from_numpy(lattice.coordinate_tables(bounds, gen))
        it = sd.ShardedLatticeIteration(1, 0, d, gen, tab, exchange="allreduce", granule=4096)
        it.step()
        it.step()
        pool, nfev = it.global_pool(), it.global_nfev()
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), pool=pool, nfev=nfev, owned=int(it.owned().sum()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,d,gen", [(2, 5, 2), (4, 2, 6), (2, 3, 0), (3, 4, 3)])
def test_sharded_lattice_generation_matches_single_rank(world, d, gen):
    """SURVEY.md 8(e), lattice path: granule-cyclic ownership (4096 ids per granule: the cases
    span 2, 3 and 1 granules); the union of the ranks' pools equals the oracle's pool of the whole
    generation."""
    from oracle import oracle as orc
    with tempfile.TemporaryDirectory() as tmp:
        init_file = os.path.join(tmp, "init")
        mp.spawn(_lattice_worker, args=(world, init_file, tmp, d, gen), nprocs=world, join=True)
        z = np.load(os.path.join(tmp, "out.npz"))
    bounds = [(-5.12, 5.12)] * d
    x = orc.lattice_coords(bounds, gen)
    f, _ = orc.eval_points("rastrigin", None, x)
    ref = orc.compact_pool(orc.lattice_star(d, gen, f))
    assert np.array_equal(z["pool"], ref)
    assert int(z["nfev"]) == len(x)


def test_split_range_covers_everything():
    from shgo_b200.dist import split_range
    for n in (0, 1, 7, 1024, 1025):
        for world in (1, 2, 3, 8):
            parts = [split_range(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1


def test_shard_range():
    from shgo_b200.dist import shard_range
    assert [shard_range(1 << 10, 4, r) for r in range(4)] == [(0, 256), (256, 256), (512, 256), (768, 256)]
    with pytest.raises(ValueError):
        shard_range(10, 4, 0)
