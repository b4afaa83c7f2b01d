"""N > 1 host-side logic on CPU (gloo, world_size 2/3): the image sharding rule of the library (jaicov_shard_images,
a pure host function) and the sharding algebra of the assembly -- the all-reduced sum of the per-rank normal
equations (each rank stacks only its own observation range, here with the oracle as the stand-in for the device
sweeps) equals the unsharded system."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import bundle_adjustment_b200 as ba


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import ctypes
    from oracle.oracle import FlatProblem, lib
    from tests.scenes import synthetic_scene
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    scene, _ = synthetic_scene(4, images=11, targets=40, visibility=0.7)
    fp = FlatProblem(scene)
    b, e = ba._lib.shard_images(fp.pt_ptr, world, rank)
    j0, j1 = int(fp.pt_ptr[b]), int(fp.pt_ptr[e])
    n = fp.n
    N = np.zeros(n * (n + 1) // 2)
    nv = np.zeros(n)
    p = fp.cstruct()
    lib().orc_stack_image_points(ctypes.byref(p), fp.bk.sigma2apriori, N.ctypes.data, nv.ctypes.data, j0, j1)
    tN, tn = torch.from_numpy(N), torch.from_numpy(nv)
    dist.all_reduce(tN)
    dist.all_reduce(tn)
    cover = torch.tensor([float(j1 - j0)], dtype=torch.float64)
    dist.all_reduce(cover)
    if rank == 0:
        Nf = np.zeros_like(N)
        nf = np.zeros_like(nv)
        lib().orc_stack_image_points(ctypes.byref(p), fp.bk.sigma2apriori, Nf.ctypes.data, nf.ctypes.data, 0, fp.m)
        idx = np.arange(n)
        dg = np.sqrt(np.abs(Nf[idx + idx * (idx + 1) // 2]))
        dg[dg == 0] = 1
        iu = np.triu_indices(n)
        k = iu[0] + iu[1] * (iu[1] + 1) // 2
        q.put((float((np.abs(N[k] - Nf[k]) / (dg[iu[0]] * dg[iu[1]])).max()), float(cover[0]), fp.m,
               float(np.abs(nv - nf).max() / np.abs(nf).max())))
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_assembly_algebra(built, world):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    errN, cover, m, errn = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert cover == m            # every observation belongs to exactly one rank
    assert errN < 1e-13 and errn < 1e-12


def test_shard_rule(built):
    pt_ptr = np.array([0, 10, 10, 35, 60, 61, 100])
    for world in (1, 2, 3, 4, 8):
        ranges = [ba._lib.shard_images(pt_ptr, world, r) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == 6
        for (b0, e0), (b1, e1) in zip(ranges, ranges[1:]):
            assert e0 == b1 and b0 <= e0
    assert [ba._lib.shard_images(pt_ptr, 2, r) for r in range(2)] == [(0, 4), (4, 6)]   # first image boundary at or after 50 of 100 observations
