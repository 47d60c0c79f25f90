"""N>1 host logic on CPU: world_size-2 gloo.  Each rank migrates its shard of the shots (here
with the oracle standing in for the GPU engine, tiny NT), the partial stacks are summed with
the same single reduce bench.py/driver use, and rank 0 finalises the image."""
import dataclasses
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import rtm_gpu_b200 as R
from rtm_gpu_b200.dist import finalize, partition_shots, reduce_stack


def test_partition_covers_all_shots_once():
    for nrec in (1, 2, 7, 64, 240):
        for world in (1, 2, 3, 8):
            got = [m for r in range(world) for m in partition_shots(nrec, world, r)]
            assert got == list(range(nrec))
            sizes = [len(partition_shots(nrec, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(__file__))
    import oraclelib as O
    from golden_cases import GOLDEN_CASES
    from refcase import data_tiny, velocity_tiny
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = dataclasses.replace(GOLDEN_CASES["tiny_te_compen"], NT1=60, nrec=3, depths=[300.0, 500.0, 700.0])
    v = R.pad_velocity(velocity_tiny(case), case.N2, case.ifv)
    vmin, vmax, _, _ = R.velocity_bins(v, case.dv)
    c = R.taylor_operator(case.nfdmax)
    p = O.make_params(case, vmin, vmax, contract=1)
    n = case.mod_NX * case.mod_NZ
    part = torch.zeros(2 * n, dtype=torch.float32)
    mine = partition_shots(case.nrec, world, rank)
    imgs = {}
    for m in mine:
        up, down, *_ = O.migrate_shot(p, v, c, None, case.r_u[m], case.r_x0, data_tiny(case, case.depths[m])[:, :60])
        part[:n] += torch.from_numpy(up.ravel())
        part[n:] += torch.from_numpy(down.ravel())
        imgs[m] = (up, down)
    allimgs = [None] * world
    dist.all_gather_object(allimgs, imgs)
    total = reduce_stack(part, len(mine), dst=0)
    if rank == 0:
        img, ill = finalize(part, total, case.iNorm, case.mod_NX, case.mod_NZ)
        merged = {}
        for d in allimgs:
            merged.update(d)
        ups = [merged[m][0] for m in range(case.nrec)]
        downs = [merged[m][1] for m in range(case.nrec)]
        ref, refd = O.stack(ups, downs, case.iNorm)
        q.put((total, float(np.abs(img - ref).max() / np.abs(ref).max()), float(np.abs(ill - refd).max() / np.abs(refd).max())))
    dist.destroy_process_group()


def test_two_rank_stack_matches_serial_stack():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    total, e_img, e_ill = q.get(timeout=10)
    assert total == 3
    # only the summation order differs from the serial m=0..nrec-1 loop: a few ulp
    assert e_img < 5e-6 and e_ill < 1e-6
