"""CPU, world_size 2 over gloo: the data-parallel contract of SURVEY 8(e) -- every rank takes rows r, r+world, ...
of each N_rand-row batch, the loss is normalised by the GLOBAL ray count, and ONE all-reduce(SUM) of the flat
gradient bucket reproduces the single-process gradient; the per-leaf table is reduced with MAX."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import nerf_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    B, rows = 10, 7                                   # ragged last batch: 7 of N_rand=10 rows
    rays = torch.cat([torch.zeros(B, 3), torch.nn.functional.normalize(torch.randn(B, 3), dim=-1),
                      2 * torch.ones(B, 1), 6 * torch.ones(B, 1), torch.nn.functional.normalize(torch.randn(B, 3), dim=-1)], -1)
    tgt = torch.rand(B, 3)
    gid = torch.randint(0, 4, (B,))
    pc = {k: v.clone().requires_grad_(True) for k, v in O.init_params(1).items()}
    local = list(range(rank, rows, world))            # rows first + r, first + r + world, ...
    o = O.render_rays(rays[local], pc, None, 8, 0, white_bkgd=True)
    # local contribution to the GLOBAL mean: sum over local rays / (3 * rows)
    loss = ((o["rgb_map"] - tgt[local]) ** 2).sum() / (3 * rows)
    loss.backward()
    bucket = torch.cat([p.grad.reshape(-1) for p in pc.values()])
    dist.all_reduce(bucket)                           # the one collective of the step
    table = torch.full((4,), -1.0)
    stat = (tgt[local] - o["rgb_map"].detach()).abs().max(-1)[0]
    for g_, s_ in zip(gid[local].tolist(), stat.tolist()):
        table[g_] = max(float(table[g_]), s_)
    dist.all_reduce(table, op=dist.ReduceOp.MAX)
    lsum = loss.detach().clone()
    dist.all_reduce(lsum)
    if rank == 0:
        # single-process reference on the same 7 rows
        pr = {k: v.clone().requires_grad_(True) for k, v in O.init_params(1).items()}
        o1 = O.render_rays(rays[:rows], pr, None, 8, 0, white_bkgd=True)
        l1 = O.mse(o1["rgb_map"], tgt[:rows])
        l1.backward()
        ref = torch.cat([p.grad.reshape(-1) for p in pr.values()])
        t1 = torch.full((4,), -1.0)
        s1 = (tgt[:rows] - o1["rgb_map"].detach()).abs().max(-1)[0]
        for g_, s_ in zip(gid[:rows].tolist(), s1.tolist()):
            t1[g_] = max(float(t1[g_]), s_)
        out.put((float((bucket - ref).abs().max()), float(ref.abs().max()), float(abs(lsum - l1)),
                 bool(torch.equal(table, t1)), sorted(local)))
    dist.barrier()
    dist.destroy_process_group()


def test_ray_sharding_and_single_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, scale, lerr, table_ok, local0 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert local0 == [0, 2, 4, 6]
    assert err <= 1e-6 * max(scale, 1.0) + 1e-7, (err, scale)
    assert lerr < 1e-7 and table_ok
