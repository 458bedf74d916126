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


class _HostNet(torch.nn.Module):
    """The flat-storage contract of model.NeRF (flat_parameters / _ensure_flat / _flat_grad / weights_version / mode) on CPU
    tensors: lets the PRODUCT's Trainer / FusedAdam host logic run under gloo without a GPU (no kernel is launched)."""

    def __init__(self, seed):
        super().__init__()
        from flnerf_b200.lib import MLP_PARAMS
        g = torch.Generator().manual_seed(seed)
        self._flat = torch.randn(MLP_PARAMS, generator=g)
        self.w = torch.nn.Parameter(self._flat.view(-1))
        self._flat_grad = torch.zeros(MLP_PARAMS)
        self.weights_version, self.mode = 0, 1

    def _ensure_flat(self):
        pass

    def flat_parameters(self):
        return self.w.data

    def _grad_bucket(self):
        return self._flat_grad


def _trainer_worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "fast-learning-nerf_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from flnerf_b200.engine import FusedAdam, Trainer
    nc, nf = _HostNet(100 + rank), _HostNet(200 + rank)           # every rank starts from DIFFERENT weights
    opt = FusedAdam(list(nc.parameters()) + list(nf.parameters()), [nc, nf], lr=5e-4 * (rank + 1))
    opt._m[id(nc)].fill_(float(rank + 1)); opt._step[id(nf)] = 7 * (rank + 1)
    before = nc.flat_parameters().clone()
    tr = Trainer(nc, nf, opt, 100, 100, np.eye(3), 2.0, 6.0, 64, 128, world_size=world, rank=rank, graph=True)
    w = torch.cat([nc.flat_parameters(), nf.flat_parameters(), opt._m[id(nc)]])
    gathered = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(gathered, w)
    loss = tr.global_loss(torch.tensor([0.25, 0.5]) * (rank + 1))
    # the batch-sharding arithmetic of step_from_tree: rows first + r, first + r + world, ... of a ragged 7-row tail
    local = (7 - rank + world - 1) // world
    out.put((rank, all(torch.equal(gathered[0], g) for g in gathered), bool(torch.equal(before, nc.flat_parameters())),
             opt._step[id(nf)], opt.param_groups[0]["lr"], loss.tolist(), tr.use_graph, local, tr.bucket.numel()))
    dist.barrier()
    dist.destroy_process_group()


def test_product_trainer_syncs_replicas_and_reduces_the_logged_loss():
    """engine.Trainer (the product) under world_size 2 on CPU: construction broadcasts rank 0's weights, Adam moments, step
    counters and learning rate (nn.DataParallel had ONE copy: run_nerf.py:82,90); global_loss sums the per-rank shares; CUDA
    graphs are switched off under data parallelism."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_trainer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    (_, same0, kept0, step0, lr0, loss0, g0, n0, nb), (_, same1, kept1, step1, lr1, loss1, g1, n1, _) = res
    assert same0 and same1 and kept0 and not kept1                 # everybody holds rank 0's weights and moments
    assert step0 == step1 == 7 and lr0 == lr1 == 5e-4
    assert loss0 == loss1 == [0.75, 1.5]
    assert not g0 and not g1 and (n0, n1) == (4, 3) and nb == 2 * 595844
