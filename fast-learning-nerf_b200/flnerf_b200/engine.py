"""Host-side runtime of the training hot loop (run_nerf.py:470-516): one call = render coarse+fine, both
MSE losses, backward, (gradient all-reduce), Adam -- as a straight sequence of libflnerf.so kernels with no
autograd graph, no per-iteration D2H copy and no host synchronisation.  ``FusedAdam`` keeps
torch.optim.Adam's state_dict format (checkpoints stay interchangeable with the reference).
"""
import math
import os
from typing import List, Optional

import torch

from . import ops
from .lib import MLP_PARAMS, FlnerfError


def share_grad_bucket(nets) -> torch.Tensor:
    """Puts the flat gradients of all nets into ONE contiguous fp32 bucket so that data parallelism needs a
    single NCCL all-reduce per step."""
    dev = nets[0].flat_parameters().device
    bucket = torch.zeros(len(nets) * MLP_PARAMS, dtype=torch.float32, device=dev)
    for i, net in enumerate(nets):
        net._ensure_flat()
        net._flat_grad = bucket[i * MLP_PARAMS:(i + 1) * MLP_PARAMS]
        off = 0
        for p in net.parameters():
            n = p.numel()
            p.grad = net._flat_grad[off:off + n].view(p.shape)
            off += n
    return bucket


class FusedAdam(torch.optim.Adam):
    """torch.optim.Adam(lr, betas, eps=1e-8) (run_nerf.py:99) whose step() is one fused kernel per net over the
    flat parameter / gradient / moment buffers.  state_dict()/load_state_dict() are the stock ones: exp_avg and
    exp_avg_sq of every parameter are views into the flat moment buffers."""

    def __init__(self, params, nets, lr=5e-4, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, lr=lr, betas=betas, eps=eps)
        self._nets = list(nets)
        self._m, self._v, self._step = {}, {}, {}
        self._bind()

    def _bind(self):
        for net in self._nets:
            flat = net.flat_parameters()
            m = self._m.get(id(net))
            if m is None or m.device != flat.device:
                m, v = torch.zeros_like(flat), torch.zeros_like(flat)
                self._m[id(net)], self._v[id(net)] = m, v
            else:
                v = self._v[id(net)]
            step = None
            off = 0
            for p in net.parameters():
                n = p.numel()
                st = self.state[p]
                if "exp_avg" in st and st["exp_avg"].data_ptr() != m[off:off + n].data_ptr():
                    m[off:off + n].copy_(st["exp_avg"].reshape(-1))
                    v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                if step is None:
                    step = int(float(st["step"])) if "step" in st else 0
                st["exp_avg"] = m[off:off + n].view(p.shape)
                st["exp_avg_sq"] = v[off:off + n].view(p.shape)
                st["step"] = torch.tensor(float(step))      # one tensor per parameter, like stock Adam
                off += n
            self._step[id(net)] = step

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._bind()

    def state_dict(self):
        for net in self._nets:      # the fused step keeps ONE counter per net; publish it per parameter
            for p in net.parameters():
                self.state[p]["step"] = torch.tensor(float(self._step[id(net)]))
        return super().state_dict()

    def zero_grad(self, set_to_none: bool = True):
        for net in self._nets:
            net._grad_bucket().zero_()

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        for net in self._nets:
            self._step[id(net)] += 1
            ops.adam_step(net.flat_parameters(), self._m[id(net)], self._v[id(net)], net._grad_bucket(), float(g["lr"]),
                          b1, b2, g["eps"], self._step[id(net)])
            net.weights_version += 1

    def replayed(self):
        """Host bookkeeping for one optimiser step that ran on the device without this object (CUDA-graph replay)."""
        for net in self._nets:
            self._step[id(net)] += 1
            net.weights_version += 1


class Trainer:
    """The fused training step.  Rays come either from the caller (``step``) or from the GPU-resident quadtree
    index buffer (``step_from_tree``)."""

    def __init__(self, net_coarse, net_fine, optimizer: FusedAdam, H, W, K, near, far, N_samples=64, N_importance=128,
                 white_bkgd=True, perturb=1.0, lindisp=False, ndc=False, raw_noise_std=0.0, seed=0, world_size=1,
                 rank=0, graph=None):
        if N_importance <= 0 or net_fine is None:
            raise FlnerfError("Trainer implements the coarse+fine loop (N_importance > 0), like run_nerf.train()")
        self.nc, self.nf, self.opt = net_coarse, net_fine, optimizer
        self.H, self.W, self.K = int(H), int(W), K
        self.near, self.far = float(near), float(far)
        self.Nc, self.Nf = int(N_samples), int(N_importance)
        self.white, self.perturb, self.lindisp, self.ndc = bool(white_bkgd), float(perturb), bool(lindisp), bool(ndc)
        self.noise_std = float(raw_noise_std)
        self.seed, self.calls = int(seed), 0
        self.world, self.rank = int(world_size), int(rank)
        self.bucket = share_grad_bucket([net_coarse, net_fine])
        self.last = {}
        # one CUDA graph per full batch of step_from_tree: opt-in, because a replayed step returns
        # the SAME loss / output tensors every time (the graph's static outputs)
        self.use_graph = bool(int(os.environ.get("FLNERF_GRAPH", "0"))) if graph is None else bool(graph)
        # Single process by default: a CUDA graph that captured NCCL kernels keeps the communicator alive -- on 2 x B200 the
        # step replayed correctly (1.77 M vs 1.75 M rays/s, profiles/r02o_*) but destroy_process_group() never returned
        # unless every captured step had been dropped first (release_graph()).  Under data parallelism the step is therefore
        # launched kernel by kernel (costs ~1.5 %) unless $FLNERF_GRAPH_DP=1 opts in.
        if self.world > 1 and not int(os.environ.get("FLNERF_GRAPH_DP", "0")):
            self.use_graph = False
        self._graph = self._graph_key = self._graph_out = self._rec = None
        self._graph_seen, self._graph_launches = {}, 0
        self.sync_replicas()

    def release_graph(self):
        """Drop the captured step (and its private memory pool); the next full batch is captured again."""
        self._graph = self._graph_key = self._graph_out = None
        self._graph_seen = {}

    def sync_replicas(self):
        """nn.DataParallel held ONE copy of the weights (run_nerf.py:82,90).  Replicas built from unseeded RNGs or resumed
        from different files would never converge to each other (only gradients are exchanged), so rank 0's parameters,
        Adam moments and step counters are broadcast once -- at construction and after any checkpoint load."""
        if self.world <= 1 or not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return
        dist = torch.distributed
        for net in (self.nc, self.nf):
            flat = net.flat_parameters()
            dist.broadcast(flat, 0)
            net.weights_version += 1
            m, v = self.opt._m[id(net)], self.opt._v[id(net)]
            dist.broadcast(m, 0)
            dist.broadcast(v, 0)
            st = torch.tensor([float(self.opt._step[id(net)])], dtype=torch.float64, device=flat.device)
            dist.broadcast(st, 0)
            self.opt._step[id(net)] = int(st.item())
        lr = torch.tensor([float(self.opt.param_groups[0]["lr"])], dtype=torch.float64, device=self.bucket.device)
        dist.broadcast(lr, 0)
        for g in self.opt.param_groups:
            g["lr"] = float(lr.item())

    def global_loss(self, loss):
        """``step`` returns this rank's SHARE of the two MSEs (its rows, divided by the global ray count); their sum over
        ranks is the reference's img_loss / img_loss0.  One 8-byte all-reduce -- call it at log time only."""
        if self.world <= 1:
            return loss
        out = loss.clone()
        torch.distributed.all_reduce(out)
        return out

    # -- one MLP evaluation over [B,S] samples without autograd
    def _forward_net(self, net, rays11, z):
        B, S = z.shape
        flat, packed = net._weights()
        if net.mode == ops.MODE_FP32:
            x, dirpe = ops.encode_f32(rays11, z), None
        else:
            x, dirpe = ops.encode_tc(rays11, z, net.mode)
        raw, stash = ops.mlp_forward(net.mode, flat, packed, x, dirpe, B * S, S, True)
        return raw, (x, dirpe, stash)

    def _backward_net(self, net, saved, draw, n, S):
        x, dirpe, stash = saved
        flat, packed = net._weights()
        ops.mlp_backward(net.mode, flat, packed, x, dirpe, stash, draw, net._flat_grad, n, S)
        net._mask_grads()       # (a use_viewdirs=False model's frozen adapter; no-op otherwise)

    @torch.no_grad()
    def step(self, rays_o, rays_d, target, leaf_gid=None, leaf_max=None, global_batch: Optional[int] = None):
        """Returns loss[2] = (fine mse, coarse mse) as a DEVICE tensor (no sync); under data parallelism it is this
        rank's share (see ``global_loss``).  A rank whose share of a ragged last batch is EMPTY still joins the
        collectives and the optimiser step with a zero gradient."""
        B = rays_o.shape[0]
        Nc, Nf = self.Nc, self.Nf
        half = self.bucket.numel() // 2
        work = None
        if B == 0:
            self.opt.zero_grad()
            loss = torch.zeros(2, dtype=torch.float32, device=self.bucket.device)
        else:
            rays11 = ops.pack_rays(rays_o, rays_d, self.near, self.far, self.ndc, self.H, self.W, float(self.K[0][0]))
            off = self.calls
            self.calls += B * (Nc + Nf)
            z_c = ops.coarse_depths(rays11, Nc, self.perturb > 0, self.lindisp, None, self.seed, off)
            noise_c = noise_f = None
            if self.noise_std > 0:
                noise_c = torch.randn(B, Nc, device=rays11.device) * self.noise_std
                noise_f = torch.randn(B, Nc + Nf, device=rays11.device) * self.noise_std
            raw_c, sv_c = self._forward_net(self.nc, rays11, z_c)
            rgb0, _, _, w_c, _ = ops.composite_forward(raw_c, z_c, rays11[:, 3:6], noise_c, self.white, rays_d_stride=11)
            z_f, _, _ = ops.sample_pdf_merge(z_c, w_c, Nf, self.perturb == 0, None, self.seed + 1, off, want_samples=False)
            raw_f, sv_f = self._forward_net(self.nf, rays11, z_f)
            rgb, disp, acc, _, _ = ops.composite_forward(raw_f, z_f, rays11[:, 3:6], noise_f, self.white, rays_d_stride=11,
                                                         want_weights=False)
            denom = int(global_batch) if global_batch is not None else B * self.world
            loss, d_rgb, d_rgb0 = ops.mse_leafmax(rgb, rgb0, target, denom, leaf_gid, leaf_max)
            self.opt.zero_grad()
            draw_f = ops.composite_backward(raw_f, z_f, rays11[:, 3:6], noise_f, self.white, d_rgb, None, None, None,
                                            rays_d_stride=11)
            self._backward_net(self.nf, sv_f, draw_f, B * (Nc + Nf), Nc + Nf)
            if self.world > 1:
                # the fine net's half of the bucket is complete: its all-reduce (NCCL's own stream) overlaps the coarse
                # net's backward pass
                work = torch.distributed.all_reduce(self.bucket[half:], async_op=True)
            draw_c = ops.composite_backward(raw_c, z_c, rays11[:, 3:6], noise_c, self.white, d_rgb0, None, None, None,
                                            rays_d_stride=11)
            self._backward_net(self.nc, sv_c, draw_c, B * Nc, Nc)
            self.last = {"rgb": rgb, "rgb0": rgb0, "disp": disp, "acc": acc}
        if self.world > 1:
            if work is None:
                torch.distributed.all_reduce(self.bucket[half:])
            torch.distributed.all_reduce(self.bucket[:half])
            if work is not None:
                work.wait()
        self.opt.step()
        return loss

    @torch.no_grad()
    def step_from_tree(self, mgr, first, n_rand):
        """One batch of ``n_rand`` rows of the quadtree index buffer starting at ``first``; under data parallelism
        rank r consumes rows first + r, first + r + world, ... (every rank holds the same index buffer)."""
        rows = min(n_rand, mgr.n_rays - first)
        if self.use_graph and rows == n_rand and rows % self.world == 0:
            return self._graph_step(mgr, first, n_rand)
        local = (rows - self.rank + self.world - 1) // self.world
        o, d, tgt, gid = mgr.batch(first + self.rank, local, self.world)
        return self.step(o, d, tgt, gid, mgr.leaf_max, global_batch=rows)

    # -- the same step as ONE CUDA graph: the ~30 launches of a step carry no per-step host value once the batch start,
    #    the Philox offsets and the Adam scalars come from the device-side step record (include/flnerf.h)
    def _record_step(self, mgr, first, n_rand):
        g = self.opt.param_groups[0]
        t = self.opt._step[id(self.nc)] + 1
        ops.step_record_write(self._rec, first, self.calls, float(g["lr"]), g["betas"][0], g["betas"][1], t)

    def _step_on_record(self, mgr, n_rand):
        """step_from_tree with every per-step scalar taken from the attached record (host values are zero / biases)."""
        local = n_rand // self.world
        calls = self.calls
        self.calls = 0                      # Philox offset = rec.rng_offset + 0
        ops.set_step_record(self._rec)
        try:
            o, d, tgt, gid = mgr.batch(self.rank, local, self.world)        # rows rec.first + rank + k * world
            loss = self.step(o, d, tgt, gid, mgr.leaf_max, global_batch=n_rand)
        finally:
            ops.set_step_record(None, self.bucket.device)
            self.calls = calls
        return loss

    def _graph_step(self, mgr, first, n_rand):
        from . import lib
        key = (id(mgr), mgr.ray_pix.data_ptr(), mgr.ray_gid.data_ptr(), mgr.leaf_max.data_ptr(), n_rand,
               float(self.opt.param_groups[0]["betas"][0]), self.nc.mode, self.nf.mode)
        if self._rec is None:
            self._rec = torch.zeros(ops.STEP_RECORD_BYTES, dtype=torch.uint8, device=self.bucket.device)
        local = n_rand // self.world
        if self._graph_key != key:
            seen = self._graph_seen.get(key, 0)
            self._graph_seen[key] = seen + 1
            if seen == 0:
                # first batch of this shape: run it eagerly (lazy initialisation of the library, allocator warm-up)
                o, d, tgt, gid = mgr.batch(first + self.rank, local, self.world)
                return self.step(o, d, tgt, gid, mgr.leaf_max, global_batch=n_rand)
            # second batch: capture.  Capturing executes nothing, so the host counters are restored afterwards and the
            # step then runs as the first replay.
            steps = {id(n): self.opt._step[id(n)] for n in (self.nc, self.nf)}
            versions = (self.nc.weights_version, self.nf.weights_version)
            self._record_step(mgr, first, n_rand)
            torch.cuda.synchronize()
            c0 = lib.launch_count()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._step_on_record(mgr, n_rand)
            self._graph_launches = lib.launch_count() - c0
            lib.load().flnerf_launch_count_add(-self._graph_launches)         # captured, not launched
            for n in (self.nc, self.nf):
                self.opt._step[id(n)] = steps[id(n)]
            self.nc.weights_version, self.nf.weights_version = versions
            self._graph, self._graph_key, self._graph_out = graph, key, (out, dict(self.last))
        self._record_step(mgr, first, n_rand)
        self._graph.replay()
        lib.load().flnerf_launch_count_add(self._graph_launches)
        self.calls += local * (self.Nc + self.Nf)
        self.opt.replayed()
        loss, self.last = self._graph_out
        return loss


def lr_at(lrate, lrate_decay, global_iter):
    """run_nerf.py:498-502"""
    return lrate * (0.1 ** (global_iter / (lrate_decay * 1000)))
