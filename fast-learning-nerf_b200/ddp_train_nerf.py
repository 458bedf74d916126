"""Drop-in for the training side of nerf++-ours/ddp_train_nerf.py (BASELINE configs[4]): ``config_parser``,
``create_nerf(rank, args) -> (start, models)``, ``train_step(models, rays_o, rays_d, target_rgb, args)`` and the epoch loop
``ddp_train_nerf(args)`` with the quadtree ray selector of the fork (prob=True pixel sampling, refinement on the MEAN leaf loss:
nerf++-ours/tree.py:567-584, 622).  ``models`` keeps the reference's layout -- 'cascade_level', 'cascade_samples', 'net_m',
'optim_m' -- and its checkpoint file ``model_{epoch:04d}.pth`` with ``net_m`` / ``optim_m`` state dicts under the reference's
parameter names (``module.nerf_net.{fg,bg}_net.base_layers.i.0.weight`` ..., stock Adam state in MLPNet registration order), so
checkpoints move both ways.  All arithmetic runs in libflnerf.so through flnerf_b200.nerfpp (foreground + 84-channel background
MLPs on tensor cores or fp32 CUDA cores, inverted-sphere sampling, fg/bg compositing, count-based resampling).

Differences: nn.DataParallel (ddp_train_nerf.py:153) becomes one process per GPU with one gradient all-reduce per cascade level;
data comes either from the caller (rays) or from ``dataset_type = synthetic`` (cameras inside the unit sphere, as
intersect_sphere requires: ddp_train_nerf.py:65-66); the reference's data_loader_split.py file loaders are out of scope.
"""
import argparse
import os
import time
from collections import OrderedDict

import numpy as np
import torch

from argument_parser import ConfigFileParser
from flnerf_b200 import nerfpp, ops, synthetic
from flnerf_b200.lib import FlnerfError
from tree import QuadTreeManager

# MLPNet registers base_layers, sigma_layers, base_remap_layers, rgb_layers (nerf_network.py:86-118); the flat buffer follows
# the nerf-ours order (nerfpp._ORDER).  _REG lists the reference names in registration order = stock Adam's parameter order.
_REG = ["base_layers.%d.0" % i for i in range(8)] + ["sigma_layers.0", "base_remap_layers.0", "rgb_layers.0", "rgb_layers.2"]


def _shapes(in_pts):
    sh = OrderedDict()
    for i in range(8):
        k = in_pts if i == 0 else (256 + in_pts if i == 5 else 256)
        sh["base_layers.%d.0" % i] = (256, k)
    sh["sigma_layers.0"], sh["base_remap_layers.0"], sh["rgb_layers.0"], sh["rgb_layers.2"] = (1, 256), (256, 256), (128, 283), (3, 128)
    return sh


class NerfNetModule:
    """One cascade level: NerfNetWithAutoExpo(args) without auto-exposure (ddp_model.py:146-176) as two flat parameter
    buffers + gradients in ONE contiguous bucket; state_dict() speaks the reference's names."""

    def __init__(self, args, device, seed):
        self.device = device
        self.n_fg = int(ops.L.load().flnerf_mlp_param_count_g(63, 27))
        self.n_bg = int(ops.L.load().flnerf_mlp_param_count_g(84, 27))
        g = torch.Generator().manual_seed(seed)
        self.flat = torch.cat([self._init(63, g), self._init(84, g)]).to(device)
        self.grad = torch.zeros_like(self.flat)
        self.net = nerfpp.NerfNet(self.flat[:self.n_fg], self.flat[self.n_fg:], precision=args.precision or "bf16")
        self.net.grad_fg, self.net.grad_bg = self.grad[:self.n_fg], self.grad[self.n_fg:]

    @staticmethod
    def _init(in_pts, g):
        """nn.Linear's default initialisation (kaiming_uniform(a=sqrt 5) = U(+-1/sqrt(fan_in)) for weight and bias), in the flat
        nerf-ours order."""
        sh = _shapes(in_pts)
        parts = []
        for a, _ in nerfpp._ORDER:
            out, k = sh[a]
            bound = 1.0 / np.sqrt(k)
            parts.append((torch.rand(out * k, generator=g) * 2 - 1) * bound)
            parts.append((torch.rand(out, generator=g) * 2 - 1) * bound)
        return torch.cat(parts).float()

    def _named(self, flat_fg, flat_bg):
        out = OrderedDict()
        for tag, flat, in_pts in (("fg_net", flat_fg, 63), ("bg_net", flat_bg, 84)):
            like = {a + s: torch.empty(sh if s == ".weight" else sh[:1]) for a, sh in _shapes(in_pts).items() for s in (".weight", ".bias")}
            named = nerfpp.mlpnet_from_flat(flat, like)
            for a in _REG:
                for s in (".weight", ".bias"):
                    out["module.nerf_net.%s.%s%s" % (tag, a, s)] = named[a + s]
        return out

    def state_dict(self):
        return OrderedDict((k, v.detach().clone()) for k, v in self._named(self.flat[:self.n_fg], self.flat[self.n_fg:]).items())

    def load_state_dict(self, sd):
        with torch.no_grad():
            for k, v in self._named(self.flat[:self.n_fg], self.flat[self.n_fg:]).items():
                v.copy_(sd[k].to(self.device))

    def parameters(self):
        return list(self._named(self.flat[:self.n_fg], self.flat[self.n_fg:]).values())


class FlatAdamOptim:
    """torch.optim.Adam(net.parameters(), lr) (ddp_train_nerf.py:155) over the level's flat buffer; state_dict() is stock Adam's."""

    def __init__(self, module: NerfNetModule, lr):
        self.module = module
        self.adam = nerfpp.FlatAdam([module.flat], lr=lr)

    def zero_grad(self):
        self.module.grad.zero_()

    def step(self):
        self.adam.step([self.module.grad])

    def state_dict(self):
        m = self.module
        state = {}
        if self.adam.t > 0:
            ms = list(m._named(self.adam.m[0][:m.n_fg], self.adam.m[0][m.n_fg:]).values())
            vs = list(m._named(self.adam.v[0][:m.n_fg], self.adam.v[0][m.n_fg:]).values())
            for i, (a, b) in enumerate(zip(ms, vs)):
                state[i] = {"step": torch.tensor(float(self.adam.t)), "exp_avg": a.detach().clone(), "exp_avg_sq": b.detach().clone()}
        group = {"lr": self.adam.lr, "betas": tuple(self.adam.betas), "eps": self.adam.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(2 * 2 * len(_REG)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        m = self.module
        self.adam.lr = float(sd["param_groups"][0]["lr"])
        st = sd["state"]
        if len(st) == 0:
            return
        ms = list(m._named(self.adam.m[0][:m.n_fg], self.adam.m[0][m.n_fg:]).values())
        vs = list(m._named(self.adam.v[0][:m.n_fg], self.adam.v[0][m.n_fg:]).values())
        with torch.no_grad():
            for i, (a, b) in enumerate(zip(ms, vs)):
                a.copy_(st[i]["exp_avg"].to(m.device)); b.copy_(st[i]["exp_avg_sq"].to(m.device))
        self.adam.t = int(float(st[0]["step"]))


def config_parser():
    """The flags of ddp_train_nerf.py:430-500 this path reads (same names and defaults), + dataset_type / precision."""
    p = ConfigFileParser()
    p.add_argument("--config", type=str, default=None, help="config file path")
    p.add_argument("--expname", type=str, default="nerfpp")
    p.add_argument("--basedir", type=str, default="./logs/")
    p.add_argument("--datadir", type=str, default=None)
    p.add_argument("--scene", type=str, default=None)
    p.add_argument("--netdepth", type=int, default=8)
    p.add_argument("--netwidth", type=int, default=256)
    p.add_argument("--use_viewdirs", action="store_true")
    p.add_argument("--init_level", type=int, default=3)
    p.add_argument("--subdivide_every", type=int, default=1)
    p.add_argument("--subdivide_thres", type=float, default=0.015)
    p.add_argument("--rays_downscale", type=int, default=1)
    p.add_argument("--randSamp_perc", type=float, default=0.5)
    p.add_argument("--dset_name", type=str, default="Truck")
    p.add_argument("--no_reload", action="store_true")
    p.add_argument("--ckpt_path", type=str, default=None)
    p.add_argument("--N_rand", type=int, default=32 * 32 * 2)
    p.add_argument("--chunk_size", type=int, default=1024 * 8)
    p.add_argument("--batch_size", type=int, default=2880)
    p.add_argument("--n_epoch", type=int, default=6)
    p.add_argument("--cascade_level", type=int, default=2)
    p.add_argument("--cascade_samples", type=str, default="64,64")
    p.add_argument("--optim_autoexpo", action="store_true")
    p.add_argument("--lrate", type=float, default=5e-4)
    p.add_argument("--max_freq_log2", type=int, default=10)
    p.add_argument("--max_freq_log2_viewdirs", type=int, default=4)
    p.add_argument("--dataset_type", type=str, default="synthetic", help="flnerf addition: 'synthetic' needs no files")
    p.add_argument("--precision", type=str, default=None, help="flnerf addition: bf16 | bf16x3 | fp32")
    p.add_argument("--no_graph", action="store_true", help="flnerf addition: launch every kernel of a batch from the host")
    return p


def create_nerf(rank, args):
    """ddp_train_nerf.py:134-181 -> (start, models)."""
    if args.netdepth != 8 or args.netwidth != 256 or args.max_freq_log2 != 10 or args.max_freq_log2_viewdirs != 4:
        raise FlnerfError("the flnerf nerf++ kernels implement netdepth 8, netwidth 256, max_freq_log2 10 / 4 (tat_training_truck.txt)")
    if args.optim_autoexpo:
        raise FlnerfError("optim_autoexpo is not implemented")
    device = torch.device("cuda", rank)
    models = OrderedDict()
    models["cascade_level"] = args.cascade_level
    models["cascade_samples"] = [int(x.strip()) for x in args.cascade_samples.split(",")]
    for m in range(models["cascade_level"]):
        net = NerfNetModule(args, device, seed=777 + m)      # one seed for every process (ddp_train_nerf.py:137)
        models["net_%d" % m] = net
        models["optim_%d" % m] = FlatAdamOptim(net, args.lrate)
    start = 0
    d = os.path.join(args.basedir, args.expname)
    if args.ckpt_path is not None and os.path.isfile(args.ckpt_path):
        ckpts = [args.ckpt_path]
    else:
        ckpts = [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(".pth")] if os.path.isdir(d) else []
    path2iter = lambda p: int(os.path.basename(p)[:-4].rsplit("_", 1)[1])
    ckpts = sorted(ckpts, key=path2iter)
    print("Found ckpts: {}".format(ckpts))
    if len(ckpts) > 0 and not args.no_reload:
        print("Reloading from: {}".format(ckpts[-1]))
        start = path2iter(ckpts[-1])
        to_load = torch.load(ckpts[-1], map_location=device, weights_only=False)
        for m in range(models["cascade_level"]):
            for name in ("net_%d" % m, "optim_%d" % m):
                models[name].load_state_dict(to_load[name])
    return start, models


@torch.no_grad()
def cascade_batch(nets, optims, samples, o, d, gt, global_batch, seed, offset, world=1):
    """The body of the reference's batch loop (ddp_train_nerf.py:346-404): for each cascade level place the samples (level 0:
    sphere intersection + stratified jitter; level 1: count-based inverse CDF of level 0's fg / bg weights, merged), render
    through the level's NerfNet, MSE against the batch, backward, (all-reduce), Adam.  Returns (losses[L] on the device, the
    last level's outputs)."""
    losses, ret, fg_z, bg_z, fg_far = [], None, None, None, None
    for m in range(len(nets)):
        if m == 0:
            fg_far, fg_z, bg_z = ops.pp_depths0(o, d, samples[0], True, None, None, seed, offset)
        else:
            fg_z, _ = ops.pp_sample_pdf_merge(fg_z, ret["fg_weights"], samples[1], False, None, seed + 1, offset)
            bg_z, _ = ops.pp_sample_pdf_merge(bg_z, ret["bg_weights"], samples[1], False, None, seed + 2, offset)
        optims[m].zero_grad()
        ret = nets[m].net.forward(o, d, fg_far, fg_z, bg_z)
        loss, d_rgb, _ = ops.mse_leafmax(ret["rgb"], None, gt, global_batch)        # mean over the GLOBAL batch
        nets[m].net.backward(d_rgb)
        if world > 1:
            torch.distributed.all_reduce(nets[m].grad)
        optims[m].step()
        losses.append(loss[0:1])
    return losses, ret


class CascadeStep:
    """cascade_batch for repeated batches of one shape.  Single process: from the second full batch on, the whole batch -- both
    cascade levels, ~45 launches -- is ONE CUDA graph: the rays are copied into static buffers, the
    Philox offset and Adam's bias-corrected scalars come from the device-side step record.  The returned losses / outputs are
    then the graph's static tensors (the same memory every call).  Data-parallel runs and odd batches take the eager path."""

    def __init__(self, nets, optims, samples, seed, world=1, graph=True):
        self.nets, self.optims, self.samples, self.seed, self.world = nets, optims, list(samples), int(seed), int(world)
        self.use_graph = bool(graph) and self.world == 1
        self._graph = self._key = self._rec = self._out = self._in = None
        self._seen, self._launches = {}, 0

    def _adam(self):
        a = self.optims[0].adam
        same = all(o.adam.t == a.t and o.adam.lr == a.lr and tuple(o.adam.betas) == tuple(a.betas) for o in self.optims)
        return a if same else None

    def __call__(self, o, d, gt, global_batch, offset):
        a = self._adam() if self.use_graph else None
        if a is None:
            return cascade_batch(self.nets, self.optims, self.samples, o, d, gt, global_batch, self.seed, offset, self.world)
        from flnerf_b200 import lib
        key = (o.shape[0], int(global_batch), self.seed, tuple(a.betas), tuple(n.net.mode for n in self.nets))
        if self._key != key:
            seen = self._seen.get(key, 0)
            self._seen[key] = seen + 1
            if seen == 0:            # first batch of this shape / seed: eager (lazy initialisation, allocator warm-up)
                return cascade_batch(self.nets, self.optims, self.samples, o, d, gt, global_batch, self.seed, offset, 1)
            if self._rec is None:
                self._rec = torch.zeros(ops.STEP_RECORD_BYTES, dtype=torch.uint8, device=o.device)
            self._in = tuple(torch.empty_like(x) for x in (o, d, gt))
            steps = [op.adam.t for op in self.optims]
            torch.cuda.synchronize()
            c0 = lib.launch_count()
            graph = torch.cuda.CUDAGraph()
            ops.set_step_record(self._rec)
            try:
                with torch.cuda.graph(graph):
                    out = cascade_batch(self.nets, self.optims, self.samples, *self._in, global_batch, self.seed, 0, 1)
            finally:
                ops.set_step_record(None, o.device)
            self._launches = lib.launch_count() - c0
            lib.load().flnerf_launch_count_add(-self._launches)              # captured, not launched
            for op, t in zip(self.optims, steps):
                op.adam.t = t
            self._graph, self._key, self._out = graph, key, out
        for dst, src in zip(self._in, (o, d, gt)):
            dst.copy_(src, non_blocking=True)
        ops.step_record_write(self._rec, 0, offset, a.lr, a.betas[0], a.betas[1], a.t + 1)
        self._graph.replay()
        lib.load().flnerf_launch_count_add(self._launches)
        for op in self.optims:
            op.adam.t += 1
        return self._out


@torch.no_grad()
def train_step(models, rays_o, rays_d, target_rgb, args, tree_mgr=None, seed=0):
    """ddp_train_nerf.py:327-424: one pass over the epoch's rays in batches of args.batch_size; returns the last cascade
    level's predictions in emission order.  Under torchrun every rank takes rows r::world of each batch and the level's
    gradients are all-reduced once per level.  tree_mgr (optional) accumulates the quadtree's refinement statistic per batch
    on the device instead of collecting predictions."""
    world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
    rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
    L = models["cascade_level"]
    nets = [models["net_%d" % m] for m in range(L)]
    optims = [models["optim_%d" % m] for m in range(L)]
    samples = models["cascade_samples"]
    if L != 2:
        raise FlnerfError("train_step implements cascade_level = 2 (coarse + fine), like every config of the fork")
    epoch_size, bs = rays_o.shape[0], args.batch_size
    step = getattr(nets[0], "_cascade_step", None)       # kept on the first level's module: `models` stays the reference's dict
    if step is None or step.nets != nets or step.optims != optims:
        step = nets[0]._cascade_step = CascadeStep(nets, optims, samples, seed, world, graph=not getattr(args, "no_graph", False))
    step.seed = seed
    preds, it, offset = [], 0, 0
    for b0 in range(0, epoch_size, bs):
        b1 = min(b0 + bs, epoch_size)
        sl = slice(b0 + rank, b1, world)
        o, d, gt = rays_o[sl].contiguous(), rays_d[sl].contiguous(), target_rgb[sl].contiguous()
        B = o.shape[0]
        losses, ret = step(o, d, gt, b1 - b0, offset)
        offset += B * (samples[0] + samples[1])
        if tree_mgr is not None:
            tree_mgr.accumulate(ret["rgb"], gt, tree_mgr.ray_gid[sl].contiguous())
        else:
            preds.append(ret["rgb"].clone())          # a replayed step returns the graph's static output
        if it % 400 == 0:
            l = torch.cat(losses)
            if world > 1:
                torch.distributed.all_reduce(l)
            l = l.tolist()
            if rank == 0:
                print("{}//iter {}: level1/loss {:.4f}, level1/psnr {:.4f}, level2/loss {:.4f}, level2/psnr {:.4f}".format(
                    time.strftime("%Y-%m-%d %H:%M:%S", time.localtime()), it, l[0], -10 * np.log10(max(l[0], 1e-12)), l[1],
                    -10 * np.log10(max(l[1], 1e-12))))
        it += 1
    print("{}//total: {} iters.".format(time.strftime("%Y-%m-%d %H:%M:%S", time.localtime()), it))
    return torch.cat(preds, 0) if preds else None


def synthetic_truck(H=270, W=480, n_views=8, device="cuda"):
    """SURVEY 8d configs[4]: 960x540 frames halved by the fork's hard-coded resolution_level=2 (data_loader_split.py:102),
    cameras INSIDE the unit sphere (radius 0.6, looking at the origin).  The images are those of the analytic blob scene seen
    from radius 4: a uniform scaling of the world by 0.15 leaves every perspective image unchanged."""
    K = synthetic.intrinsics(H, W, 0.6 * W)
    far_poses = synthetic.lego_like_poses(n_views, phi=-20.0, radius=4.0)
    imgs = synthetic.render_scene(H, W, K, far_poses, device=device)
    poses = far_poses.copy()
    poses[:, :3, 3] *= 0.15
    return H, W, K, imgs, poses


def ddp_train_nerf(args):
    """ddp_train_nerf.py:184-324 on the synthetic scene."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl")
    if args.dataset_type != "synthetic":
        raise FlnerfError("ddp_train_nerf: only dataset_type = synthetic is built in (the fork's file loaders are out of scope)")
    os.makedirs(os.path.join(args.basedir, args.expname), exist_ok=True)
    H, W, K, imgs, poses = synthetic_truck(int(os.environ.get("FLNERF_PP_H", 270)), int(os.environ.get("FLNERF_PP_W", 480)),
                                           int(os.environ.get("FLNERF_PP_VIEWS", 8)), device=torch.device("cuda", local))
    start, models = create_nerf(local, args)
    max_level = args.init_level + sum(1 for i in range(1, args.n_epoch + 1) if i % args.subdivide_every == 0 and i < args.n_epoch - 1)
    mgr = QuadTreeManager(H, W, K, imgs, torch.as_tensor(poses[:, :3, :4]), mseThres=0.0, max_depth=args.init_level,
                          max_level=max_level, device=torch.device("cuda", local), use_mean=True)
    for epoch_id in range(start + 1, args.n_epoch + 1):
        t0 = time.time()
        print("*" * 46 + "\nEpoch " + str(epoch_id) + "\n" + "*" * 46)
        last = epoch_id == args.n_epoch
        if last:
            print("last epoch: use all rays to train.")
        n = mgr.emit_epoch(down_scale=args.rays_downscale, last_epoch=last, prob=not last, randSamp_proc=args.randSamp_perc)
        o, d, rgb, _ = mgr.batch(0, n, 1)
        print("training rays num: " + str(n))
        mgr.reset_leaf_stats()
        train_step(models, o, d, rgb, args, tree_mgr=mgr, seed=epoch_id * 7919)
        if epoch_id % args.subdivide_every == 0 and epoch_id < args.n_epoch - 1:
            if world > 1:
                torch.distributed.all_reduce(mgr.leaf_sum)
                torch.distributed.all_reduce(mgr.leaf_cnt)
            mgr.refine(args.subdivide_thres)
            print("After sudivide, there are {} child nodes".format(int(mgr.counts.sum().item())))
        if rank == 0:
            to_save = OrderedDict()
            for m in range(models["cascade_level"]):
                for name in ("net_%d" % m, "optim_%d" % m):
                    to_save[name] = models[name].state_dict()
            torch.save(to_save, os.path.join(args.basedir, args.expname, "model_{:04d}.pth".format(epoch_id)))
        print("one step finished. cost time: {}s.".format(int(time.time() - t0)))
    return models


if __name__ == "__main__":
    ddp_train_nerf(config_parser().parse_args())
