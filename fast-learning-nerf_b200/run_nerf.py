"""Drop-in for nerf-ours/run_nerf.py: same CLI (``python run_nerf.py --config configs/lego.txt``), same
``create_nerf`` / ``run_network`` / ``batchify`` / ``train`` entry points, same checkpoint ({epoch:03d}.tar with
``module.``-prefixed state dicts + stock Adam state) and tree pickle (treeDivide_{epoch:04d}.pkl) formats.

What changed underneath: rays, targets and quadtrees stay on the GPU; the batch loop is
``flnerf_b200.engine.Trainer.step_from_tree`` (no per-iteration D2H copies: the |gt-pred| statistic
adjust_tree needs is reduced on the device); multi-GPU is one process per GPU (torchrun) with a single
all-reduce of the flat gradient bucket instead of nn.DataParallel.
"""
import os
import pickle
import sys
import time

import numpy as np
import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

from run_nerf_helpers import *  # noqa: F401,F403,E402
from run_nerf_helpers import get_embedder, img2mse, mse2psnr, to8b  # noqa: E402
from argument_parser import config_parser  # noqa: E402
from model import NeRF  # noqa: E402
from render import render_rays, render_path, render  # noqa: F401,E402
from tree import QuadTreeManager, get_children  # noqa: F401,E402
from flnerf_b200.engine import FusedAdam, Trainer, lr_at  # noqa: E402
from flnerf_b200 import synthetic  # noqa: E402

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


class ModuleHolder(nn.Module):
    """Keeps the ``module.`` prefix nn.DataParallel put on checkpoint keys (run_nerf.py:82,90) without any of its
    scatter/replicate/gather machinery."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


def _unwrap(net):
    return net.module if hasattr(net, "module") else net


def batchify(fn, chunk):
    """run_nerf.py:40-47"""
    if chunk is None:
        return fn
    return lambda inputs: torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)


def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """run_nerf.py:50-64 for explicit sample points: PE kernels + MLP in netchunk slices."""
    flat = inputs.reshape(-1, inputs.shape[-1])
    emb = embed_fn(flat)
    if viewdirs is not None:
        dirs = viewdirs[:, None].expand(inputs.shape).reshape(-1, inputs.shape[-1])
        emb = torch.cat([emb, embeddirs_fn(dirs)], -1)
    out = batchify(fn, netchunk)(emb)
    return out.reshape(list(inputs.shape[:-1]) + [out.shape[-1]])


class NetworkQuery:
    """``network_query_fn(inputs, viewdirs, network_fn)`` of the reference, plus the fused entry render_rays uses
    (sample points are never materialised: one kernel goes from rays + depths to the MLP's input tiles)."""

    def __init__(self, embed_fn, embeddirs_fn, netchunk):
        self.embed_fn, self.embeddirs_fn, self.netchunk = embed_fn, embeddirs_fn, netchunk

    def __call__(self, inputs, viewdirs, network_fn):
        return run_network(inputs, viewdirs, network_fn, self.embed_fn, self.embeddirs_fn, self.netchunk)

    def fused_rays(self, rays11, z, network_fn):
        return _unwrap(network_fn).query_rays(rays11, z)


def create_nerf(args):
    """run_nerf.py:67-153; returns the same 6-tuple."""
    embed_fn, input_ch = get_embedder(args.multires, args.i_embed)
    input_ch_views, embeddirs_fn = 0, None
    if args.use_viewdirs:
        embeddirs_fn, input_ch_views = get_embedder(args.multires_views, args.i_embed)
    output_ch = 5 if args.N_importance > 0 else 4
    precision = getattr(args, "precision", None)

    def make(depth, width):
        return ModuleHolder(NeRF(D=depth, W=width, input_ch=input_ch, output_ch=output_ch, skips=[4],
                                 input_ch_views=input_ch_views, use_viewdirs=args.use_viewdirs,
                                 precision=precision).to(device))

    model = make(args.netdepth, args.netwidth)
    grad_vars = list(model.parameters())
    model_fine = None
    if args.N_importance > 0:
        model_fine = make(args.netdepth_fine, args.netwidth_fine)
        grad_vars += list(model_fine.parameters())
    network_query_fn = NetworkQuery(embed_fn, embeddirs_fn, args.netchunk)
    # the nets whose flat buffers the kernels and the fused optimiser work on (use_viewdirs=False: the inner net behind the
    # reference-shaped module, model.NeRF.kernel_net; its Adam state is then NOT the reference's 20-tensor layout)
    nets = [_unwrap(m).kernel_net for m in (model, model_fine) if m is not None]
    optimizer = FusedAdam(grad_vars if args.use_viewdirs else [p for n in nets for p in n.parameters()], nets,
                          lr=args.lrate, betas=(0.9, 0.999))

    start_epoch, start_iter = 0, 0
    basedir, expname = args.basedir, args.expname
    if args.ft_path is not None and args.ft_path != 'None':
        ckpts = [args.ft_path]
    else:
        d = os.path.join(basedir, expname)
        ckpts = [os.path.join(d, f) for f in sorted(os.listdir(d)) if 'tar' in f] if os.path.isdir(d) else []
    print('Found ckpts', ckpts)
    if len(ckpts) > 0 and not args.no_reload:
        print('Reloading from', ckpts[-1])
        ckpt = torch.load(ckpts[-1], map_location=device, weights_only=False)
        start_epoch, start_iter = ckpt['global_epoch'], ckpt['global_iter']
        optimizer.load_state_dict(ckpt['optimizer_state_dict'])
        model.load_state_dict(ckpt['network_fn_state_dict'])
        if model_fine is not None:
            model_fine.load_state_dict(ckpt['network_fine_state_dict'])

    render_kwargs_train = {
        'network_query_fn': network_query_fn, 'perturb': args.perturb, 'N_importance': args.N_importance,
        'network_fine': model_fine, 'N_samples': args.N_samples, 'network_fn': model,
        'use_viewdirs': args.use_viewdirs, 'white_bkgd': args.white_bkgd, 'raw_noise_std': args.raw_noise_std,
    }
    if args.dataset_type != 'llff' or args.no_ndc:
        print('Not ndc!')
        render_kwargs_train['ndc'] = False
        render_kwargs_train['lindisp'] = args.lindisp
    render_kwargs_test = dict(render_kwargs_train)
    render_kwargs_test['perturb'] = False
    render_kwargs_test['raw_noise_std'] = 0.
    return render_kwargs_train, render_kwargs_test, start_epoch, start_iter, grad_vars, optimizer


def load_dataset(args):
    """Returns images[N,H,W,3 or 4] (numpy), poses[N,3or4,4], render_poses, hwf, i_split, near, far.
    ``dataset_type = synthetic`` needs no files; the real loaders are the reference's own load_*.py (host-side
    I/O, out of scope here) found on $FLNERF_LOADERS (a directory holding load_blender.py / load_llff.py)."""
    if args.dataset_type == 'synthetic':
        H = W = int(os.environ.get("FLNERF_SYN_RES", 400 if args.half_res else 800))
        focal = 0.5 * W / np.tan(0.5 * 0.6911112070083618)
        n_train, n_test = int(os.environ.get("FLNERF_SYN_VIEWS", 100)), 8
        poses = np.concatenate([synthetic.lego_like_poses(n_train), synthetic.lego_like_poses(n_test, phi=-20.0)], 0)
        K = synthetic.intrinsics(H, W, focal)
        images = synthetic.render_scene(H, W, K, poses, device=device).cpu().numpy()
        i_split = [np.arange(n_train), np.arange(n_train, n_train + n_test), np.arange(n_train, n_train + n_test)]
        return images, poses, torch.as_tensor(poses[n_train:]), [H, W, focal], i_split, 2., 6.
    loaders = os.environ.get("FLNERF_LOADERS")
    if not loaders:
        raise RuntimeError("dataset_type=%s needs the reference loaders: set FLNERF_LOADERS to the directory that "
                           "holds load_blender.py / load_llff.py (e.g. <reference>/nerf-ours)" % args.dataset_type)
    sys.path.append(loaders)
    if args.dataset_type == 'blender':
        from load_blender import load_blender_data
        images, poses, render_poses, hwf, i_split = load_blender_data(args.datadir, args.half_res, args.testskip)
        if args.white_bkgd:
            images = images[..., :3] * images[..., -1:] + (1. - images[..., -1:])       # run_nerf.py:199-202
        else:
            images = images[..., :3]
        return images, poses, render_poses, hwf, i_split, 2., 6.
    if args.dataset_type == 'llff':
        from load_llff import load_llff_data
        images, poses, bds, render_poses, i_test = load_llff_data(args.datadir, args.factor, recenter=True,
                                                                  bd_factor=.75, spherify=args.spherify)
        hwf = poses[0, :3, -1]
        poses = poses[:, :3, :4]
        if not isinstance(i_test, list):
            i_test = [i_test]
        if args.llffhold > 0:
            i_test = np.arange(images.shape[0])[::args.llffhold]
        i_val = i_test
        i_train = np.array([i for i in np.arange(int(images.shape[0])) if (i not in i_test and i not in i_val)])
        near, far = (np.ndarray.min(bds) * .9, np.ndarray.max(bds) * 1.) if args.no_ndc else (0., 1.)
        return images, poses, render_poses, hwf, [i_train, i_val, i_test], near, far
    raise RuntimeError('Unknown dataset type %s' % args.dataset_type)


def train(argv=None):
    args = config_parser().parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not torch.distributed.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    images, poses, render_poses, hwf, i_split, near, far = load_dataset(args)
    i_train, i_val, i_test = i_split
    H, W, focal = int(hwf[0]), int(hwf[1]), float(hwf[2])
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])                   # run_nerf.py:237-242
    if args.render_test:
        render_poses = np.array(poses[i_test])
    basedir, expname = args.basedir, args.expname
    if rank == 0:
        os.makedirs(os.path.join(basedir, expname), exist_ok=True)
        with open(os.path.join(basedir, expname, 'args.txt'), 'w') as f:
            for arg in sorted(vars(args)):
                f.write('{} = {}\n'.format(arg, getattr(args, arg)))
        if args.config is not None:
            with open(os.path.join(basedir, expname, 'config.txt'), 'w') as f:
                f.write(open(args.config, 'r').read())

    render_kwargs_train, render_kwargs_test, global_epoch, global_iter, grad_vars, optimizer = create_nerf(args)
    bds = {'near': near, 'far': far}
    render_kwargs_train.update(bds)
    render_kwargs_test.update(bds)
    render_poses = torch.as_tensor(np.asarray(render_poses), dtype=torch.float32).to(device)

    if args.render_only:
        print('RENDER ONLY')
        gt = images[i_test] if args.render_test else None
        savedir = os.path.join(basedir, expname, 'renderonly_{}_{:06d}'.format('test' if args.render_test else 'path', global_iter))
        os.makedirs(savedir, exist_ok=True)
        rgbs, _ = render_path(render_poses, hwf, K, args.chunk, render_kwargs_test, gt_imgs=gt, savedir=savedir,
                              render_factor=args.render_factor)
        print('Done rendering', savedir)
        return

    N_rand = args.N_rand
    train_images = torch.as_tensor(np.asarray(images)[i_train][..., :3], dtype=torch.float32)
    train_poses = torch.as_tensor(np.asarray(poses)[i_train], dtype=torch.float32)
    print('Begin')
    print('TRAIN views are', i_train)
    print('TEST views are', i_test)
    print('VAL views are', i_val)

    # deepest level the schedule can reach (run_nerf.py:347-355) sizes the leaf capacity
    max_level = args.init_level + sum(1 for i in range(1, args.n_epoch + 1)
                                      if i % args.subdivide_every == 0 and i < args.n_epoch - 1)
    treeManager = QuadTreeManager(H, W, K, train_images, train_poses, mseThres=0.0, max_depth=args.init_level,
                                  max_level=max_level, device=device)
    tree_pkl = os.path.join(basedir, expname, 'treeDivide_{:04d}.pkl'.format(global_epoch))
    if os.path.exists(tree_pkl):
        with open(tree_pkl, 'rb') as f:
            treeManager.quadTrees = pickle.load(f)
            treeManager.cur_level = global_epoch
            print("load '" + tree_pkl + "'")

    nc, nf = _unwrap(render_kwargs_train['network_fn']).kernel_net, _unwrap(render_kwargs_train['network_fine']).kernel_net
    trainer = Trainer(nc, nf, optimizer, H, W, K, near, far, args.N_samples, args.N_importance, args.white_bkgd,
                      args.perturb, args.lindisp, render_kwargs_train.get('ndc', True), args.raw_noise_std,
                      world_size=world, rank=rank, graph=not getattr(args, "no_graph", False))

    def log(tag, it, loss):
        l = trainer.global_loss(loss).tolist()   # the only host sync (+ an 8-byte all-reduce), every 50/400 iterations
        if rank != 0:
            return
        print('{}//iter {}: coarse/loss {:.4f}, coarse/psnr {:.4f}, fine/loss {:.4f}, fine/psnr {:.4f}'.format(
            time.strftime("%Y-%m-%d %H:%M:%S", time.localtime()), it, l[0], -10 * np.log10(max(l[0], 1e-12)), l[1],
            -10 * np.log10(max(l[1], 1e-12))))

    if global_epoch == 0:
        # centre-crop warm-up (run_nerf.py:367-423): the same central pixels for every image, image by image
        print('Center Cropping for 500 iters...')
        t0 = time.time()
        dH, dW = H // 4, W // 4
        rows = torch.arange(H // 2 - dH, H // 2 + dH)
        cols = torch.arange(W // 2 - dW, W // 2 + dW)
        coords = torch.stack(torch.meshgrid(rows, cols, indexing='ij'), -1).reshape(-1, 2)
        randNum = int(N_rand * 500 / treeManager.n_images)
        sel = coords[np.random.choice(coords.shape[0], size=[min(randNum, coords.shape[0])], replace=False)]
        pix = (sel[:, 0] * W + sel[:, 1]).to(torch.int32)
        if world > 1:       # every rank must hold the SAME index buffer (rank r consumes rows first + r, first + r + world, ...)
            pix = pix.to(device)
            torch.distributed.broadcast(pix, 0)
            pix = pix.cpu()
        n_img = treeManager.n_images
        ray_pix = pix.repeat(n_img).to(device)
        ray_gid = (torch.arange(n_img, dtype=torch.int32).repeat_interleave(pix.shape[0]) * treeManager.cap).to(device)
        treeManager.ray_pix, treeManager.ray_gid, treeManager.n_rays = ray_pix, ray_gid, int(ray_pix.shape[0])
        it = 0
        for first in range(0, treeManager.n_rays, N_rand):
            rows_ = min(N_rand, treeManager.n_rays - first)
            local = (rows_ - rank + world - 1) // world
            o, d, tgt, _ = treeManager.batch(first + rank, local, world)
            loss = trainer.step(o, d, tgt, None, None, global_batch=rows_)      # lr is not decayed here (quirk 5)
            if it % 50 == 0:
                log('crop', it, loss)
            it += 1
        torch.cuda.synchronize()
        print('pre Center Cropping finished. cost time: {}s.'.format(time.time() - t0))

    for epoch_id in range(global_epoch + 1, args.n_epoch + 1):
        print('*' * 46 + '\nEpoch ' + str(epoch_id) + '\n' + '*' * 46)
        t_epoch = time.time()
        print('generating rays... cur level=' + str(treeManager.cur_level))
        last = epoch_id == args.n_epoch
        if last:
            print('last epoch: use all rays to train.')
        n_rays = treeManager.emit_epoch(down_scale=1, last_epoch=last)
        torch.cuda.synchronize()
        print('shuffling rays costs {:.2f}s.'.format(time.time() - t_epoch))
        print('training rays num: ' + str(n_rays))
        treeManager.reset_leaf_stats()
        it = 0
        for first in range(0, n_rays, N_rand):
            loss = trainer.step_from_tree(treeManager, first, N_rand)
            new_lrate = lr_at(args.lrate, args.lrate_decay, global_iter)        # applied after the step (:498-502)
            for g in optimizer.param_groups:
                g['lr'] = new_lrate
            global_iter += 1
            if it % 400 == 0:
                log('train', it, loss)
            it += 1
        print('{}//total: {} iters.'.format(time.strftime("%Y-%m-%d %H:%M:%S", time.localtime()), it))
        if args.subdivide_every > 0 and epoch_id % args.subdivide_every == 0 and epoch_id < args.n_epoch - 1:
            t1 = time.time()
            if world > 1:
                torch.distributed.all_reduce(treeManager.leaf_max, op=torch.distributed.ReduceOp.MAX)
            treeManager.refine(args.subdivide_thres)
            torch.cuda.synchronize()
            print('After sudivide, there are {} child nodes'.format(int(treeManager.counts.sum().item())))
            print('adjust quadTree cost {:.2f}s.'.format(time.time() - t1))
        if rank == 0:
            path = os.path.join(basedir, expname, '{:03d}.tar'.format(epoch_id))
            torch.save({
                'global_epoch': epoch_id,
                'global_iter': global_iter,
                'network_fn_state_dict': render_kwargs_train['network_fn'].state_dict(),
                'network_fine_state_dict': render_kwargs_train['network_fine'].state_dict(),
                'optimizer_state_dict': optimizer.state_dict(),
            }, path)
            print('Saved checkpoints at', path)
            with open(os.path.join(basedir, expname, 'treeDivide_{:04d}.pkl'.format(epoch_id)), 'wb') as f:
                pickle.dump(treeManager.quadTrees, f)
        print('one step finished. cost time: {}s.'.format(int(time.time() - t_epoch)))


if __name__ == '__main__':
    t0 = time.time()
    train()
    print('train complete. time={:.1f}s.'.format(time.time() - t0))
