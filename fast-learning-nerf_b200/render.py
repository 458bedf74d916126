"""Drop-in for nerf-ours/render.py: ``render``, ``batchify_rays``, ``render_rays``, ``raw2outputs``,
``render_path`` with the reference signatures and return conventions.  Stage arithmetic runs in
libflnerf.so: ray packing / NDC (csrc/rays.cu), stratified depths + fused sample-point/PE encoding,
the MLP (csrc/mlp_tc.cu | mlp_simt.cu), warp-scan compositing and inverse-CDF resampling + merge
(csrc/composite.cu).  Tensors stay on the GPU from the ray batch to rgb_map.
"""
import os
import time

import numpy as np
import torch

from run_nerf_helpers import get_rays, to8b, compute_ssim  # noqa: F401  (same star-import surface as the reference)
from run_nerf_helpers import *  # noqa: F401,F403
from flnerf_b200 import ops
from flnerf_b200.lib import FlnerfError

_SEED = {"base": int(os.environ.get("FLNERF_SEED", "0")), "calls": 0}


def _next_offset(n):
    """Philox counter space for the in-kernel uniforms (the reference draws torch.rand, render.py:258)."""
    off = _SEED["calls"]
    _SEED["calls"] += int(n)
    return off


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """Render rays in chunks (render.py:12-24); chunking only bounds memory."""
    parts = {}
    for i in range(0, rays_flat.shape[0], chunk):
        out = render_rays(rays_flat[i:i + chunk], **kwargs)
        for k, v in out.items():
            parts.setdefault(k, []).append(v)
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in parts.items()}


def render(H, W, K, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, **kwargs):
    """render.py:26-91.  Returns [rgb_map, disp_map, acc_map, extras]."""
    if c2w is not None and c2w_staticcam is None and use_viewdirs and _frame_path_ok(near, far, kwargs):
        return _render_frame(H, W, K, chunk, c2w, ndc, float(near), float(far), kwargs)
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, K, c2w)
    else:
        rays_o, rays_d = rays
    # use_viewdirs=False (render.py:59-66 skipped): the packed rays still carry d/|d|, a model built without view
    # directions multiplies it by zero weights (model.NeRF)
    view_src = rays_d
    if c2w_staticcam is not None:
        rays_o, rays_d = get_rays(H, W, K, c2w_staticcam)
    sh = rays_d.shape
    scalar_bounds = not (torch.is_tensor(near) or torch.is_tensor(far))
    rays11 = ops.pack_rays(rays_o, rays_d, float(near) if scalar_bounds else 0.0, float(far) if scalar_bounds else 1.0,
                           ndc, H, W, float(K[0][0]))
    if not scalar_bounds:
        rays11[:, 6] = torch.as_tensor(near, device=rays11.device, dtype=torch.float32).reshape(-1)
        rays11[:, 7] = torch.as_tensor(far, device=rays11.device, dtype=torch.float32).reshape(-1)
    if c2w_staticcam is not None:   # view directions keep following the moving camera (render.py:61-66)
        rays11[:, 8:11] = ops.pack_rays(rays_o, view_src, 0.0, 1.0, False, H, W, float(K[0][0]))[:, 8:11]
    all_ret = batchify_rays(rays11, chunk, **kwargs)
    for k in all_ret:
        all_ret[k] = all_ret[k].reshape(list(sh[:-1]) + list(all_ret[k].shape[1:]))
    head = ['rgb_map', 'disp_map', 'acc_map']
    return [all_ret[k] for k in head] + [{k: v for k, v in all_ret.items() if k not in head}]


def _frame_path_ok(near, far, kw):
    """Full-frame rendering from a pose with un-jittered depths through a tensor-core network: the fused front end applies."""
    if torch.is_tensor(near) or torch.is_tensor(far) or kw.get("perturb", 0.) > 0. or kw.get("pytest", False):
        return False
    if getattr(kw.get("network_query_fn"), "fused_rays", None) is None:
        return False
    net = kw.get("network_fn")
    net = net.module if hasattr(net, "module") else net
    return getattr(net, "mode", ops.MODE_FP32) != ops.MODE_FP32


def _render_frame(H, W, K, chunk, c2w, ndc, near, far, kw):
    """render(c2w=...) (render.py:52-91) with the coarse pass's inputs produced by ONE kernel per chunk of pixels
    (ops.encode_frame_tc) instead of get_rays over the whole frame + packing + depths + PE."""
    net = kw["network_fn"]
    net = net.module if hasattr(net, "module") else net
    kw = dict(kw)
    N_samples, lindisp = kw["N_samples"], kw.get("lindisp", False)
    parts = {}
    for p0 in range(0, H * W, chunk):
        B = min(chunk, H * W - p0)
        rays11, z, tiles, dirpe = ops.encode_frame_tc(net.mode, H, W, K, torch.as_tensor(c2w), near, far, ndc, lindisp, p0, B,
                                                      N_samples)
        out = render_rays(rays11, _coarse=(z, tiles, dirpe), **kw)
        for k, v in out.items():
            parts.setdefault(k, []).append(v)
    all_ret = {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in parts.items()}
    for k in all_ret:
        all_ret[k] = all_ret[k].reshape([H, W] + list(all_ret[k].shape[1:]))
    head = ['rgb_map', 'disp_map', 'acc_map']
    return [all_ret[k] for k in head] + [{k: v for k, v in all_ret.items() if k not in head}]


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False):
    """render.py:149-192 -> (rgb_map, disp_map, acc_map, weights, depth_map)."""
    noise = None
    if raw_noise_std > 0.:
        if pytest:
            np.random.seed(0)
            noise = torch.tensor(np.random.rand(*list(raw[..., 3].shape)) * raw_noise_std, dtype=torch.float32,
                                 device=raw.device)
        else:
            noise = torch.randn(raw[..., 3].shape, device=raw.device) * raw_noise_std
    rgb, disp, acc, w, depth = ops.CompositeFn.apply(raw, z_vals, rays_d, noise, bool(white_bkgd))
    return rgb, disp, acc, w, depth


def _query(network_query_fn, rays11, z, net):
    fused = getattr(network_query_fn, "fused_rays", None)
    if fused is not None:
        return fused(rays11, z, net)
    # user-supplied query function: hand it explicit sample points like the reference does (render.py:268-272)
    pts = rays11[:, None, 0:3] + rays11[:, None, 3:6] * z[:, :, None]
    return network_query_fn(pts, rays11[:, 8:11], net)


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, retraw=False, lindisp=False, perturb=0.,
                N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., verbose=False, pytest=False,
                _coarse=None):
    """Volumetric rendering of a ray batch (render.py:195-305); same keys in the returned dict."""
    if ray_batch.shape[-1] == 8:          # [o, d, near, far] (use_viewdirs=False, render.py:74-80): no view direction
        ray_batch = torch.cat([ray_batch, ray_batch.new_zeros(ray_batch.shape[0], 3)], -1)
    if ray_batch.shape[-1] < 11:
        raise FlnerfError("flnerf render_rays needs [o,d,near,far(,viewdir)] rays")
    rays11 = ray_batch.float().contiguous()
    B = rays11.shape[0]
    rays_d = rays11[:, 3:6]
    if _coarse is not None:     # eval path: depths + the coarse network's input tiles came with the rays (ops.encode_frame_tc)
        z_vals, tiles, dirpe = _coarse
        net = network_fn.module if hasattr(network_fn, "module") else network_fn
        raw = net.query_tiles(tiles, dirpe, B, N_samples)
    else:
        t_rand = None
        if perturb > 0. and pytest:
            np.random.seed(0)
            t_rand = torch.tensor(np.random.rand(B, N_samples), dtype=torch.float32, device=rays11.device)
        z_vals = ops.coarse_depths(rays11, N_samples, perturb > 0., lindisp, t_rand, _SEED["base"],
                                   _next_offset(B * N_samples) if (perturb > 0. and t_rand is None) else 0)
        raw = _query(network_query_fn, rays11, z_vals, network_fn)
    rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkgd, pytest)
    if N_importance > 0:
        rgb0, disp0, acc0 = rgb_map, disp_map, acc_map
        det = (perturb == 0.)
        u = None
        if pytest:
            np.random.seed(0)
            u = torch.tensor(np.broadcast_to(np.linspace(0., 1., N_importance), (B, N_importance)).copy() if det
                             else np.random.rand(B, N_importance), dtype=torch.float32, device=rays11.device)
        z_vals, z_samples, z_std = ops.sample_pdf_merge(
            z_vals, weights.detach(), N_importance, det and u is None, u, _SEED["base"] + 1,
            _next_offset(B * N_importance) if (not det and u is None) else 0)
        run_fn = network_fn if network_fine is None else network_fine
        raw = _query(network_query_fn, rays11, z_vals, run_fn)
        rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkgd,
                                                                     pytest)
    ret = {'rgb_map': rgb_map, 'disp_map': disp_map, 'acc_map': acc_map}
    if retraw:
        ret['raw'] = raw
    if N_importance > 0:
        ret['rgb0'], ret['disp0'], ret['acc0'], ret['z_std'] = rgb0, disp0, acc0, z_std
    return ret


def render_path(render_poses, hwf, K, chunk, render_kwargs, gt_imgs=None, savedir=None, render_factor=0):
    """render.py:94-146: renders every pose, reports PSNR / SSIM (/ LPIPS when the lpips package exists)."""
    H, W, focal = hwf
    if render_factor != 0:
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor
    try:
        import lpips
        lpips_vgg = lpips.LPIPS(net="vgg").eval().cuda()
    except Exception:
        lpips_vgg = None
    rgbs, disps, psnrs, ssims, lps = [], [], [], [], []
    t0 = time.time()
    for i, c2w in enumerate(render_poses):
        with torch.no_grad():
            rgb, disp, acc, _ = render(H, W, K, chunk=chunk, c2w=c2w[:3, :4], **render_kwargs)
        rgbs.append(rgb.cpu().numpy())
        disps.append(disp.cpu().numpy())
        if i == 0:
            print(rgb.shape, disp.shape)
        if gt_imgs is not None and render_factor == 0:
            gt = torch.as_tensor(gt_imgs[i]).float().to(rgb.device)[..., :3]
            ssim, psnr = ops.ssim_psnr(gt, rgb).tolist()         # one kernel on the device (csrc/eval.cu)
            lp = float('nan')
            if lpips_vgg is not None:
                lp = lpips_vgg(gt.permute(2, 0, 1).contiguous(), rgb.permute(2, 0, 1).contiguous(), normalize=True).item()
            print('img-{}: psnr={}, ssim={}, lpips={}'.format(i, psnr, ssim, lp))
            psnrs.append(psnr); ssims.append(ssim); lps.append(lp)
        if savedir is not None:
            _imwrite(os.path.join(savedir, '{:03d}.png'.format(i)), to8b(rgbs[-1]))
    results = 'mean PSNR: {}\nmean SSIM: {}\nmean LPIPS: {}'.format(
        np.mean(psnrs) if psnrs else float('nan'), np.mean(ssims) if ssims else float('nan'),
        np.mean(lps) if lps else float('nan'))
    print(results, '({:.2f}s)'.format(time.time() - t0))
    if savedir is not None:
        with open(os.path.join(savedir, 'results.txt'), 'w') as f:
            f.write(results)
    return np.stack(rgbs, 0), np.stack(disps, 0)


def _imwrite(path, img8):
    try:
        import imageio
        imageio.imwrite(path, img8)
    except (ImportError, AttributeError):      # absent, or a stub module without imwrite
        import cv2
        cv2.imwrite(path, np.ascontiguousarray(img8[..., ::-1]))
