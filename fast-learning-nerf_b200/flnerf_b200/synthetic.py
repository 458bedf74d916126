"""Synthetic stand-ins for the datasets the image does not ship (no network): blender-style cameras on a
sphere (load_blender.py:10-34 conventions) looking at an analytic scene of coloured Gaussian blobs, rendered
with the library's own compositing kernel so that a NeRF can fit it and PSNR is meaningful."""
import math

import numpy as np
import torch

from . import ops


def pose_spherical(theta, phi, radius):
    """Camera-to-world of a camera at (theta, phi, radius) looking at the origin (blender convention)."""
    t = np.eye(4); t[2, 3] = radius
    ph = phi / 180.0 * np.pi
    rp = np.array([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0], [0, np.sin(ph), np.cos(ph), 0], [0, 0, 0, 1.0]])
    th = theta / 180.0 * np.pi
    rt = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1.0]])
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1.0]])
    return (flip @ rt @ rp @ t).astype(np.float32)


def intrinsics(H, W, focal):
    return np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], dtype=np.float64)


def lego_like_poses(n, phi=-30.0, radius=4.0):
    return np.stack([pose_spherical(a, phi, radius) for a in np.linspace(-180, 180, n + 1)[:-1]], 0)


_BLOBS = np.array([  # centre xyz, sigma, colour rgb, density
    [0.0, 0.0, 0.0, 0.55, 0.9, 0.7, 0.1, 25.0],
    [0.7, 0.3, 0.2, 0.30, 0.1, 0.6, 0.9, 40.0],
    [-0.6, -0.4, 0.3, 0.35, 0.8, 0.1, 0.2, 35.0],
    [0.1, 0.8, -0.5, 0.25, 0.2, 0.9, 0.3, 50.0],
    [-0.3, 0.5, 0.7, 0.20, 0.9, 0.9, 0.9, 60.0],
], dtype=np.float32)


@torch.no_grad()
def render_scene(H, W, K, poses, near=2.0, far=6.0, n_samples=96, device="cuda", white_bkgd=True, rows_per_chunk=200):
    """images [n,H,W,3] fp32 on ``device`` of the analytic blob scene."""
    blobs = torch.tensor(_BLOBS, device=device)
    out = torch.empty(len(poses), H, W, 3, dtype=torch.float32, device=device)
    for i, c2w in enumerate(poses):
        o, d = ops.raygen(H, W, K, torch.as_tensor(c2w[:3, :4], dtype=torch.float32, device=device))
        for r0 in range(0, H, rows_per_chunk):
            oo, dd = o[r0:r0 + rows_per_chunk].reshape(-1, 3), d[r0:r0 + rows_per_chunk].reshape(-1, 3)
            r11 = ops.pack_rays(oo, dd, near, far, False, H, W, float(K[0][0]))
            z = ops.coarse_depths(r11, n_samples, False, False)
            pts = oo[:, None, :] + dd[:, None, :] * z[:, :, None]
            d2 = ((pts[:, :, None, :] - blobs[None, None, :, 0:3]) ** 2).sum(-1)
            dens = blobs[:, 7] * torch.exp(-0.5 * d2 / blobs[:, 3] ** 2)              # [B,S,nb]
            sigma = dens.sum(-1)
            col = (dens[..., None] * blobs[:, 4:7]).sum(-2) / sigma.clamp(min=1e-8)[..., None]
            col = col.clamp(1e-4, 1 - 1e-4)
            raw = torch.cat([torch.log(col / (1 - col)), sigma[..., None]], -1).contiguous()
            rgb = ops.composite_forward(raw, z, dd.contiguous(), None, white_bkgd, want_weights=False)[0]
            out[i, r0:r0 + rows_per_chunk] = rgb.reshape(-1, W, 3)
    return out.clamp_(0, 1)
