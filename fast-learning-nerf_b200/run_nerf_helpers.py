"""Drop-in for nerf-ours/run_nerf_helpers.py: same names and argument meaning, arithmetic in
libflnerf.so (csrc/rays.cu, csrc/composite.cu).  Inputs living on the CPU are moved to the current
CUDA device for the kernel and the result is returned on the caller's device.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from flnerf_b200 import ops
from flnerf_b200.lib import FlnerfError


def _cuda_device():
    if not torch.cuda.is_available():
        raise FlnerfError("flnerf needs a CUDA (sm_100a) device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _to_cuda(t):
    return t if t.is_cuda else t.to(_cuda_device())


# --- misc (run_nerf_helpers.py:9-11): scalar metrics, not on the hot path
def img2mse(x, y):
    return ((x - y) ** 2).mean()


def mse2psnr(x):
    return -10.0 * torch.log(x) / math.log(10.0) * torch.ones(1, device=x.device) if x.dim() == 0 else \
        -10.0 * torch.log(x) / math.log(10.0)


def to8b(x):
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


# --- positional encoding (run_nerf_helpers.py:15-63)
class Embedder:
    """Same kwargs as the reference; the sin/cos bank with log-sampled power-of-two bands on 3-vectors is what
    the CUDA kernel implements (every config of the reference uses exactly that)."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        d = kwargs["input_dims"]
        self.n_freq = kwargs["num_freqs"]
        ok = (d == 3 and kwargs["include_input"] and kwargs["log_sampling"] and kwargs["max_freq_log2"] == self.n_freq - 1
              and list(kwargs["periodic_fns"]) == [torch.sin, torch.cos])
        if not ok:
            raise FlnerfError("flnerf Embedder supports input_dims=3, include_input, log_sampling, [sin, cos]")
        self.out_dim = d + 2 * d * self.n_freq

    def embed(self, inputs):
        dev = inputs.device
        return ops.posenc(_to_cuda(inputs), self.n_freq).to(dev)


def get_embedder(multires, i=0):
    if i == -1:
        return torch.nn.Identity(), 3
    eo = Embedder(include_input=True, input_dims=3, max_freq_log2=multires - 1, num_freqs=multires, log_sampling=True,
                  periodic_fns=[torch.sin, torch.cos])
    fn = lambda x, eo=eo: eo.embed(x)
    fn.n_freq = multires        # lets render.run_network recognise the fused (sample point + PE + MLP) path
    return fn, eo.out_dim


# --- rays (run_nerf_helpers.py:68-108)
def get_rays(H, W, K, c2w):
    c2w = torch.as_tensor(c2w)
    dev = c2w.device
    o, d = ops.raygen(int(H), int(W), K, _to_cuda(c2w.float()))
    return o.to(dev), d.to(dev)


def get_rays_np(H, W, K, c2w):
    o, d = ops.raygen(int(H), int(W), K, torch.as_tensor(np.asarray(c2w), dtype=torch.float32, device=_cuda_device()))
    return o.cpu().numpy(), d.cpu().numpy()


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    if float(near) != 1.0:
        raise FlnerfError("flnerf ndc_rays implements near=1. (the only value the reference passes, render.py:71)")
    dev, shp = rays_o.device, rays_o.shape
    r11 = ops.pack_rays(_to_cuda(rays_o), _to_cuda(rays_d), 0.0, 1.0, True, H, W, float(focal))
    return r11[:, 0:3].reshape(shp).to(dev), r11[:, 3:6].reshape(shp).to(dev)


# --- hierarchical sampling (run_nerf_helpers.py:112-155)
def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    """bins [B, M+1], weights [B, M] -> [B, N_samples] (inverse-CDF sampling).  With
    pytest=True the uniforms come from numpy seed 0 exactly like the reference hook (helpers:127-135)."""
    dev = bins.device
    bins_c, w_c = _to_cuda(bins).float().contiguous(), _to_cuda(weights).float().contiguous()
    B, nb = bins_c.shape
    u = None
    if pytest:
        np.random.seed(0)
        shp = [B, N_samples]
        u = torch.tensor(np.broadcast_to(np.linspace(0., 1., N_samples), shp).copy() if det else np.random.rand(*shp),
                         dtype=torch.float32, device=bins_c.device)
    return ops.sample_pdf_bins(bins_c, w_c, N_samples, det and not pytest, u).to(dev)


# --- SSIM (run_nerf_helpers.py:158-234): evaluation metric, separable Gaussian window as in tf.image.ssim
def compute_ssim(img0, img1, max_val=1.0, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03, return_map=False):
    if (img0.is_cuda and img0.dim() == 3 and img0.shape[-1] == 3 and not return_map
            and (filter_size, filter_sigma, k1, k2) == (11, 1.5, 0.01, 0.03)):
        return ops.ssim_psnr(img0, img1, max_val)[0:1].float()       # csrc/eval.cu: one kernel, no conv library
    # general arguments / CPU tensors / the per-pixel map: the reference's formula on torch ops
    w_, h_, c_ = img0.shape[-3:]
    a = img0.reshape(-1, w_, h_, c_).permute(0, 3, 1, 2)
    b = img1.reshape(-1, w_, h_, c_).permute(0, 3, 1, 2)
    half = filter_size // 2
    shift = (2 * half - filter_size + 1) / 2
    taps = torch.exp(-0.5 * ((torch.arange(filter_size, device=a.device) - half + shift) / filter_sigma) ** 2)
    taps = taps / taps.sum()
    kv = taps.view(1, 1, -1, 1).repeat(c_, 1, 1, 1)
    kh = taps.view(1, 1, 1, -1).repeat(c_, 1, 1, 1)

    def blur(t):
        return F.conv2d(F.conv2d(t, kh, padding=[0, half], groups=c_), kv, padding=[half, 0], groups=c_)

    m0, m1 = blur(a), blur(b)
    v0 = (blur(a * a) - m0 * m0).clamp(min=0.0)
    v1 = (blur(b * b) - m1 * m1).clamp(min=0.0)
    cov = blur(a * b) - m0 * m1
    cov = torch.sign(cov) * torch.minimum(torch.sqrt(v0 * v1), cov.abs())
    c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
    smap = ((2 * m0 * m1 + c1) * (2 * cov + c2)) / ((m0 * m0 + m1 * m1 + c1) * (v0 + v1 + c2))
    return smap if return_map else smap.reshape(-1, c_ * w_ * h_).mean(-1)
