"""TEST INFRASTRUCTURE -- CPU restatement of the reference's nerf++ render path (nerf++-ours/, SURVEY 8f rank 1:
the next row after the nerf-ours hot path; BASELINE.json configs[4]).

Nothing here is shipped or measured: it is the checker the future CUDA path for this row will be diffed against, pinned
against the UNMODIFIED reference (oracle/ref_shim.load_nerfpp) by tests/test_nerfpp_oracle.py and the fixtures in
tests/golden/nerfpp.npz.  Every function cites the reference lines it follows.

Layout note (what makes this row cheap on the GPU side): ``MLPNet`` has the GEMM structure of nerf-ours' ``NeRF`` --
``mlp_params_to_nerf_layout`` maps its state_dict onto the flat parameter order the tensor-core kernels already use
(foreground net: identical shapes; background net: 84 instead of 63 position channels)."""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

TINY_NUMBER = 1e-6      # utils.py:8
HUGE_NUMBER = 1e10      # utils.py:7


# ------------------------------------------------------------------------------------------------ sampling helpers
def intersect_sphere(ray_o: torch.Tensor, ray_d: torch.Tensor) -> torch.Tensor:
    """ddp_train_nerf.py:54-69: depth at which the ray leaves the unit sphere (cameras must lie inside it)."""
    d1 = -torch.sum(ray_d * ray_o, dim=-1) / torch.sum(ray_d * ray_d, dim=-1)
    p = ray_o + d1.unsqueeze(-1) * ray_d
    ray_d_cos = 1.0 / torch.norm(ray_d, dim=-1)
    p_norm_sq = torch.sum(p * p, dim=-1)
    if bool((p_norm_sq >= 1.0).any()):
        raise ValueError("camera outside the unit sphere (ddp_train_nerf.py:65-66)")
    return d1 + torch.sqrt(1.0 - p_norm_sq) * ray_d_cos


def perturb_samples(z_vals: torch.Tensor, t_rand: torch.Tensor) -> torch.Tensor:
    """ddp_train_nerf.py:72-81 with the uniforms made explicit (the reference draws torch.rand_like)."""
    mids = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])
    upper = torch.cat([mids, z_vals[..., -1:]], dim=-1)
    lower = torch.cat([z_vals[..., 0:1], mids], dim=-1)
    return lower + (upper - lower) * t_rand


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, n_samples: int, u: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ddp_train_nerf.py:84-133: count-based inverse CDF (no searchsorted), TINY_NUMBER floor on the weights AND on the
    bin width.  ``u`` None = deterministic linspace(0, 1, n_samples) (det=True); else the uniforms torch.rand would draw."""
    weights = weights + TINY_NUMBER
    pdf = weights / torch.sum(weights, dim=-1, keepdim=True)
    cdf = torch.cumsum(pdf, dim=-1)
    cdf = torch.cat([torch.zeros_like(cdf[..., 0:1]), cdf], dim=-1)
    dots = list(weights.shape[:-1])
    M = weights.shape[-1]
    if u is None:
        u = torch.linspace(0.0, 1.0, n_samples).view([1] * len(dots) + [n_samples]).expand(dots + [n_samples])
    above = torch.sum(u.unsqueeze(-1) >= cdf[..., :M].unsqueeze(-2), dim=-1).long()
    below = torch.clamp(above - 1, min=0)
    idx = torch.stack((below, above), dim=-1)
    cdf_g = torch.gather(cdf.unsqueeze(-2).expand(dots + [n_samples, M + 1]), -1, idx)
    bins_g = torch.gather(bins.unsqueeze(-2).expand(dots + [n_samples, M + 1]), -1, idx)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom = torch.where(denom < TINY_NUMBER, torch.ones_like(denom), denom)
    t = (u - cdf_g[..., 0]) / denom
    return bins_g[..., 0] + t * (bins_g[..., 1] - bins_g[..., 0] + TINY_NUMBER)


def depth2pts_outside(ray_o: torch.Tensor, ray_d: torch.Tensor, depth: torch.Tensor):
    """ddp_model.py:16-45: inverted-sphere parametrisation of a background sample; depth = 1/r in [0,1].  Returns
    ([..., 4] = unit direction of the sample + 1/r, conventional depth)."""
    d1 = -torch.sum(ray_d * ray_o, dim=-1) / torch.sum(ray_d * ray_d, dim=-1)
    p_mid = ray_o + d1.unsqueeze(-1) * ray_d
    p_mid_norm = torch.norm(p_mid, dim=-1)
    ray_d_cos = 1.0 / torch.norm(ray_d, dim=-1)
    d2 = torch.sqrt(1.0 - p_mid_norm * p_mid_norm) * ray_d_cos
    p_sphere = ray_o + (d1 + d2).unsqueeze(-1) * ray_d
    rot_axis = torch.cross(ray_o, p_sphere, dim=-1)
    rot_axis = rot_axis / torch.norm(rot_axis, dim=-1, keepdim=True)
    phi = torch.asin(p_mid_norm)
    theta = torch.asin(p_mid_norm * depth)
    rot_angle = (phi - theta).unsqueeze(-1)
    p_new = p_sphere * torch.cos(rot_angle) + torch.cross(rot_axis, p_sphere, dim=-1) * torch.sin(rot_angle) + \
        rot_axis * torch.sum(rot_axis * p_sphere, dim=-1, keepdim=True) * (1.0 - torch.cos(rot_angle))
    p_new = p_new / torch.norm(p_new, dim=-1, keepdim=True)
    pts = torch.cat((p_new, depth.unsqueeze(-1)), dim=-1)
    depth_real = 1.0 / (depth + TINY_NUMBER) * torch.cos(theta) * ray_d_cos + d1
    return pts, depth_real


# ------------------------------------------------------------------------------------------------ embedding + MLP
def embed(x: torch.Tensor, n_freqs: int) -> torch.Tensor:
    """nerf_network.py:11-59 with max_freq_log2 = n_freqs - 1, log sampling: [x, sin(2^0 x), cos(2^0 x), ..., cos(2^(L-1) x)]
    -- the same channel order as nerf-ours' Embedder for any input dimension (3 -> 63 at L=10, 4 -> 84, 3 -> 27 at L=4)."""
    out = [x]
    for k in range(n_freqs):
        f = float(2.0 ** k)
        out += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(out, dim=-1)


def init_mlp_params(seed: int, input_ch: int = 63, input_ch_viewdirs: int = 27, D: int = 8, W: int = 256,
                    skips: Sequence[int] = (4,)) -> Dict[str, torch.Tensor]:
    """MLPNet.state_dict() keys / shapes (nerf_network.py:86-118) with nn.Linear-style U(+-1/sqrt(fan_in)) values drawn
    from numpy's MT19937 (fixtures store only the seed)."""
    rs = np.random.RandomState(seed)
    p: Dict[str, torch.Tensor] = {}

    def lin(name, fin, fout):
        b = 1.0 / math.sqrt(fin)
        p[name + ".weight"] = torch.from_numpy(rs.uniform(-b, b, (fout, fin)).astype(np.float32))
        p[name + ".bias"] = torch.from_numpy(rs.uniform(-b, b, (fout,)).astype(np.float32))

    dim = input_ch
    for i in range(D):
        lin("base_layers.%d.0" % i, dim, W)
        dim = W
        if i in skips and i != D - 1:
            dim += input_ch
    lin("sigma_layers.0", dim, 1)
    lin("base_remap_layers.0", dim, 256)
    lin("rgb_layers.0", 256 + input_ch_viewdirs, W // 2)
    lin("rgb_layers.2", W // 2, 3)
    return p


def mlp_forward(p: Dict[str, torch.Tensor], x: torch.Tensor, input_ch: int, skips: Sequence[int] = (4,)):
    """MLPNet.forward (nerf_network.py:120-142): returns (rgb [...,3] after the sigmoid, sigma [...] = |Linear|)."""
    D = sum(1 for k in p if k.startswith("base_layers.") and k.endswith(".weight"))
    pts = x[..., :input_ch]
    lin = lambda h, n: torch.nn.functional.linear(h, p[n + ".weight"], p[n + ".bias"])
    base = torch.relu(lin(pts, "base_layers.0.0"))
    for i in range(D - 1):
        if i in skips:
            base = torch.cat((pts, base), dim=-1)
        base = torch.relu(lin(base, "base_layers.%d.0" % (i + 1)))
    sigma = torch.abs(lin(base, "sigma_layers.0")).squeeze(-1)
    remap = lin(base, "base_remap_layers.0")
    n_view = p["rgb_layers.0.weight"].shape[1] - 256
    h = torch.relu(lin(torch.cat((remap, x[..., -n_view:]), dim=-1), "rgb_layers.0"))
    return torch.sigmoid(lin(h, "rgb_layers.2")), sigma


def mlp_params_to_nerf_layout(p: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """MLPNet state_dict -> nerf-ours NeRF state_dict names (model.py:20-34): the two modules compute the same GEMM chain
    (skip concat [input, h] in both, model.py:47 / nerf_network.py:129; rgb head input [feature, views] in both), so the
    foreground network can run on the existing kernels with raw = (pre-sigmoid rgb, pre-|.| sigma)."""
    out = {}
    for k, v in p.items():
        k2 = k
        if k.startswith("base_layers."):
            i = int(k.split(".")[1])
            k2 = "pts_linears.%d.%s" % (i, k.split(".")[-1])
        k2 = k2.replace("sigma_layers.0", "alpha_linear").replace("base_remap_layers.0", "feature_linear")
        k2 = k2.replace("rgb_layers.0", "views_linears.0").replace("rgb_layers.2", "rgb_linear")
        out[k2] = v
    return out


# ------------------------------------------------------------------------------------------------ NerfNet.forward
def nerfnet_forward(p_fg, p_bg, ray_o, ray_d, fg_z_max, fg_z_vals, bg_z_vals, n_freqs: int = 10, n_freqs_view: int = 4):
    """NerfNet.forward (ddp_model.py:74-143): foreground inside the unit sphere (last interval = fg_z_max - z_last),
    background on the inverted sphere (samples flipped so that they run near -> far, last interval HUGE_NUMBER),
    rgb = fg + bg_lambda * bg with bg_lambda = the foreground's final transmittance."""
    ray_d_norm = torch.norm(ray_d, dim=-1, keepdim=True)
    viewdirs = ray_d / ray_d_norm
    dots = list(ray_d.shape[:-1])
    # foreground
    N = fg_z_vals.shape[-1]
    o = ray_o.unsqueeze(-2).expand(dots + [N, 3])
    d = ray_d.unsqueeze(-2).expand(dots + [N, 3])
    v = viewdirs.unsqueeze(-2).expand(dots + [N, 3])
    fg_pts = o + fg_z_vals.unsqueeze(-1) * d
    rgb, sigma = mlp_forward(p_fg, torch.cat((embed(fg_pts, n_freqs), embed(v, n_freqs_view)), dim=-1), 3 + 6 * n_freqs)
    fg_dists = fg_z_vals[..., 1:] - fg_z_vals[..., :-1]
    fg_dists = ray_d_norm * torch.cat((fg_dists, fg_z_max.unsqueeze(-1) - fg_z_vals[..., -1:]), dim=-1)
    fg_alpha = 1.0 - torch.exp(-sigma * fg_dists)
    T = torch.cumprod(1.0 - fg_alpha + TINY_NUMBER, dim=-1)
    bg_lambda = T[..., -1]
    T = torch.cat((torch.ones_like(T[..., 0:1]), T[..., :-1]), dim=-1)
    fg_weights = fg_alpha * T
    fg_rgb = torch.sum(fg_weights.unsqueeze(-1) * rgb, dim=-2)
    fg_depth = torch.sum(fg_weights * fg_z_vals, dim=-1)
    # background
    N = bg_z_vals.shape[-1]
    o = ray_o.unsqueeze(-2).expand(dots + [N, 3])
    d = ray_d.unsqueeze(-2).expand(dots + [N, 3])
    v = viewdirs.unsqueeze(-2).expand(dots + [N, 3])
    bg_pts, _ = depth2pts_outside(o, d, bg_z_vals)
    x = torch.flip(torch.cat((embed(bg_pts, n_freqs), embed(v, n_freqs_view)), dim=-1), dims=[-2])
    bg_z = torch.flip(bg_z_vals, dims=[-1])
    bg_dists = bg_z[..., :-1] - bg_z[..., 1:]
    bg_dists = torch.cat((bg_dists, HUGE_NUMBER * torch.ones_like(bg_dists[..., 0:1])), dim=-1)
    rgb, sigma = mlp_forward(p_bg, x, 4 + 8 * n_freqs)
    bg_alpha = 1.0 - torch.exp(-sigma * bg_dists)
    T = torch.cumprod(1.0 - bg_alpha + TINY_NUMBER, dim=-1)[..., :-1]
    T = torch.cat((torch.ones_like(T[..., 0:1]), T), dim=-1)
    bg_weights = bg_alpha * T
    bg_rgb = bg_lambda.unsqueeze(-1) * torch.sum(bg_weights.unsqueeze(-1) * rgb, dim=-2)
    bg_depth = bg_lambda * torch.sum(bg_weights * bg_z, dim=-1)
    return {"rgb": fg_rgb + bg_rgb, "fg_weights": fg_weights, "bg_weights": bg_weights, "fg_rgb": fg_rgb,
            "fg_depth": fg_depth, "bg_rgb": bg_rgb, "bg_depth": bg_depth, "bg_lambda": bg_lambda}


def cascade_depths(ray_o, ray_d, n0: int, n1: int, ret0=None, fg_prev=None, bg_prev=None, t_fg=None, t_bg=None,
                   u_fg=None, u_bg=None):
    """The per-level sample placement of train_step (ddp_train_nerf.py:352-382).  Level 0 (ret0 None): fg depths linear
    from 1e-4 to the sphere exit (n0 samples, jittered by t_fg), bg inverse depths linspace(0,1) jittered by t_bg.
    Level 1: both resampled from the previous level's weights (mid-point bins, weights[..., 1:-1]) and merged with the
    previous depths by a sort.  QUIRK reproduced as is: NerfNet.forward returns bg_weights in the FLIPPED (1 -> 0) sample
    order (ddp_model.py:112-124) and this fork's train_step does not flip them back before resampling the 0 -> 1 ordered
    bg depths (:377-382; upstream nerf++ does)."""
    fg_far = intersect_sphere(ray_o, ray_d)
    if ret0 is None:
        fg_near = 1e-4 * torch.ones_like(fg_far)
        step = (fg_far - fg_near) / (n0 - 1)
        fg = torch.stack([fg_near + i * step for i in range(n0)], dim=-1)
        bg = torch.linspace(0.0, 1.0, n0).view([1] * (ray_d.dim() - 1) + [n0]).expand(list(ray_d.shape[:-1]) + [n0])
        if t_fg is not None:
            fg = perturb_samples(fg, t_fg)
        if t_bg is not None:
            bg = perturb_samples(bg, t_bg)
        return fg_far, fg, bg
    fg_w = ret0["fg_weights"].clone().detach()
    fg_mid = 0.5 * (fg_prev[..., 1:] + fg_prev[..., :-1])
    fg_s = sample_pdf(fg_mid, fg_w[..., 1:-1], n1, u_fg).detach()
    fg, _ = torch.sort(torch.cat((fg_prev, fg_s), dim=-1))
    bg_w = ret0["bg_weights"].clone().detach()
    bg_mid = 0.5 * (bg_prev[..., 1:] + bg_prev[..., :-1])
    bg_s = sample_pdf(bg_mid, bg_w[..., 1:-1], n1, u_bg).detach()
    bg, _ = torch.sort(torch.cat((bg_prev, bg_s), dim=-1))
    return fg_far, fg, bg


def train_step(levels, adams, ray_o, ray_d, rgb_gt, samples=(64, 128), t_fg=None, t_bg=None, u_fg=None, u_bg=None):
    """One iteration of nerf++-ours train_step for ONE batch (ddp_train_nerf.py:346-404, auto-exposure off): cascade level
    m has its own NerfNet ``levels[m] = (p_fg, p_bg)`` and its own Adam ``adams[m]`` (nerf_oracle.AdamState over
    list(p_fg.values()) + list(p_bg.values())); level 1 resamples both depth sets from level 0's weights (with the fork's
    un-flipped bg_weights, see cascade_depths).  Explicit uniforms replace torch.rand.  Returns per-level loss / gradients."""
    out, ret, fg_z, bg_z = [], None, None, None
    for m, (p_fg, p_bg) in enumerate(levels):
        params = list(p_fg.values()) + list(p_bg.values())
        for q in params:
            q.requires_grad_(True)
            q.grad = None
        if m == 0:
            fg_far, fg_z, bg_z = cascade_depths(ray_o, ray_d, samples[0], 0, t_fg=t_fg, t_bg=t_bg)
        else:
            fg_far, fg_z, bg_z = cascade_depths(ray_o, ray_d, samples[0], samples[1], ret0=ret, fg_prev=fg_z, bg_prev=bg_z,
                                                u_fg=u_fg, u_bg=u_bg)
        ret = nerfnet_forward(p_fg, p_bg, ray_o, ray_d, fg_far, fg_z, bg_z)
        loss = torch.mean((ret["rgb"] - rgb_gt) * (ret["rgb"] - rgb_gt))          # utils.img2mse (utils.py:12-14)
        loss.backward()
        grads = [q.grad.clone() for q in params]
        with torch.no_grad():
            adams[m].step(grads)
        for q in params:
            q.requires_grad_(False)
        ret = {k: v.detach() for k, v in ret.items()}
        out.append({"loss": float(loss.detach()), "grads": grads, "rgb": ret["rgb"], "fg_z": fg_z, "bg_z": bg_z})
    return out
