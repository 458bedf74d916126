"""Thin torch-facing wrappers over the C ABI: torch owns device memory and streams, libflnerf.so does
the arithmetic.  Every function here launches hand-written sm_100a kernels; nothing falls back to
torch math.  Autograd is provided for the two differentiable stages of the reference render loop
(the MLP query and raw2outputs); everything else is data (render.py:281 detaches the samples).
"""
import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import lib as L

MODE_FP32, MODE_BF16, MODE_BF16X3 = L.MODE_FP32, L.MODE_BF16, L.MODE_BF16X3


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ctx(t: torch.Tensor):
    if not t.is_cuda:
        raise L.FlnerfError("flnerf ops need CUDA tensors (got %s); there is no CPU fallback" % t.device)
    return C.c_void_p(L.context(t.device.index if t.device.index is not None else torch.cuda.current_device()))


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _alloc_bytes(nbytes: int, device) -> torch.Tensor:
    """1 KB-aligned byte buffer (torch's caching allocator hands out 512-byte aligned blocks)."""
    buf = torch.empty(int(nbytes) + 1024, dtype=torch.uint8, device=device)
    off = (-buf.data_ptr()) % 1024
    return buf[off:off + int(nbytes)]


# ------------------------------------------------------------------------------------------ rays
def raygen(H: int, W: int, K, c2w: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """get_rays (run_nerf_helpers.py:68-78)."""
    dev = c2w.device
    o = torch.empty(H, W, 3, dtype=torch.float32, device=dev)
    d = torch.empty(H, W, 3, dtype=torch.float32, device=dev)
    Kh = (C.c_double * 9)(*[float(K[i][j]) for i in range(3) for j in range(3)])
    Ph = (C.c_float * 12)(*c2w[:3, :4].detach().float().cpu().reshape(-1).tolist())
    L.check(L.load().flnerf_raygen(_ctx(o), H, W, Kh, Ph, _ptr(o), _ptr(d), _stream()), "flnerf_raygen")
    return o, d


def pack_rays(rays_o, rays_d, near: float, far: float, ndc: bool, H: int, W: int, focal: float) -> torch.Tensor:
    """ndc_rays + [o,d,near,far,viewdir] packing (run_nerf_helpers.py:91-108, render.py:59-80)."""
    o, d = _f32c(rays_o.reshape(-1, 3)), _f32c(rays_d.reshape(-1, 3))
    B = o.shape[0]
    out = torch.empty(B, 11, dtype=torch.float32, device=o.device)
    L.check(L.load().flnerf_pack_rays(_ctx(o), B, _ptr(o), _ptr(d), float(near), float(far), int(bool(ndc)), int(H),
                                      int(W), float(focal), _ptr(out), _stream()), "flnerf_pack_rays")
    return out


_TVALS = {}


def _t_vals(n: int, device) -> torch.Tensor:
    key = (n, str(device))
    if key not in _TVALS:
        _TVALS[key] = torch.linspace(0.0, 1.0, steps=n).to(device)   # render.py:244 (host, bit-identical to torch)
    return _TVALS[key]


def coarse_depths(rays11, n_samples: int, perturb: bool, lindisp: bool, t_rand=None, seed: int = 0, offset: int = 0):
    """render.py:244-266."""
    B = rays11.shape[0]
    z = torch.empty(B, n_samples, dtype=torch.float32, device=rays11.device)
    tr = None if t_rand is None else _f32c(t_rand)
    L.check(L.load().flnerf_coarse_depths(_ctx(z), B, n_samples, _ptr(rays11), _ptr(_t_vals(n_samples, z.device)),
                                          _ptr(tr), int(bool(perturb)), int(bool(lindisp)), seed, offset, _ptr(z),
                                          _stream()), "flnerf_coarse_depths")
    return z


def posenc(x: torch.Tensor, n_freq: int) -> torch.Tensor:
    """Embedder.embed (run_nerf_helpers.py:15-45) for 3-vectors."""
    shp = x.shape
    xf = _f32c(x.reshape(-1, 3))
    out = torch.empty(xf.shape[0], 3 + 6 * n_freq, dtype=torch.float32, device=x.device)
    L.check(L.load().flnerf_posenc(_ctx(xf), xf.shape[0], n_freq, _ptr(xf), _ptr(out), _stream()), "flnerf_posenc")
    return out.reshape(*shp[:-1], 3 + 6 * n_freq)


def encode_f32(rays11, z) -> torch.Tensor:
    B, S = z.shape
    x = torch.empty(B * S, 90, dtype=torch.float32, device=z.device)
    L.check(L.load().flnerf_encode_f32(_ctx(z), B, S, _ptr(rays11), _ptr(z), _ptr(x), _stream()), "flnerf_encode_f32")
    return x


def padded_rows(n: int) -> int:
    return (n + 255) // 256 * 256


def encode_tc(rays11, z, mode=MODE_BF16):
    """Sample points + PE as tensor-core tiles; MODE_BF16X3 writes the hi and the lo tile set back to back."""
    B, S = z.shape
    x3 = mode == MODE_BF16X3
    tiles = _alloc_bytes(padded_rows(B * S) // 128 * 16384 * (2 if x3 else 1), z.device)
    dirpe = torch.empty(B, 32, dtype=torch.float32, device=z.device)
    fn = L.load().flnerf_encode_tc_x3 if x3 else L.load().flnerf_encode_tc
    L.check(fn(_ctx(z), B, S, _ptr(rays11), _ptr(z), _ptr(tiles), _ptr(dirpe), _stream()), "flnerf_encode_tc")
    return tiles, dirpe


def encode_frame_tc(mode, H, W, K, c2w, near, far, ndc, lindisp, pixel0: int, B: int, n_samples: int):
    """The eval path's fused front end (render.py:52-80, 244-249, run_nerf.py:50-64): pixels [pixel0, pixel0+B) of the frame
    -> (rays11[B,11], z[B,S], pe_tiles, dirpe[B,32]) in ONE launch (no [H,W,3] ray images, no separate packing / depth /
    direction-PE kernels)."""
    dev = c2w.device
    x3 = mode == MODE_BF16X3
    rays11 = torch.empty(B, 11, dtype=torch.float32, device=dev)
    z = torch.empty(B, n_samples, dtype=torch.float32, device=dev)
    tiles = _alloc_bytes(padded_rows(B * n_samples) // 128 * 16384 * (2 if x3 else 1), dev)
    dirpe = torch.empty(B, 32, dtype=torch.float32, device=dev)
    Kh = (C.c_double * 9)(*[float(K[i][j]) for i in range(3) for j in range(3)])
    Ph = (C.c_float * 12)(*c2w[:3, :4].detach().float().cpu().reshape(-1).tolist())
    L.check(L.load().flnerf_encode_frame_tc(_ctx(rays11), int(x3), int(H), int(W), Kh, Ph, float(near), float(far), int(bool(ndc)),
                                            int(bool(lindisp)), int(pixel0), int(B), int(n_samples),
                                            _ptr(_t_vals(n_samples, dev)), _ptr(rays11), _ptr(z), _ptr(tiles), _ptr(dirpe),
                                            _stream()), "flnerf_encode_frame_tc")
    return rays11, z, tiles, dirpe


def ssim_psnr(img0, img1, max_val: float = 1.0):
    """(mean SSIM, PSNR) of two [H,W,3] images as a 2-element float64 DEVICE tensor (compute_ssim of run_nerf_helpers.py:158-228
    and render.py:120) -- one kernel, no host sync."""
    a, b = _f32c(img0), _f32c(img1)
    H, W, _ = a.shape
    sums = torch.empty(2, dtype=torch.float64, device=a.device)
    L.check(L.load().flnerf_ssim_psnr(_ctx(a), int(H), int(W), _ptr(a), _ptr(b), float(max_val), _ptr(sums), _stream()),
            "flnerf_ssim_psnr")
    n = 3.0 * H * W
    return torch.stack([sums[0] / n, -10.0 * torch.log10(sums[1] / n)])


def pack_x90(x90, mode=MODE_BF16):
    n = x90.shape[0]
    x3 = mode == MODE_BF16X3
    tiles = _alloc_bytes(padded_rows(n) // 128 * 16384 * (2 if x3 else 1), x90.device)
    dirpe = torch.empty(n, 32, dtype=torch.float32, device=x90.device)
    fn = L.load().flnerf_pack_x90_x3 if x3 else L.load().flnerf_pack_x90
    L.check(fn(_ctx(x90), n, _ptr(x90), _ptr(tiles), _ptr(dirpe), _stream()), "flnerf_pack_x90")
    return tiles, dirpe


# ------------------------------------------------------------------------------------------ MLP
def mlp_pack_weights(flat_params: torch.Tensor, packed: Optional[torch.Tensor] = None) -> torch.Tensor:
    if packed is None:
        packed = _alloc_bytes(L.load().flnerf_mlp_packed_bytes(), flat_params.device)
    L.check(L.load().flnerf_mlp_pack_weights(_ctx(flat_params), _ptr(flat_params), _ptr(packed), _stream()),
            "flnerf_mlp_pack_weights")
    return packed


def mlp_forward(mode, flat_params, packed, x, dirpe, n: int, S: int, training: bool):
    """Returns (raw[n,4], stash)."""
    dev = flat_params.device
    raw = torch.empty(n, 4, dtype=torch.float32, device=dev)
    stash = _alloc_bytes(L.load().flnerf_mlp_stash_bytes(mode, n, S, int(training)), dev)
    L.check(L.load().flnerf_mlp_forward(_ctx(raw), mode, _ptr(flat_params), _ptr(packed), n, S, _ptr(x), _ptr(dirpe),
                                        _ptr(raw), _ptr(stash), int(training), _stream()), "flnerf_mlp_forward")
    return raw, stash


def mlp_backward(mode, flat_params, packed, x, dirpe, stash, draw, flat_grad, n: int, S: int, stages: int = 7, ws=None):
    """flat_grad += dL/dparams.  ``stages`` selects kernels (1 dgrad, 2 wgrad, 4 heads) for per-kernel timing."""
    ws_bytes = L.load().flnerf_mlp_bwd_workspace_bytes(mode, n)
    if ws is None:
        ws = _alloc_bytes(ws_bytes, flat_params.device)
    draw = _f32c(draw.reshape(n, 4))
    L.check(L.load().flnerf_mlp_backward_stages(_ctx(draw), mode, _ptr(flat_params), _ptr(packed), n, S, _ptr(x),
                                                _ptr(dirpe), _ptr(stash), _ptr(draw), _ptr(flat_grad), _ptr(ws),
                                                ws_bytes, int(stages), _stream()), "flnerf_mlp_backward")
    return ws


# ------------------------------------------------------------------------------------------ compositing
def composite_forward(raw, z, rays_d, noise, white_bkgd: bool, rays_d_stride: int = 3, want_weights: bool = True):
    B, S = z.shape
    dev = z.device
    rgb = torch.empty(B, 3, dtype=torch.float32, device=dev)
    disp = torch.empty(B, dtype=torch.float32, device=dev)
    acc = torch.empty(B, dtype=torch.float32, device=dev)
    depth = torch.empty(B, dtype=torch.float32, device=dev)
    w = torch.empty(B, S, dtype=torch.float32, device=dev) if want_weights else None
    L.check(L.load().flnerf_composite_forward(_ctx(z), B, S, _ptr(raw), _ptr(z), _ptr(rays_d), rays_d_stride,
                                              _ptr(noise), int(bool(white_bkgd)), _ptr(rgb), _ptr(disp), _ptr(acc),
                                              _ptr(depth), _ptr(w), _stream()), "flnerf_composite_forward")
    return rgb, disp, acc, w, depth


def composite_backward(raw, z, rays_d, noise, white_bkgd, g_rgb, g_disp, g_acc, g_depth, rays_d_stride: int = 3):
    B, S = z.shape
    draw = torch.empty(B, S, 4, dtype=torch.float32, device=z.device)
    gs = [None if g is None else _f32c(g) for g in (g_rgb, g_disp, g_acc, g_depth)]
    L.check(L.load().flnerf_composite_backward(_ctx(z), B, S, _ptr(raw), _ptr(z), _ptr(rays_d), rays_d_stride,
                                               _ptr(noise), int(bool(white_bkgd)), _ptr(gs[0]), _ptr(gs[1]),
                                               _ptr(gs[2]), _ptr(gs[3]), _ptr(draw), _stream()),
            "flnerf_composite_backward")
    return draw


class CompositeFn(torch.autograd.Function):
    """raw2outputs (render.py:149-192) with the analytic backward of SURVEY appendix A.2."""

    @staticmethod
    def forward(ctx, raw, z, rays_d, noise, white_bkgd):
        raw, z, rays_d = _f32c(raw), _f32c(z), _f32c(rays_d)
        noise = None if noise is None else _f32c(noise)
        rgb, disp, acc, w, depth = composite_forward(raw, z, rays_d, noise, white_bkgd)
        ctx.save_for_backward(raw, z, rays_d, noise)
        ctx.white = bool(white_bkgd)
        ctx.mark_non_differentiable(w)     # only consumed by sample_pdf, which the reference detaches
        return rgb, disp, acc, w, depth

    @staticmethod
    def backward(ctx, g_rgb, g_disp, g_acc, g_w, g_depth):
        raw, z, rays_d, noise = ctx.saved_tensors
        draw = composite_backward(raw, z, rays_d, noise, ctx.white, g_rgb, g_disp, g_acc, g_depth)
        return draw, None, None, None, None


def sample_pdf_merge(z, weights, n_fine: int, det: bool, u=None, seed: int = 0, offset: int = 0, want_samples=True):
    """sample_pdf + sort-merge + z_std (run_nerf_helpers.py:112-155, render.py:279-284,299)."""
    B, Nc = z.shape
    dev = z.device
    merged = torch.empty(B, Nc + n_fine, dtype=torch.float32, device=dev)
    zs = torch.empty(B, n_fine, dtype=torch.float32, device=dev) if want_samples else None
    zstd = torch.empty(B, dtype=torch.float32, device=dev)
    uu = None if u is None else _f32c(u)
    L.check(L.load().flnerf_sample_pdf_merge(_ctx(z), B, Nc, n_fine, _ptr(_f32c(z)), _ptr(_f32c(weights)), _ptr(uu),
                                             int(bool(det)), seed, offset, _ptr(merged), _ptr(zs), _ptr(zstd),
                                             _stream()), "flnerf_sample_pdf_merge")
    return merged, zs, zstd


def sample_pdf_bins(bins, weights, n_samples: int, det: bool, u=None, seed: int = 0, offset: int = 0):
    """sample_pdf on explicit bins (run_nerf_helpers.py:112-155)."""
    B, nb = bins.shape
    out = torch.empty(B, n_samples, dtype=torch.float32, device=bins.device)
    uu = None if u is None else _f32c(u)
    L.check(L.load().flnerf_sample_pdf(_ctx(bins), B, nb, n_samples, _ptr(_f32c(bins)), _ptr(_f32c(weights)), _ptr(uu),
                                       int(bool(det)), seed, offset, _ptr(out), _stream()), "flnerf_sample_pdf")
    return out


# ------------------------------------------------------------------------------------------ loss / optimiser
def mse_leafmax(rgb, rgb0, target, denom: int, leaf_gid=None, leaf_max=None, want_grads=True):
    B = rgb.shape[0]
    dev = rgb.device
    loss = torch.empty(2, dtype=torch.float32, device=dev)
    d_rgb = torch.empty(B, 3, dtype=torch.float32, device=dev) if want_grads else None
    d_rgb0 = torch.empty(B, 3, dtype=torch.float32, device=dev) if (want_grads and rgb0 is not None) else None
    L.check(L.load().flnerf_mse_leafmax(_ctx(rgb), B, _ptr(_f32c(rgb)), _ptr(None if rgb0 is None else _f32c(rgb0)),
                                        _ptr(_f32c(target)), int(denom), _ptr(leaf_gid), _ptr(loss), _ptr(d_rgb),
                                        _ptr(d_rgb0), _ptr(leaf_max), _stream()), "flnerf_mse_leafmax")
    return loss, d_rgb, d_rgb0


def leaf_sum(pred, target, leaf_gid, leaf_sum_t, leaf_cnt_t):
    """Accumulates the nerf++ refinement statistic (nerf++-ours/tree.py:613-622) of one batch into the per-leaf tables."""
    B = pred.shape[0]
    L.check(L.load().flnerf_leaf_sum(_ctx(pred), B, _ptr(_f32c(pred)), _ptr(_f32c(target)), _ptr(leaf_gid), _ptr(leaf_sum_t),
                                     _ptr(leaf_cnt_t), _stream()), "flnerf_leaf_sum")


def leaf_mean(leaf_sum_t, leaf_cnt_t, leaf_stat):
    L.check(L.load().flnerf_leaf_mean(_ctx(leaf_stat), leaf_stat.numel(), _ptr(leaf_sum_t), _ptr(leaf_cnt_t), _ptr(leaf_stat),
                                      _stream()), "flnerf_leaf_mean")


def adam_step(param, m, v, grad, lr: float, b1: float, b2: float, eps: float, t: int):
    L.check(L.load().flnerf_adam_step(_ctx(param), param.numel(), _ptr(param), _ptr(m), _ptr(v), _ptr(grad), lr, b1, b2,
                                      eps, int(t), _stream()), "flnerf_adam_step")


# ------------------------------------------------------------------------------------------ step record (CUDA graphs)
STEP_RECORD_BYTES = 24


def set_step_record(rec: Optional[torch.Tensor], device=None):
    """Attach (or with None detach) the device-side per-step scalars (include/flnerf.h: flnerf_step_record)."""
    idx = rec.device.index if rec is not None else (device.index if device is not None else torch.cuda.current_device())
    L.check(L.load().flnerf_set_step_record(C.c_void_p(L.context(idx)), _ptr(rec)), "flnerf_set_step_record")


def step_record_write(rec: torch.Tensor, first: int, rng_offset: int, lr: float, b1: float, b2: float, t: int):
    L.check(L.load().flnerf_step_record_write(_ctx(rec), _ptr(rec), int(first), int(rng_offset), float(lr), float(b1),
                                              float(b2), int(t), _stream()), "flnerf_step_record_write")


# ------------------------------------------------------------------------------------------ quadtree
def qt_init(n_images, cap, H, W, max_depth, boxes, count, min_area):
    L.check(L.load().flnerf_qt_init(_ctx(boxes), n_images, cap, H, W, max_depth, _ptr(boxes), _ptr(count),
                                    _ptr(min_area), _stream()), "flnerf_qt_init")


def qt_refine(n_images, cap, boxes_in, count_in, min_area, leaf_max, thres, boxes_out, count_out):
    L.check(L.load().flnerf_qt_refine(_ctx(boxes_in), n_images, cap, _ptr(boxes_in), _ptr(count_in), _ptr(min_area),
                                      _ptr(leaf_max), float(thres), _ptr(boxes_out), _ptr(count_out), _stream()),
            "flnerf_qt_refine")


def qt_count(n_images, cap, boxes, count, min_area, rays_per_pixel, ray_offset):
    L.check(L.load().flnerf_qt_count(_ctx(boxes), n_images, cap, _ptr(boxes), _ptr(count), _ptr(min_area),
                                     float(rays_per_pixel), _ptr(ray_offset), _stream()), "flnerf_qt_count")


def qt_emit(n_images, cap, W, boxes, count, ray_offset, n_rays, seed, ray_pix, ray_gid):
    L.check(L.load().flnerf_qt_emit(_ctx(boxes), n_images, cap, W, _ptr(boxes), _ptr(count), _ptr(ray_offset),
                                    int(n_rays), int(seed), _ptr(ray_pix), _ptr(ray_gid), _stream()), "flnerf_qt_emit")


def sharp_map(images):
    """image_process.py:26-39 for every image: [n,H,W,3] fp32 -> [n,H,W] fp32 sharpness (gray local std)."""
    n, H, W, _ = images.shape
    out = torch.empty(n, H, W, dtype=torch.float32, device=images.device)
    L.check(L.load().flnerf_sharp_map(_ctx(images), int(n), int(H), int(W), _ptr(images), _ptr(out), _stream()),
            "flnerf_sharp_map")
    return out


def qt_prob_prepare(n_images, cap, H, W, boxes, count, sharp):
    """Per-leaf to_prob_v2 tables (image_process.py:59-74): returns (row_offset, row_cdf, leaf_thr)."""
    dev = boxes.device
    row_offset = torch.empty(n_images * cap + 1, dtype=torch.int64, device=dev)
    L.check(L.load().flnerf_qt_prob_rows(_ctx(boxes), n_images, cap, _ptr(boxes), _ptr(count), _ptr(row_offset), _stream()),
            "flnerf_qt_prob_rows")
    rows = int(row_offset[-1].item())
    row_cdf = torch.empty(max(rows, 1), dtype=torch.float64, device=dev)
    leaf_thr = torch.zeros(n_images * cap, dtype=torch.float64, device=dev)
    L.check(L.load().flnerf_qt_prob_prepare(_ctx(boxes), n_images, cap, int(H), int(W), _ptr(boxes), _ptr(count), _ptr(sharp),
                                            _ptr(row_offset), _ptr(row_cdf), _ptr(leaf_thr), _stream()),
            "flnerf_qt_prob_prepare")
    return row_offset, row_cdf, leaf_thr


def qt_emit_prob(n_images, cap, H, W, boxes, count, ray_offset, n_rays, seed, rand_frac, sharp, tables, ray_pix, ray_gid,
                 u=None, shuffle=True):
    row_offset, row_cdf, leaf_thr = tables
    L.check(L.load().flnerf_qt_emit_prob(_ctx(boxes), n_images, cap, int(H), int(W), _ptr(boxes), _ptr(count),
                                         _ptr(ray_offset), int(n_rays), int(seed), float(rand_frac), _ptr(sharp),
                                         _ptr(row_offset), _ptr(row_cdf), _ptr(leaf_thr), _ptr(u), 1 if shuffle else 0,
                                         _ptr(ray_pix), _ptr(ray_gid), _stream()), "flnerf_qt_emit_prob")


def gather_batch(B, first, stride, ray_pix, ray_gid, cap, H, W, K, poses, images, want_gid=True, lut=None):
    """images: fp32 [n,H,W,3], or uint8 [n,H,W,3] with lut[256] = float32(u / 255.) (the loaders' division, done per batch)."""
    dev = images.device
    o = torch.empty(B, 3, dtype=torch.float32, device=dev)
    d = torch.empty(B, 3, dtype=torch.float32, device=dev)
    t = torch.empty(B, 3, dtype=torch.float32, device=dev)
    gid = torch.empty(B, dtype=torch.int32, device=dev) if want_gid else None
    Kh = (C.c_double * 9)(*[float(K[i][j]) for i in range(3) for j in range(3)])
    if images.dtype == torch.uint8:
        L.check(L.load().flnerf_gather_batch_u8(_ctx(images), int(B), int(first), int(stride), _ptr(ray_pix), _ptr(ray_gid),
                                                int(cap), int(H), int(W), Kh, _ptr(poses), _ptr(images), _ptr(lut), _ptr(o),
                                                _ptr(d), _ptr(t), _ptr(gid), _stream()), "flnerf_gather_batch_u8")
        return o, d, t, gid
    L.check(L.load().flnerf_gather_batch(_ctx(images), int(B), int(first), int(stride), _ptr(ray_pix), _ptr(ray_gid),
                                         int(cap), int(H), int(W), Kh, _ptr(poses), _ptr(images), _ptr(o), _ptr(d),
                                         _ptr(t), _ptr(gid), _stream()), "flnerf_gather_batch")
    return o, d, t, gid


def qt_emit_sub(n_images, cap, boxes, ray_offset, n_rays, seed, ray_xy, ray_gid):
    """gen_rays_v3's sub-pixel emission (tree.py:231-268): ray_xy[N,2] on a 1/1000-pixel grid, shuffled like qt_emit."""
    L.check(L.load().flnerf_qt_emit_sub(_ctx(boxes), n_images, cap, _ptr(boxes), _ptr(ray_offset), int(n_rays), int(seed),
                                        _ptr(ray_xy), _ptr(ray_gid), _stream()), "flnerf_qt_emit_sub")


def gather_sub(ray_xy, ray_gid, cap, H, W, K, poses, images, lut=None):
    """gen_rays_v3's gather (tree.py:270-285): F.grid_sample(bilinear, zeros, align_corners=False) of the colour / direction /
    origin images at sub-pixel positions, with the reference's transposed grid -> (origins, dirs, rgb) [N,3]."""
    B, dev = ray_xy.shape[0], images.device
    o = torch.empty(B, 3, dtype=torch.float32, device=dev)
    d = torch.empty(B, 3, dtype=torch.float32, device=dev)
    t = torch.empty(B, 3, dtype=torch.float32, device=dev)
    Kh = (C.c_double * 9)(*[float(K[i][j]) for i in range(3) for j in range(3)])
    L.check(L.load().flnerf_gather_sub(_ctx(images), int(B), _ptr(_f32c(ray_xy)), _ptr(ray_gid), int(cap), int(H), int(W), Kh,
                                       _ptr(poses), _ptr(images), _ptr(lut if images.dtype == torch.uint8 else None), _ptr(o),
                                       _ptr(d), _ptr(t), _stream()), "flnerf_gather_sub")
    return o, d, t


# ------------------------------------------------------------------------------------------ nerf++ building blocks
# (the next row: parity-tested kernels, no complete path yet -- see DESIGN.md section 8)
def pp_depths0(rays_o, rays_d, N, perturb, t_fg=None, t_bg=None, seed=0, offset=0):
    """ddp_train_nerf.py:352-366 -> (fg_far[B], fg_z[B,N], bg_z[B,N])."""
    B, dev = rays_o.shape[0], rays_o.device
    fg_far = torch.empty(B, dtype=torch.float32, device=dev)
    fg_z = torch.empty(B, N, dtype=torch.float32, device=dev)
    bg_z = torch.empty(B, N, dtype=torch.float32, device=dev)
    L.check(L.load().flnerf_pp_depths0(_ctx(rays_o), B, int(N), _ptr(_f32c(rays_o)), _ptr(_f32c(rays_d)), _ptr(t_fg), _ptr(t_bg),
                                       int(bool(perturb)), int(seed), int(offset), _ptr(fg_far), _ptr(fg_z), _ptr(bg_z),
                                       _stream()), "flnerf_pp_depths0")
    return fg_far, fg_z, bg_z


def pp_bg_encode(rays_o, rays_d, bg_z, want_pts=False):
    """ddp_model.py:16-45,109-117 -> (x111[B,N,111] flipped, bg_z_flip[B,N], pts4[B,N,4] or None)."""
    B, N = bg_z.shape
    dev = bg_z.device
    x = torch.empty(B, N, 111, dtype=torch.float32, device=dev)
    zf = torch.empty(B, N, dtype=torch.float32, device=dev)
    pts = torch.empty(B, N, 4, dtype=torch.float32, device=dev) if want_pts else None
    L.check(L.load().flnerf_pp_bg_encode(_ctx(bg_z), B, N, _ptr(_f32c(rays_o)), _ptr(_f32c(rays_d)), _ptr(_f32c(bg_z)), _ptr(x),
                                         _ptr(zf), _ptr(pts), _stream()), "flnerf_pp_bg_encode")
    return x, zf, pts


def pp_composite_forward(raw_fg, fg_z, fg_far, raw_bg, bg_z_flip, rays_d):
    """ddp_model.py:93-133 on raw network outputs -> (rgb[B,3], fg_weights, bg_weights, aux9[B,9])."""
    B, Sf = fg_z.shape
    Sb = bg_z_flip.shape[1]
    dev = fg_z.device
    rgb = torch.empty(B, 3, dtype=torch.float32, device=dev)
    fw = torch.empty(B, Sf, dtype=torch.float32, device=dev)
    bw = torch.empty(B, Sb, dtype=torch.float32, device=dev)
    aux = torch.empty(B, 9, dtype=torch.float32, device=dev)
    L.check(L.load().flnerf_pp_composite_forward(_ctx(fg_z), B, Sf, Sb, _ptr(_f32c(raw_fg)), _ptr(_f32c(fg_z)), _ptr(_f32c(fg_far)),
                                                 _ptr(_f32c(raw_bg)), _ptr(_f32c(bg_z_flip)), _ptr(_f32c(rays_d)), _ptr(rgb),
                                                 _ptr(fw), _ptr(bw), _ptr(aux), _stream()), "flnerf_pp_composite_forward")
    return rgb, fw, bw, aux


def pp_composite_backward(raw_fg, fg_z, fg_far, raw_bg, bg_z_flip, rays_d, g_rgb):
    B, Sf = fg_z.shape
    Sb = bg_z_flip.shape[1]
    dfg, dbg = torch.empty_like(raw_fg, dtype=torch.float32), torch.empty_like(raw_bg, dtype=torch.float32)
    L.check(L.load().flnerf_pp_composite_backward(_ctx(fg_z), B, Sf, Sb, _ptr(_f32c(raw_fg)), _ptr(_f32c(fg_z)), _ptr(_f32c(fg_far)),
                                                  _ptr(_f32c(raw_bg)), _ptr(_f32c(bg_z_flip)), _ptr(_f32c(rays_d)),
                                                  _ptr(_f32c(g_rgb)), _ptr(dfg), _ptr(dbg), _stream()),
            "flnerf_pp_composite_backward")
    return dfg, dbg


def pp_sample_pdf_merge(z, weights, Nf, det, u=None, seed=0, offset=0):
    """ddp_train_nerf.py:369-382 -> (z_merged[B,Nc+Nf], z_samples[B,Nf])."""
    B, Nc = z.shape
    dev = z.device
    zm = torch.empty(B, Nc + Nf, dtype=torch.float32, device=dev)
    zs = torch.empty(B, Nf, dtype=torch.float32, device=dev)
    L.check(L.load().flnerf_pp_sample_pdf_merge(_ctx(z), B, Nc, int(Nf), _ptr(_f32c(z)), _ptr(_f32c(weights)), _ptr(u),
                                                int(bool(det)), int(seed), int(offset), _ptr(zm), _ptr(zs), _stream()),
            "flnerf_pp_sample_pdf_merge")
    return zm, zs


def pack_xrows(mode, in_pts, x, S):
    """Already-embedded fp32 rows x[n, in_pts + 27] (S rows per ray) -> (pe_tiles, dirpe[n/S, 32]) for the tensor-core MLP of a
    network with in_pts position channels (ceil(in_pts/64) slabs per 128-row tile; MODE_BF16X3: hi set then lo set)."""
    n = x.shape[0]
    x3 = mode == MODE_BF16X3
    slabs = (in_pts + 63) // 64
    tiles = _alloc_bytes(padded_rows(n) // 128 * 16384 * slabs * (2 if x3 else 1), x.device)
    dirpe = torch.empty(n // S, 32, dtype=torch.float32, device=x.device)
    L.check(L.load().flnerf_pack_xrows(_ctx(x), int(x3), int(in_pts), 27, int(n), int(S), _ptr(_f32c(x)), _ptr(tiles), _ptr(dirpe),
                                       _stream()), "flnerf_pack_xrows")
    return tiles, dirpe


def mlp_pack_weights_g(in_pts, flat_params, packed=None):
    if packed is None:
        packed = _alloc_bytes(L.load().flnerf_mlp_packed_bytes_g(int(in_pts)), flat_params.device)
    L.check(L.load().flnerf_mlp_pack_weights_g(_ctx(flat_params), int(in_pts), _ptr(flat_params), _ptr(packed), _stream()),
            "flnerf_mlp_pack_weights_g")
    return packed


def mlp_forward_g(mode, in_pts, flat_params, packed, tiles, dirpe, n: int, S: int, training: bool):
    """The tensor-core MLP of a network with in_pts (63 / 84) position channels -> (raw[n,4], stash)."""
    dev = flat_params.device
    raw = torch.empty(n, 4, dtype=torch.float32, device=dev)
    stash = _alloc_bytes(L.load().flnerf_mlp_stash_bytes(mode, n, S, int(training)), dev)
    L.check(L.load().flnerf_mlp_forward_g(_ctx(raw), mode, int(in_pts), _ptr(flat_params), _ptr(packed), n, S, _ptr(tiles),
                                          _ptr(dirpe), _ptr(raw), _ptr(stash), int(training), _stream()), "flnerf_mlp_forward_g")
    return raw, stash


def mlp_backward_g(mode, in_pts, flat_params, packed, tiles, dirpe, stash, draw, flat_grad, n: int, S: int):
    ws_bytes = L.load().flnerf_mlp_bwd_workspace_bytes(mode, n)
    ws = _alloc_bytes(ws_bytes, flat_params.device)
    draw = _f32c(draw.reshape(n, 4))
    L.check(L.load().flnerf_mlp_backward_g(_ctx(draw), mode, int(in_pts), _ptr(flat_params), _ptr(packed), n, S, _ptr(tiles),
                                           _ptr(dirpe), _ptr(stash), _ptr(draw), _ptr(flat_grad), _ptr(ws), ws_bytes, _stream()),
            "flnerf_mlp_backward_g")


def mlp_fp32_forward_g(in_pts, in_views, flat_params, x, n):
    """The fp32 MLP with in_pts position / in_views view channels (84 / 27 = the nerf++ background network)."""
    dev = flat_params.device
    raw = torch.empty(n, 4, dtype=torch.float32, device=dev)
    stash = _alloc_bytes(L.load().flnerf_mlp_stash_bytes(MODE_FP32, n, 1, 1), dev)
    L.check(L.load().flnerf_mlp_fp32_forward_g(_ctx(flat_params), int(in_pts), int(in_views), _ptr(flat_params), int(n),
                                               _ptr(_f32c(x)), _ptr(raw), _ptr(stash), _stream()), "flnerf_mlp_fp32_forward_g")
    return raw, stash


def mlp_fp32_backward_g(in_pts, in_views, flat_params, x, stash, draw, grads, n):
    nbytes = L.load().flnerf_mlp_bwd_workspace_bytes(MODE_FP32, n)
    ws = _alloc_bytes(nbytes, flat_params.device)
    L.check(L.load().flnerf_mlp_fp32_backward_g(_ctx(flat_params), int(in_pts), int(in_views), _ptr(flat_params), int(n),
                                                _ptr(_f32c(x)), _ptr(stash), _ptr(_f32c(draw)), _ptr(grads), _ptr(ws),
                                                int(nbytes), _stream()), "flnerf_mlp_fp32_backward_g")
