"""The nerf++ row: NerfNet.forward (nerf++-ours/ddp_model.py:74-143) and its backward as a sequence of
libflnerf.so kernels -- foreground network on the nerf-ours encode + MLP kernels, background network (84 position channels:
two 64-wide input slabs) through ``flnerf_pp_bg_encode`` and the generic-width MLP entry points, both composited by
``flnerf_pp_composite_forward/backward`` -- in any of the three MLP modes: fp32 (CUDA cores), bf16 and bf16x3 (tcgen05; the
split-precision mode meets the fp32 tolerances).  ``cascade_train_step`` is one iteration of ddp_train_nerf.train_step."""
from typing import Dict

import torch

from . import ops
from .lib import FlnerfError

_ORDER = [("base_layers.%d.0" % i, "pts_linears.%d" % i) for i in range(8)] + \
         [("rgb_layers.0", "views_linears.0"), ("base_remap_layers.0", "feature_linear"), ("sigma_layers.0", "alpha_linear"),
          ("rgb_layers.2", "rgb_linear")]


def flat_from_mlpnet(state: Dict[str, torch.Tensor], device) -> torch.Tensor:
    """MLPNet.state_dict() (nerf_network.py:86-118) -> the flat fp32 parameter buffer of the flnerf kernels (nerf-ours
    parameters() order: pts_linears.0..7, views_linears.0, feature_linear, alpha_linear, rgb_linear)."""
    return torch.cat([state[a + s].reshape(-1).float() for a, _ in _ORDER for s in (".weight", ".bias")]).to(device).contiguous()


def mlpnet_from_flat(flat: torch.Tensor, like: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Inverse of flat_from_mlpnet (used to hand gradients back under the reference's parameter names)."""
    out, off = {}, 0
    for a, _ in _ORDER:
        for s in (".weight", ".bias"):
            n = like[a + s].numel()
            out[a + s] = flat[off:off + n].view(like[a + s].shape)
            off += n
    if off != flat.numel():
        raise FlnerfError("parameter buffer has %d floats, the MLPNet needs %d" % (flat.numel(), off))
    return out


_MODES = {"fp32": ops.MODE_FP32, "bf16": ops.MODE_BF16, "bf16x3": ops.MODE_BF16X3}


class NerfNet:
    """NerfNet.forward + backward without autograd over two flat parameter buffers (foreground 63, background 84 position
    channels).  precision: "fp32" (CUDA-core parity path), "bf16" or "bf16x3" (tensor cores)."""

    def __init__(self, flat_fg: torch.Tensor, flat_bg: torch.Tensor, in_pts_bg: int = 84, in_views: int = 27,
                 precision: str = "fp32"):
        self.fg, self.bg, self.in_bg, self.in_views = flat_fg, flat_bg, int(in_pts_bg), int(in_views)
        self.grad_fg, self.grad_bg = torch.zeros_like(flat_fg), torch.zeros_like(flat_bg)
        self.mode = _MODES[precision]
        self._packed_fg = self._packed_bg = None
        self._saved = None

    def _pack(self):
        """fp32 master weights -> tensor-core images (once per optimiser step; the buffers keep their addresses)."""
        self._packed_fg = ops.mlp_pack_weights(self.fg, self._packed_fg)
        self._packed_bg = ops.mlp_pack_weights_g(self.in_bg, self.bg, self._packed_bg)

    @torch.no_grad()
    def forward(self, ray_o, ray_d, fg_z_max, fg_z_vals, bg_z_vals, training=True):
        B, Sf = fg_z_vals.shape
        Sb = bg_z_vals.shape[1]
        rays11 = ops.pack_rays(ray_o, ray_d, 0.0, 1.0, False, 1, 1, 1.0)       # o, d, -, -, viewdir = d/|d|
        x_bg, bg_flip, _ = ops.pp_bg_encode(ray_o, ray_d, bg_z_vals)
        if self.mode == ops.MODE_FP32:
            x_fg, dp_fg, dp_bg = ops.encode_f32(rays11, fg_z_vals), None, None
            raw_fg, st_fg = ops.mlp_forward(ops.MODE_FP32, self.fg, None, x_fg, None, B * Sf, Sf, True)
            x_bg = x_bg.view(B * Sb, -1)
            raw_bg, st_bg = ops.mlp_fp32_forward_g(self.in_bg, self.in_views, self.bg, x_bg, B * Sb)
        else:
            self._pack()
            x_fg, dp_fg = ops.encode_tc(rays11, fg_z_vals, self.mode)
            raw_fg, st_fg = ops.mlp_forward(self.mode, self.fg, self._packed_fg, x_fg, dp_fg, B * Sf, Sf, training)
            x_bg, dp_bg = ops.pack_xrows(self.mode, self.in_bg, x_bg.view(B * Sb, -1), Sb)
            raw_bg, st_bg = ops.mlp_forward_g(self.mode, self.in_bg, self.bg, self._packed_bg, x_bg, dp_bg, B * Sb, Sb, training)
        rgb, fw, bw, aux = ops.pp_composite_forward(raw_fg.view(B, Sf, 4), fg_z_vals, fg_z_max, raw_bg.view(B, Sb, 4), bg_flip,
                                                    ray_d)
        self._saved = (B, Sf, Sb, x_fg, dp_fg, st_fg, raw_fg, x_bg, dp_bg, st_bg, raw_bg, fg_z_vals, fg_z_max, bg_flip, ray_d)
        return {"rgb": rgb, "fg_weights": fw, "bg_weights": bw, "fg_rgb": aux[:, 0:3], "fg_depth": aux[:, 3],
                "bg_rgb": aux[:, 4:7], "bg_depth": aux[:, 7], "bg_lambda": aux[:, 8]}

    @torch.no_grad()
    def backward(self, g_rgb):
        """Accumulates d(loss)/d(params) into grad_fg / grad_bg for a loss whose gradient w.r.t. ret['rgb'] is g_rgb."""
        if self._saved is None:
            raise FlnerfError("NerfNet.backward() without a forward()")
        B, Sf, Sb, x_fg, dp_fg, st_fg, raw_fg, x_bg, dp_bg, st_bg, raw_bg, fg_z, fg_far, bg_flip, ray_d = self._saved
        d_fg, d_bg = ops.pp_composite_backward(raw_fg.view(B, Sf, 4), fg_z, fg_far, raw_bg.view(B, Sb, 4), bg_flip, ray_d, g_rgb)
        if self.mode == ops.MODE_FP32:
            ops.mlp_backward(ops.MODE_FP32, self.fg, None, x_fg, None, st_fg, d_fg.view(B * Sf, 4), self.grad_fg, B * Sf, Sf)
            ops.mlp_fp32_backward_g(self.in_bg, self.in_views, self.bg, x_bg, st_bg, d_bg.view(B * Sb, 4), self.grad_bg, B * Sb)
        else:
            ops.mlp_backward(self.mode, self.fg, self._packed_fg, x_fg, dp_fg, st_fg, d_fg.view(B * Sf, 4), self.grad_fg, B * Sf, Sf)
            ops.mlp_backward_g(self.mode, self.in_bg, self.bg, self._packed_bg, x_bg, dp_bg, st_bg, d_bg.view(B * Sb, 4),
                               self.grad_bg, B * Sb, Sb)
        self._saved = None
        return self.grad_fg, self.grad_bg


class NerfNetFP32(NerfNet):
    """The fp32 parity path under its round-1 name."""

    def __init__(self, flat_fg, flat_bg, in_pts_bg: int = 84, in_views: int = 27):
        super().__init__(flat_fg, flat_bg, in_pts_bg, in_views, "fp32")


class FlatAdam:
    """torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8) over flat parameter buffers (one fused kernel per buffer)."""

    def __init__(self, flats, lr=5e-4, betas=(0.9, 0.999), eps=1e-8):
        self.flats, self.lr, self.betas, self.eps, self.t = list(flats), float(lr), betas, float(eps), 0
        self.m = [torch.zeros_like(f) for f in self.flats]
        self.v = [torch.zeros_like(f) for f in self.flats]

    def step(self, grads):
        self.t += 1
        for f, m, v, g in zip(self.flats, self.m, self.v, grads):
            ops.adam_step(f, m, v, g, self.lr, self.betas[0], self.betas[1], self.eps, self.t)


@torch.no_grad()
def cascade_train_step(nets, adams, ray_o, ray_d, rgb_gt, samples=(64, 128), t_fg=None, t_bg=None, u_fg=None, u_bg=None,
                       seed=0, offset=0):
    """One iteration of the nerf++ training loop for one batch (ddp_train_nerf.py:346-404, auto-exposure off): two cascade
    levels, each a NerfNetFP32 with its own FlatAdam; level 1 resamples the fg and bg depths from level 0's weights
    (bg_weights un-flipped, as the fork does).  Uniforms are the caller's (parity) or Philox(seed, offset).  Returns the two
    losses as DEVICE tensors and the last rgb."""
    B = ray_o.shape[0]
    losses, ret, fg_z, bg_z, fg_far = [], None, None, None, None
    for m, net in enumerate(nets):
        if m == 0:
            fg_far, fg_z, bg_z = ops.pp_depths0(ray_o, ray_d, samples[0], True, t_fg, t_bg, seed, offset)
        else:
            fg_z, _ = ops.pp_sample_pdf_merge(fg_z, ret["fg_weights"], samples[1], False, u_fg, seed + 1, offset)
            bg_z, _ = ops.pp_sample_pdf_merge(bg_z, ret["bg_weights"], samples[1], False, u_bg, seed + 2, offset)
        net.grad_fg.zero_()
        net.grad_bg.zero_()
        ret = net.forward(ray_o, ray_d, fg_far, fg_z, bg_z)
        loss, d_rgb, _ = ops.mse_leafmax(ret["rgb"], None, rgb_gt, B)
        net.backward(d_rgb)
        adams[m].step([net.grad_fg, net.grad_bg])
        losses.append(loss[0:1])
    return torch.cat(losses), ret["rgb"]
