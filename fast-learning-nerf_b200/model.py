"""Drop-in for nerf-ours/model.py: the ``NeRF`` module with the reference constructor, parameter
names/shapes (state_dict compatible, model.py:20-34) and ``forward(x[n, 63+27]) -> [n, 4]``.

The arithmetic is NOT torch: forward and backward run in libflnerf.so (csrc/mlp_simt.cu for the fp32
parity mode, csrc/mlp_tc.cu -- tcgen05 -- for the bf16 throughput mode and the split-precision bf16x3
tensor-core parity mode).  All 24 parameter tensors are
views into one flat fp32 buffer in ``parameters()`` order, their ``.grad`` are views into one flat
gradient bucket (what the fused Adam and the single NCCL all-reduce operate on).
"""
import os

import numpy as np
import torch
from torch import nn

from flnerf_b200 import ops
from flnerf_b200.lib import FlnerfError, MLP_PARAMS

DEFAULT_PRECISION = os.environ.get("FLNERF_PRECISION", "bf16")
_MODES = {"fp32": ops.MODE_FP32, "bf16": ops.MODE_BF16, "bf16x3": ops.MODE_BF16X3}


class _NerfFn(torch.autograd.Function):
    """The MLP as one autograd node.  ``proxy`` is a 1-element tensor that requires grad so that autograd
    schedules this node; the real parameter gradients are accumulated by the kernels straight into the
    module's flat gradient bucket (no input gradient exists: render.py:281)."""

    @staticmethod
    def forward(ctx, proxy, net, x, dirpe, n, S):
        training = bool(ctx.needs_input_grad[0])
        flat, packed = net._weights()
        raw, stash = ops.mlp_forward(net.mode, flat, packed, x, dirpe, n, S, training)
        ctx.net, ctx.x, ctx.dirpe, ctx.stash, ctx.n, ctx.S = net, x, dirpe, stash, n, S
        return raw

    @staticmethod
    def backward(ctx, draw):
        net = ctx.net
        flat, packed = net._weights()
        ops.mlp_backward(net.mode, flat, packed, ctx.x, ctx.dirpe, ctx.stash, draw.contiguous(), net._grad_bucket(),
                         ctx.n, ctx.S)
        ctx.stash = ctx.x = None
        return torch.zeros_like(net._proxy), None, None, None, None, None


class NeRF(nn.Module):
    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4], use_viewdirs=False,
                 precision=None):
        super().__init__()
        self.D, self.W, self.input_ch, self.input_ch_views = D, W, input_ch, input_ch_views
        self.skips, self.use_viewdirs = list(skips), use_viewdirs
        widths_in = [input_ch] + [W + input_ch if (i in self.skips) else W for i in range(D - 1)]
        self.pts_linears = nn.ModuleList([nn.Linear(k, W) for k in widths_in])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        if use_viewdirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)
        self.precision = precision or DEFAULT_PRECISION
        self._flat = self._flat_grad = self._packed = self._proxy = None
        self._packed_key = None
        self.weights_version = 0      # bumped by the fused optimiser (kernels do not touch tensor._version)

    # ------------------------------------------------------------------ flat storage
    @property
    def mode(self):
        return _MODES[self.precision]

    def _supported(self):
        return (self.D == 8 and self.W == 256 and self.input_ch == 63 and self.input_ch_views == 27
                and self.skips == [4] and self.use_viewdirs)

    def _ordered_params(self):
        return list(self.parameters())

    def _ensure_flat(self):
        ps = self._ordered_params()
        f = self._flat
        if f is not None and ps[0].data_ptr() == f.data_ptr() and ps[-1].data_ptr() == f.data_ptr() + 4 * (MLP_PARAMS - 3) \
                and ps[0].device == f.device:
            return
        if not self._supported():
            raise FlnerfError("flnerf NeRF kernels implement D=8, W=256, input_ch=63, input_ch_views=27, skips=[4], "
                              "use_viewdirs=True (the lego/fern configs); got another architecture")
        dev = ps[0].device
        if dev.type != "cuda":
            raise FlnerfError("NeRF parameters are on %s; the flnerf kernels need a CUDA (sm_100a) device" % dev)
        flat = torch.empty(MLP_PARAMS, dtype=torch.float32, device=dev)
        grad = torch.zeros(MLP_PARAMS, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in ps:
                n = p.numel()
                flat[off:off + n].copy_(p.data.reshape(-1))
                p.data = flat[off:off + n].view(p.shape)
                p.grad = grad[off:off + n].view(p.shape)
                off += n
        assert off == MLP_PARAMS
        self._flat, self._flat_grad = flat, grad
        self._proxy = torch.zeros(1, dtype=torch.float32, device=dev, requires_grad=True)
        self._packed, self._packed_key = None, None

    def _grad_bucket(self):
        """Flat gradient buffer; re-attaches the .grad views if an optimiser set them to None."""
        self._ensure_flat()
        ps = self._ordered_params()
        if any(p.grad is None for p in ps):
            self._flat_grad.zero_()
            off = 0
            for p in ps:
                n = p.numel()
                p.grad = self._flat_grad[off:off + n].view(p.shape)
                off += n
        return self._flat_grad

    def flat_parameters(self):
        self._ensure_flat()
        return self._flat

    def _weights(self):
        self._ensure_flat()
        if self.mode != ops.MODE_FP32:
            key = (self._flat._version, self.weights_version, sum(p._version for p in self.parameters()))
            if self._packed is None or key != self._packed_key:
                self._packed = ops.mlp_pack_weights(self._flat, self._packed)
                self._packed_key = key
        return self._flat, self._packed

    # ------------------------------------------------------------------ reference API
    def forward(self, x):
        """x[..., input_ch + input_ch_views] -> [..., 4] (model.py:38-63)."""
        self._ensure_flat()
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1]).float().contiguous()
        n = x2.shape[0]
        proxy = self._proxy if (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())) \
            else self._proxy.detach()
        if self.mode == ops.MODE_FP32:
            raw = _NerfFn.apply(proxy, self, x2, None, n, 1)
        else:
            tiles, dirpe = ops.pack_x90(x2, self.mode)
            raw = _NerfFn.apply(proxy, self, tiles, dirpe, n, 1)
        return raw.reshape(*lead, 4)

    def query_rays(self, rays11, z):
        """Fused run_network (run_nerf.py:50-64): sample points o+d*z, PE(63)+PE(27), MLP -> raw[B,S,4]."""
        self._ensure_flat()
        B, S = z.shape
        proxy = self._proxy if (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())) \
            else self._proxy.detach()
        if self.mode == ops.MODE_FP32:
            x = ops.encode_f32(rays11, z)
            raw = _NerfFn.apply(proxy, self, x, None, B * S, S)
        else:
            tiles, dirpe = ops.encode_tc(rays11, z, self.mode)
            raw = _NerfFn.apply(proxy, self, tiles, dirpe, B * S, S)
        return raw.reshape(B, S, 4)

    def query_tiles(self, tiles, dirpe, B, S):
        """The MLP on input tiles some other kernel already produced (the eval path's fused frame encoder)."""
        self._ensure_flat()
        proxy = self._proxy if (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())) \
            else self._proxy.detach()
        return _NerfFn.apply(proxy, self, tiles, dirpe, B * S, S).reshape(B, S, 4)

    def load_weights_from_keras(self, weights):
        """model.py:65-92: weights = [W0,b0,...] in Keras order (kernels stored [in,out])."""
        assert self.use_viewdirs, "Not implemented if use_viewdirs=False"
        def put(lin, k):
            lin.weight.data.copy_(torch.from_numpy(np.transpose(weights[k])))
            lin.bias.data.copy_(torch.from_numpy(np.transpose(weights[k + 1])))
        for i in range(self.D):
            put(self.pts_linears[i], 2 * i)
        put(self.feature_linear, 2 * self.D)
        put(self.views_linears[0], 2 * self.D + 2)
        put(self.rgb_linear, 2 * self.D + 4)
        put(self.alpha_linear, 2 * self.D + 6)
