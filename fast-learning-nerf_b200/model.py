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
        net._mask_grads()
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
            self.output_ch = output_ch
        self.precision = precision or DEFAULT_PRECISION
        self._flat = self._flat_grad = self._packed = self._proxy = None
        self._packed_key = None
        self.weights_version = 0      # bumped by the fused optimiser (kernels do not touch tensor._version)
        self._keep, self._frozen_adapter = None, False    # flat 0/1 mask of the trainable entries (inner net of a use_viewdirs=False model)
        if not use_viewdirs:
            self._build_inner()

    # ------------------------------------------------------------------ use_viewdirs=False (model.py:55-63)
    # output_linear (W -> 4 or 5) replaces the alpha / feature / views / rgb heads.  It needs no kernel of its own: rows 0..2 of
    # output_linear are rows 0..2 of an inner net's feature_linear, row 3 is its alpha_linear, and a FROZEN +-1 adapter in
    # views_linears.0 / rgb_linear (units 2k, 2k+1 = relu(+f_k), relu(-f_k); rgb_k = their difference) turns the ReLU between
    # them into the identity -- for the gradients as well.  The registered parameters (and state_dict) are the reference's;
    # the inner use_viewdirs=True net (`kernel_net`) is what the kernels, the fused optimiser and the Trainer work on, with
    # the adapter entries masked out of every gradient (`_keep`).
    def _build_inner(self):
        if self.W != 256 or self.D != 8 or self.input_ch != 63 or self.skips != [4]:
            object.__setattr__(self, "_inner", None)          # unsupported architecture: forward() will say so
            return
        with torch.random.fork_rng(devices=[]):               # the reference's constructor draws nothing for these
            inner = NeRF(D=self.D, W=self.W, input_ch=self.input_ch, input_ch_views=27, output_ch=4, skips=self.skips,
                         use_viewdirs=True, precision=self.precision)
        object.__setattr__(self, "_inner", inner)             # NOT a registered submodule: parameters() stay the reference's
        self._pub_versions = None
        self._push_to_inner()
        # (a hook, not an override: a parent module -- the "module." holder -- loads its children without calling their
        #  load_state_dict)
        self.register_load_state_dict_post_hook(lambda mod, incompatible: mod._push_to_inner())

    @property
    def kernel_net(self):
        """The module whose flat buffers the kernels / FusedAdam / Trainer use (self, unless use_viewdirs=False).  Train it with
        engine.FusedAdam / engine.Trainer: the registered (public) parameters of a use_viewdirs=False model receive no .grad."""
        return self if self.use_viewdirs else self._sync_inner()

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float(): the inner net holds the trained values -- bring the public copy up to date first; the
        # inner net follows lazily (_sync_inner sees the device change)
        if not self.use_viewdirs and getattr(self, "_inner", None) is not None and getattr(self, "_pub_versions", None) is not None:
            self._pull_from_inner()
        return super()._apply(fn, *args, **kwargs)

    def _public(self):
        return list(self.pts_linears.parameters()) + [self.output_linear.weight, self.output_linear.bias]

    @torch.no_grad()
    def _push_to_inner(self):
        """public parameters -> inner net (+ the frozen adapter)."""
        inn = self._inner
        dev = self.output_linear.weight.device
        if next(inn.parameters()).device != dev:
            inn.to(dev)
        for a, b in zip(self.pts_linears, inn.pts_linears):
            b.weight.copy_(a.weight); b.bias.copy_(a.bias)
        W, bo = self.output_linear.weight, self.output_linear.bias
        inn.feature_linear.weight.zero_(); inn.feature_linear.bias.zero_()
        inn.feature_linear.weight[0:3].copy_(W[0:3]); inn.feature_linear.bias[0:3].copy_(bo[0:3])
        inn.alpha_linear.weight.copy_(W[3:4]); inn.alpha_linear.bias.copy_(bo[3:4])
        vw = inn.views_linears[0]
        vw.weight.zero_(); vw.bias.zero_()
        inn.rgb_linear.weight.zero_(); inn.rgb_linear.bias.zero_()
        for k in range(3):
            vw.weight[2 * k, k] = 1.0; vw.weight[2 * k + 1, k] = -1.0
            inn.rgb_linear.weight[k, 2 * k] = 1.0; inn.rgb_linear.weight[k, 2 * k + 1] = -1.0
        inn.weights_version += 1
        self._pub_versions = [p._version for p in self._public()]

    @torch.no_grad()
    def _pull_from_inner(self):
        """inner net -> public parameters (what a checkpoint stores)."""
        inn = self._inner
        for a, b in zip(self.pts_linears, inn.pts_linears):
            a.weight.copy_(b.weight); a.bias.copy_(b.bias)
        self.output_linear.weight[0:3].copy_(inn.feature_linear.weight[0:3]); self.output_linear.bias[0:3].copy_(inn.feature_linear.bias[0:3])
        self.output_linear.weight[3:4].copy_(inn.alpha_linear.weight); self.output_linear.bias[3:4].copy_(inn.alpha_linear.bias)
        self._pub_versions = [p._version for p in self._public()]

    def _sync_inner(self):
        """Called before the inner net is used: picks up public parameters somebody changed in place or moved."""
        if self._inner is None:
            raise FlnerfError("flnerf NeRF kernels implement D=8, W=256, input_ch=63, skips=[4]; got another architecture")
        inn = self._inner
        inn.precision = self.precision
        if next(inn.parameters()).device != self.output_linear.weight.device or \
                self._pub_versions != [p._version for p in self._public()]:
            self._push_to_inner()
        inn._frozen_adapter = True
        return inn

    def state_dict(self, *args, **kwargs):
        if not self.use_viewdirs and self._inner is not None:
            self._pull_from_inner()
        return super().state_dict(*args, **kwargs)

    def _mask_grads(self):
        """Zeroes the gradient entries of frozen parameters (the adapter of a use_viewdirs=False model; no-op otherwise)."""
        if not self._frozen_adapter:
            return
        g = self._flat_grad
        if self._keep is None or self._keep.device != g.device:
            keep = torch.zeros(g.numel(), dtype=torch.float32, device=g.device)
            off = 0
            for name, p in self.named_parameters():          # flat order == parameters() order
                n = p.numel()
                if name.startswith("pts_linears.") or name.startswith("alpha_linear."):
                    keep[off:off + n] = 1.0
                elif name == "feature_linear.weight":
                    keep[off:off + 3 * self.W] = 1.0          # rows 0..2 = output_linear rows 0..2
                elif name == "feature_linear.bias":
                    keep[off:off + 3] = 1.0
                off += n
            self._keep = keep
        g.mul_(self._keep)

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
        """x[..., input_ch + input_ch_views] -> [..., 4] (model.py:38-63); use_viewdirs=False: [..., output_ch]."""
        if not self.use_viewdirs:
            inn = self._sync_inner()
            pts = x[..., :self.input_ch]
            raw = inn(torch.cat([pts, pts.new_zeros(list(pts.shape[:-1]) + [27])], -1))
            if self.output_ch > 4:      # rows >= 4 of output_linear never reach raw2outputs (render.py:162-171)
                raw = torch.cat([raw, raw.new_zeros(list(raw.shape[:-1]) + [self.output_ch - 4])], -1)
            return raw
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
        if not self.use_viewdirs:       # the view-direction columns of rays11 meet zero weights
            return self._sync_inner().query_rays(rays11, z)
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
        if not self.use_viewdirs:
            return self._sync_inner().query_tiles(tiles, dirpe, B, S)
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
