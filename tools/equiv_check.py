"""Diagnostic: run-to-run determinism and row-permutation equivariance of the bf16 forward (where do rows differ?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-learning-nerf_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import model
from flnerf_b200 import ops
torch.manual_seed(0)
B, Nc = int(os.environ.get("EQ_RAYS", 4096)), 64
rays = torch.cat([torch.randn(B, 3) * 0.3 + torch.tensor([0., 0., 4.]), -torch.nn.functional.normalize(torch.randn(B, 3) * 0.2 + torch.tensor([0., 0., 1.]), dim=-1),
                  2 * torch.ones(B, 1), 6 * torch.ones(B, 1), torch.nn.functional.normalize(torch.randn(B, 3), dim=-1)], -1).cuda()
z = ops.coarse_depths(rays, Nc, True, False, None, 1, 0)
net = model.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision="bf16").cuda()
with torch.no_grad():
    a = net.query_rays(rays, z)
    b = net.query_rays(rays, z)
    d = (a - b).abs().reshape(-1, 4)
    print("same input twice: max diff", float(d.max()), "rows differing", int((d.max(-1)[0] > 0).sum()), "of", d.shape[0])
    perm = torch.randperm(B, device="cuda")
    c = net.query_rays(rays[perm].contiguous(), z[perm].contiguous())
    d = (c - a[perm]).abs().reshape(-1, 4)
    bad = (d.max(-1)[0] > 0).nonzero().flatten()
    print("permuted: max diff", float(d.max()), "rows differing", bad.numel(), "per channel max", d.max(0)[0].tolist())
    if bad.numel():
        rows = bad.cpu()
        print("  first rows", rows[:16].tolist())
        print("  row%128 histogram (16 bins)", torch.histc((rows % 128).float(), 16, 0, 128).tolist())
        print("  (row//128)%4 histogram", torch.histc(((rows // 128) % 4).float(), 4, 0, 4).tolist())
        print("  per-channel differing counts", (d[bad] > 0).sum(0).tolist())
