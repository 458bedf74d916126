"""Experiment: gradient accuracy of the split-precision mode when the weight-gradient GEMM carries 1, 2 or 3 of the terms
dYhi^T Xhi + dYhi^T Xlo + dYlo^T Xhi ($FLNERF_X3_WGRAD_PASSES, read once per process).  Prints rel-L2 vs the CPU oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("fast-learning-nerf_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import nerf_oracle as O
from test_gpu_x3 import lego_rays, make_net, oracle_step_chunked, rel_l2
from flnerf_b200.engine import FusedAdam, Trainer

H, W, K, ro, rd, tgt = lego_rays(2048, seed=2)
nc, nf = make_net(41, "bf16x3"), make_net(42, "bf16x3")
opt = FusedAdam(list(nc.parameters()) + list(nf.parameters()), [nc, nf], lr=5e-4)
tr = Trainer(nc, nf, opt, H, W, K, 2.0, 6.0, 64, 128, white_bkgd=True, perturb=0.0)
loss = tr.step(ro.cuda(), rd.cuda(), tgt.cuda())
torch.cuda.synchronize()
g = tr.bucket.cpu().clone()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(5):
    tr.step(ro.cuda(), rd.cuda(), tgt.cuda())
t1.record(); torch.cuda.synchronize()
pc, pf = O.init_params(41), O.init_params(42)
res = oracle_step_chunked(O.pack_rays(H, W, K, ro, rd, 2.0, 6.0, ndc=False), tgt, pc, pf, 64, 128, white_bkgd=True)
g_or = torch.cat([x.reshape(-1) for x in res["grads"]])
half = g.numel() // 2
print("passes=%s: grad rel-L2 coarse %.3e fine %.3e | loss rel %.2e | %.2f ms/step (2048 rays)" % (
    os.environ.get("FLNERF_X3_WGRAD_PASSES", "3"), rel_l2(g[:half], g_or[:half]), rel_l2(g[half:], g_or[half:]),
    abs(float(loss.sum()) - res["loss"]) / res["loss"], t0.elapsed_time(t1) / 5))
