"""One steady-state training step per precision (bf16, bf16x3) at BASELINE configs[1] size, bracketed by
cudaProfilerStart/Stop so that `ncu --profile-from-start off` sees exactly those launches (tools/gpu_profile_r02.sh)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-learning-nerf_b200"))
import torch
import model, tree
from flnerf_b200 import synthetic
from flnerf_b200.engine import FusedAdam, Trainer

H = W = 800
K = synthetic.intrinsics(H, W, 1111.111)
poses = synthetic.lego_like_poses(4)
imgs = synthetic.render_scene(H, W, K, poses, n_samples=32)
mgr = tree.QuadTreeManager(H, W, K, imgs, torch.as_tensor(poses[:, :3, :4]), mseThres=0.0, max_depth=2, max_level=7, seed=0)
mgr.emit_epoch()
modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["bf16", "bf16x3"]
for prec in modes:
    torch.manual_seed(0)
    nc = model.NeRF(8, 256, 63, 27, 5, [4], True, precision=prec).cuda()
    nf = model.NeRF(8, 256, 63, 27, 5, [4], True, precision=prec).cuda()
    opt = FusedAdam(list(nc.parameters()) + list(nf.parameters()), [nc, nf], lr=5e-4)
    tr = Trainer(nc, nf, opt, H, W, K, 2.0, 6.0, 64, 128, white_bkgd=True, perturb=1.0, graph=False)
    first = 0
    for _ in range(3):
        tr.step_from_tree(mgr, first, 4096); first += 4096
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    tr.step_from_tree(mgr, first, 4096)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
