"""PSNR at equal iterations: the bf16 tensor-core path vs the fp32 parity path (== the reference arithmetic, see
tests/test_gpu_mlp.py) on the synthetic blob scene, same seeds / same ray batches.  Prints one JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-learning-nerf_b200"))
import numpy as np
import torch
import model, tree, render as R, run_nerf, run_nerf_helpers as H
from flnerf_b200 import synthetic
from flnerf_b200.engine import FusedAdam, Trainer

RES, VIEWS, ITERS, NRAND = int(os.environ.get("PC_RES", 160)), 24, int(os.environ.get("PC_ITERS", 1500)), 2048
dev = torch.device("cuda")
K = synthetic.intrinsics(RES, RES, 0.5 * RES / np.tan(0.5 * 0.6911112070083618))
poses = synthetic.lego_like_poses(VIEWS)
test_poses = synthetic.lego_like_poses(4, phi=-20.0)
imgs = synthetic.render_scene(RES, RES, K, poses, n_samples=128)
test_imgs = synthetic.render_scene(RES, RES, K, test_poses, n_samples=128)
out = {}
for seed in (0, 1):
    for prec in ("fp32", "bf16"):
        torch.manual_seed(seed)
        nc = model.NeRF(8, 256, 63, 27, 5, [4], True, precision=prec).to(dev)
        nf = model.NeRF(8, 256, 63, 27, 5, [4], True, precision=prec).to(dev)
        opt = FusedAdam(list(nc.parameters()) + list(nf.parameters()), [nc, nf], lr=5e-4)
        tr = Trainer(nc, nf, opt, RES, RES, K, 2.0, 6.0, 64, 128, white_bkgd=True, perturb=1.0, seed=seed)
        mgr = tree.QuadTreeManager(RES, RES, K, imgs, torch.as_tensor(poses[:, :3, :4]), mseThres=0.0, max_depth=2, max_level=5, seed=seed)
        it = 0
        while it < ITERS:
            n = mgr.emit_epoch()
            for first in range(0, n - NRAND, NRAND):
                tr.step_from_tree(mgr, first, NRAND)
                it += 1
                for g in opt.param_groups:
                    g["lr"] = 5e-4 * 0.1 ** (it / 500000)
                if it >= ITERS:
                    break
            mgr.refine(0.001)
        q = run_nerf.NetworkQuery(H.get_embedder(10)[0], H.get_embedder(4)[0], 65536)
        ps = []
        with torch.no_grad():
            for i, c2w in enumerate(test_poses):
                rgb, _, _, _ = R.render(RES, RES, K, chunk=32768, c2w=torch.as_tensor(c2w[:3, :4], device=dev), ndc=False, near=2.0,
                                        far=6.0, use_viewdirs=True, network_query_fn=q, network_fn=nc, network_fine=nf,
                                        N_samples=64, N_importance=128, white_bkgd=True, perturb=0.0)
                ps.append(float(-10 * torch.log10(torch.mean((rgb - test_imgs[i]) ** 2))))
        out["%s.seed%d" % (prec, seed)] = float(np.mean(ps))
d = [out["bf16.seed%d" % s] - out["fp32.seed%d" % s] for s in (0, 1)]
out["delta_db"] = d
out["config"] = dict(res=RES, views=VIEWS, iters=ITERS, n_rand=NRAND, leaves=int(mgr.counts.sum()))
print(json.dumps(out))
