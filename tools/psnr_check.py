"""PSNR at equal iterations: the bf16 tensor-core path vs the fp32 parity path (== the reference arithmetic, see
tests/test_gpu_mlp.py) on the synthetic blob scene; same seed => same init, same ray batches, same jitter.
Reports PSNR on training views (fit) and held-out views.  Prints one JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fast-learning-nerf_b200"))
import numpy as np
import torch
import model, tree, render as R, run_nerf, run_nerf_helpers as H
from flnerf_b200 import synthetic
from flnerf_b200.engine import FusedAdam, Trainer

RES, VIEWS = int(os.environ.get("PC_RES", 100)), int(os.environ.get("PC_VIEWS", 40))
ITERS, NRAND = int(os.environ.get("PC_ITERS", 2000)), int(os.environ.get("PC_NRAND", 1024))
SEEDS = [int(s) for s in os.environ.get("PC_SEEDS", "0,1,2").split(",")]
dev = torch.device("cuda")
K = synthetic.intrinsics(RES, RES, 0.5 * RES / np.tan(0.5 * 0.6911112070083618))
poses = synthetic.lego_like_poses(VIEWS)
test_poses = synthetic.lego_like_poses(4, phi=-25.0)
imgs = synthetic.render_scene(RES, RES, K, poses, n_samples=128)
test_imgs = synthetic.render_scene(RES, RES, K, test_poses, n_samples=128)
q = run_nerf.NetworkQuery(H.get_embedder(10)[0], H.get_embedder(4)[0], 65536)


def psnr_on(nc, nf, ps, gts):
    out = []
    with torch.no_grad():
        for c2w, gt in zip(ps, gts):
            rgb = R.render(RES, RES, K, chunk=32768, c2w=torch.as_tensor(c2w[:3, :4], device=dev), ndc=False, near=2.0, far=6.0,
                           use_viewdirs=True, network_query_fn=q, network_fn=nc, network_fine=nf, N_samples=64,
                           N_importance=128, white_bkgd=True, perturb=0.0)[0]
            out.append(float(-10 * torch.log10(torch.mean((rgb - gt) ** 2))))
    return float(np.mean(out))


res = {}
for seed in SEEDS:
    # "fp32p" = the fp32 run again with every initial weight perturbed by 1e-6 relative: its distance from "fp32" is the
    # noise floor of the comparison (training trajectories are chaotic), the yardstick for the bf16 - fp32 difference
    for prec in ("fp32", "bf16") + (("fp32p",) if os.environ.get("PC_NOISE_FLOOR", "1") == "1" else ()):
        torch.manual_seed(seed)
        nc = model.NeRF(8, 256, 63, 27, 5, [4], True, precision=prec[:4]).to(dev)
        nf = model.NeRF(8, 256, 63, 27, 5, [4], True, precision=prec[:4]).to(dev)
        if prec == "fp32p":
            with torch.no_grad():
                g = torch.Generator(device=dev).manual_seed(1234 + seed)
                for net in (nc, nf):
                    for prm in net.parameters():
                        prm.mul_(1 + 1e-6 * torch.randn(prm.shape, device=dev, generator=g))
                    if hasattr(net, "weights_version"):
                        net.weights_version += 1
        opt = FusedAdam(list(nc.parameters()) + list(nf.parameters()), [nc, nf], lr=5e-4)
        tr = Trainer(nc, nf, opt, RES, RES, K, 2.0, 6.0, 64, 128, white_bkgd=True, perturb=1.0, seed=seed)
        mgr = tree.QuadTreeManager(RES, RES, K, imgs, torch.as_tensor(poses[:, :3, :4]), mseThres=0.0, max_depth=2, max_level=5, seed=seed)
        it = 0
        while it < ITERS:
            n = mgr.emit_epoch()
            for first in range(0, n - NRAND, NRAND):
                tr.step_from_tree(mgr, first, NRAND)
                it += 1
                if it >= ITERS:
                    break
            mgr.refine(0.001)
        res["%s.seed%d" % (prec, seed)] = {"train": psnr_on(nc, nf, poses[::10], imgs[::10]), "test": psnr_on(nc, nf, test_poses, test_imgs)}
        print("partial", prec, seed, res["%s.seed%d" % (prec, seed)], file=sys.stderr, flush=True)
for k in ("train", "test"):
    d = [res["bf16.seed%d" % s][k] - res["fp32.seed%d" % s][k] for s in SEEDS]
    res["delta_%s_db" % k] = d
    res["mean_delta_%s_db" % k] = float(np.mean(d))
    if "fp32p.seed%d" % SEEDS[0] in res:
        f = [res["fp32p.seed%d" % s][k] - res["fp32.seed%d" % s][k] for s in SEEDS]
        res["noise_floor_%s_db" % k] = f
        res["mean_abs_noise_floor_%s_db" % k] = float(np.mean(np.abs(f)))
        res["mean_abs_delta_%s_db" % k] = float(np.mean(np.abs(d)))
res["config"] = dict(res=RES, views=VIEWS, iters=ITERS, n_rand=NRAND)
print(json.dumps(res))
