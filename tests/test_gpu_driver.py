"""GPU (-m gpu): the reference-facing driver end to end (run_nerf.py:156-548): centre-crop warm-up -> epochs with quadtree
refinement -> checkpoints and tree pickles -> resume -> --render_only; and interchange of those files with the UNMODIFIED
reference (its checkpoint + treeDivide pickle load here, ours load there)."""
import glob
import os
import pickle
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(ROOT, "fast-learning-nerf_b200", "configs", "synthetic_lego.txt")


@pytest.fixture()
def small_scene(monkeypatch):
    monkeypatch.setenv("FLNERF_SYN_RES", "48")
    monkeypatch.setenv("FLNERF_SYN_VIEWS", "4")


def test_train_checkpoint_resume_render_only(tmp_path, small_scene, capsys):
    import run_nerf
    import tree
    base = ["--config", CFG, "--basedir", str(tmp_path), "--expname", "drv", "--N_rand", "512", "--subdivide_every", "1",
            "--init_level", "2", "--precision", "bf16"]
    run_nerf.train(base + ["--n_epoch", "3"])
    out = capsys.readouterr().out
    d = tmp_path / "drv"
    assert "Center Cropping" in out and "Epoch 3" in out and "last epoch: use all rays to train." in out
    assert "After sudivide" in out                                       # the refinement of epoch 1 (epoch < n_epoch - 1)
    assert sorted(os.path.basename(p) for p in glob.glob(str(d / "*.tar"))) == ["001.tar", "002.tar", "003.tar"]
    assert os.path.isfile(d / "treeDivide_0003.pkl") and os.path.isfile(d / "args.txt") and os.path.isfile(d / "config.txt")
    ck = torch.load(d / "003.tar", weights_only=False)
    assert set(ck) == {"global_epoch", "global_iter", "network_fn_state_dict", "network_fine_state_dict", "optimizer_state_dict"}
    assert ck["global_epoch"] == 3 and ck["global_iter"] == 3 * 18 and len(ck["optimizer_state_dict"]["state"]) == 48
    assert all(k.startswith("module.") for k in ck["network_fn_state_dict"])
    trees = pickle.load(open(d / "treeDivide_0003.pkl", "rb"))
    assert len(trees) == 4 and type(trees[0]).__module__ == "tree"
    n_leaves = [len(tree.get_children(t.root)) for t in trees]
    assert all(4 <= n <= 16 for n in n_leaves) and max(n_leaves) > 4     # epoch 1's refinement split some 24x24 leaves
    # resume: picks up 003.tar + the pickled trees and trains epoch 4 only
    run_nerf.train(base + ["--n_epoch", "4"])
    out = capsys.readouterr().out
    assert "Reloading from" in out and "003.tar" in out and "treeDivide_0003.pkl" in out
    assert "Epoch 4" in out and "Epoch 3" not in out and "Center Cropping" not in out
    ck4 = torch.load(d / "004.tar", weights_only=False)
    assert ck4["global_epoch"] == 4 and float(ck4["optimizer_state_dict"]["state"][0]["step"]) > float(ck["optimizer_state_dict"]["state"][0]["step"])
    # render_only on the held-out poses: PSNR / SSIM of a network that has learnt something
    run_nerf.train(base + ["--n_epoch", "4", "--render_only", "--render_test"])
    out = capsys.readouterr().out
    res = glob.glob(str(d / "renderonly_test_*" / "results.txt"))
    assert "RENDER ONLY" in out and len(res) == 1
    psnr = float(open(res[0]).read().split("mean PSNR:")[1].split()[0])
    assert np.isfinite(psnr) and psnr > 3.0, psnr          # 72 iterations on 48x48 views: a sanity bound, not a quality claim


def test_files_interchange_with_the_unmodified_reference(tmp_path, small_scene):
    """A checkpoint + tree pickle WRITTEN by the reference's own objects (nn.DataParallel(NeRF), torch.optim.Adam,
    tree.QuadTreeManager) resume here; a checkpoint written here loads into the reference's modules."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shim
    if not ref_shim.available():
        pytest.skip("no reference sources (baseline/_ref is made by tools/install_reference.sh in the build container)")
    import run_nerf
    ns, rn = ref_shim.load_run_nerf()
    rn.device = torch.device("cuda")
    d = tmp_path / "ref"
    os.makedirs(d)
    ref_cfg = os.path.join(ref_shim.REF_NERF, "configs", "lego.txt")
    args = rn.config_parser().parse_args(["--config", ref_cfg, "--basedir", str(tmp_path), "--expname", "ref"])
    torch.manual_seed(0)
    kw, _, _, _, grad_vars, opt = rn.create_nerf(args)
    for p in grad_vars:                       # two stock Adam steps so that the optimiser state is populated
        p.grad = torch.randn_like(p) * 1e-3
    opt.step(); opt.step()
    torch.save({"global_epoch": 2, "global_iter": 77, "network_fn_state_dict": kw["network_fn"].state_dict(),
                "network_fine_state_dict": kw["network_fine"].state_dict(), "optimizer_state_dict": opt.state_dict()},
               d / "002.tar")                                             # run_nerf.py:532-539, by the reference's objects
    H = W = 48
    from flnerf_b200 import synthetic
    K = synthetic.intrinsics(H, W, 0.5 * W / np.tan(0.5 * 0.6911112070083618))
    poses = synthetic.lego_like_poses(4)
    imgs = np.random.RandomState(0).rand(4, H, W, 3).astype(np.float32)
    mgr = ns.tree.QuadTreeManager(H, W, K, torch.from_numpy(imgs), torch.from_numpy(poses[:, :3, :4]), mseThres=0.0, max_depth=2)
    o_, d_, rgb_ = mgr.gen_rays_v3_multiThread(down_scale=1, prob=False, last_epoch=False)
    pred = rgb_.clone(); pred[: pred.shape[0] // 2] += 0.5                # half of the rays are badly predicted -> splits
    mgr.adjust_tree_multiThread(rgb_, pred, thres=0.001)
    ref_leaves = [[(c.x0, c.y0, c.x1, c.y1) for c in ch] for ch in mgr.childrens]
    mods = {k: sys.modules.get(k) for k in ("tree", "image_process")}
    sys.modules["tree"], sys.modules["image_process"] = ns._tree, ns._image_process      # pickle by the reference's qualified names
    try:
        with open(d / "treeDivide_0002.pkl", "wb") as f:
            pickle.dump(mgr.quadTrees, f)
    finally:
        for k, v in mods.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
    # --- resume here from the reference's files, train one more epoch
    import tree
    run_nerf.train(["--config", CFG, "--basedir", str(tmp_path), "--expname", "ref", "--N_rand", "512", "--n_epoch", "3",
                    "--subdivide_every", "1", "--init_level", "2", "--precision", "bf16x3"])
    ck = torch.load(d / "003.tar", weights_only=False)
    assert ck["global_epoch"] == 3 and ck["global_iter"] == 77 + 18
    ours_trees = pickle.load(open(d / "treeDivide_0003.pkl", "rb"))
    got = [[(c.x0, c.y0, c.x1, c.y1) for c in tree.get_children(t.root)] for t in ours_trees]
    assert got == ref_leaves                                             # epoch 3 = last epoch: no further refinement
    # --- and back: our checkpoint loads into the reference's modules and stock Adam
    kw["network_fn"].load_state_dict(ck["network_fn_state_dict"])
    kw["network_fine"].load_state_dict(ck["network_fine_state_dict"])
    opt.load_state_dict(ck["optimizer_state_dict"])
    assert float(opt.state_dict()["state"][0]["step"]) == 2 + 18
