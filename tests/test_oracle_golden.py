"""CPU: the oracle restatement (oracle/nerf_oracle.py) against the committed golden vectors, which were produced
by executing the unmodified reference (oracle/make_golden.py).  Tolerances are the fp32 round-off of the
reference itself (~1e-4 max-rel on the fine output); integer / index work is exact."""
import numpy as np
import torch

import nerf_oracle as O


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_rays_and_ndc(golden):
    g = golden("rays")
    o, d = O.camera_rays(int(g["H"]), int(g["W"]), g["K"], T(g["c2w"]))
    assert torch.equal(o, T(g["rays_o"])) and torch.equal(d, T(g["rays_d"]))
    np.testing.assert_allclose(d.numpy(), g["rays_d_np"], rtol=0, atol=1e-6)      # get_rays_np agrees with get_rays
    no, nd = O.ndc_warp(int(g["Hf"]), int(g["Wf"]), g["Kf"][0][0], 1.0, T(g["fern_o"]).reshape(-1, 3), T(g["fern_d"]).reshape(-1, 3))
    assert torch.equal(no, T(g["ndc_o"])) and torch.equal(nd, T(g["ndc_d"]))


def test_posenc(golden):
    g = golden("posenc")
    assert torch.equal(O.posenc(T(g["pts"]), 10), T(g["pe_pts"]))
    assert torch.equal(O.posenc(T(g["dirs"]), 4), T(g["pe_dirs"]))
    assert g["pe_pts"].shape[1] == 63 and g["pe_dirs"].shape[1] == 27


def test_mlp_forward_and_grads(golden):
    g = golden("mlp")
    p = {k: v.clone().requires_grad_(True) for k, v in O.init_params(int(g["seed"])).items()}
    assert sum(v.numel() for v in p.values()) == 595844
    y = O.mlp_forward(p, T(g["x"]))
    np.testing.assert_allclose(y.detach().numpy(), g["y"], rtol=0, atol=2e-6)
    (y * T(g["gout"])).sum().backward()
    for k in g.files:
        if k.startswith("grad."):
            np.testing.assert_allclose(p[k[5:]].grad.numpy(), g[k], rtol=1e-4, atol=3e-6)
        elif k.startswith("gradrow."):
            np.testing.assert_allclose(p[k[8:]].grad[:4].numpy(), g[k], rtol=1e-4, atol=3e-6)
        elif k.startswith("gradnorm."):
            np.testing.assert_allclose(float(p[k[9:]].grad.double().norm()), float(g[k]), rtol=1e-5)


def test_composite(golden):
    g = golden("composite")
    raw, z, rd = T(g["raw"]), T(g["z"]), T(g["rays_d"])
    for wb in (0, 1):
        out = O.composite(raw, z, rd, None, bool(wb))
        for a, n in zip(out, ["rgb", "disp", "acc", "w", "depth"]):
            np.testing.assert_array_equal(np.nan_to_num(a.numpy(), nan=-7), np.nan_to_num(g[f"{n}_wb{wb}"], nan=-7))
    assert np.isnan(g["disp_wb0"][0])            # fully transparent ray: acc == 0 -> disp NaN (reference behaviour)
    rg = raw.clone().requires_grad_(True)
    (O.composite(rg, z, rd, None, True)[0] * T(g["g_rgb"])).sum().backward()
    np.testing.assert_allclose(rg.grad.numpy(), g["draw_wb1"], rtol=1e-5, atol=1e-7)


def test_sample_pdf_and_merge(golden):
    g = golden("sample_pdf")
    z, w = T(g["z"]), T(g["weights"])
    m, zs = O.fine_depths(z, w, 128, None)
    assert torch.equal(zs, T(g["zs_det"])) and torch.equal(m, T(g["merged_det"]))
    m, zs = O.fine_depths(z, w, 128, T(g["u"]))
    assert torch.equal(zs, T(g["zs_u"])) and torch.equal(m, T(g["merged_u"]))
    assert bool((m[:, 1:] >= m[:, :-1]).all())


def test_render_rays_and_loss(golden):
    g = golden("render_rays")
    pc, pf = O.init_params(int(g["seed_c"])), O.init_params(int(g["seed_f"]))
    rays = O.pack_rays(int(g["H"]), int(g["W"]), g["K"], T(g["rays_o"]), T(g["rays_d"]), float(g["near"]), float(g["far"]), ndc=False)
    tgt = T(g["target"])
    with torch.no_grad():
        out = O.render_rays(rays, pc, pf, 64, 128, white_bkgd=True)
    np.testing.assert_allclose(out["rgb0"].numpy(), g["det.rgb0"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["rgb_map"].numpy(), g["det.rgb"], rtol=0, atol=5e-5)
    np.testing.assert_allclose(out["z_std"].numpy(), g["det.z_std"], rtol=1e-4, atol=1e-5)
    loss = O.mse(out["rgb_map"], tgt) + O.mse(out["rgb0"], tgt)
    np.testing.assert_allclose(float(loss), float(g["det.loss"]), rtol=1e-5)
    # jittered path with the reference's pytest hook (numpy seed 0 uniforms)
    B = rays.shape[0]
    np.random.seed(0); tr = torch.Tensor(np.random.rand(B, 64))
    np.random.seed(0); uu = torch.Tensor(np.random.rand(B, 128))
    with torch.no_grad():
        out = O.render_rays(rays, pc, pf, 64, 128, white_bkgd=True, t_rand=tr, u=uu, det_fine=False)
    np.testing.assert_allclose(out["rgb0"].numpy(), g["jit.rgb0"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["rgb_map"].numpy(), g["jit.rgb"], rtol=0, atol=5e-5)


def test_train_step_gradients(golden):
    g = golden("render_rays")
    pc, pf = O.init_params(int(g["seed_c"])), O.init_params(int(g["seed_f"]))
    rays = O.pack_rays(int(g["H"]), int(g["W"]), g["K"], T(g["rays_o"]), T(g["rays_d"]), 2.0, 6.0, ndc=False)
    opt = O.AdamState(list(pc.values()) + list(pf.values()))
    before = pc["rgb_linear.bias"].clone()
    res = O.train_step(rays, T(g["target"]), pc, pf, opt, 64, 128, white_bkgd=True)
    names = list(pc.keys())
    for tag, off in (("c", 0), ("f", len(names))):
        for i, n in enumerate(names):
            gn = float(res["grads"][off + i].double().norm())
            np.testing.assert_allclose(gn, float(g[f"det.gnorm.{tag}.{n}"]), rtol=2e-3, atol=1e-9)
    np.testing.assert_allclose(res["loss"], float(g["det.loss"]), rtol=1e-5)
    # first Adam step moves every coordinate with a non-zero gradient by ~lr
    delta = (pc["rgb_linear.bias"] - before).abs()
    assert float(delta.max()) <= 5e-4 * 1.001 and float(delta.max()) > 4e-4


def test_ndc_render(golden):
    g = golden("render_ndc")
    pc, pf = O.init_params(int(g["seed_c"])), O.init_params(int(g["seed_f"]))
    rays = O.pack_rays(int(g["H"]), int(g["W"]), g["K"], T(g["rays_o"]), T(g["rays_d"]), 0.0, 1.0, ndc=True)
    with torch.no_grad():
        out = O.render_rays(rays, pc, pf, 64, 128, white_bkgd=False)
    np.testing.assert_allclose(out["rgb0"].numpy(), g["rgb0"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["rgb_map"].numpy(), g["rgb"], rtol=0, atol=5e-5)


def test_quadtree_refine_and_counts(golden):
    g = golden("quadtree")
    H, W, n = int(g["H"]), int(g["W"]), int(g["n_img"])
    leaves, ma = O.uniform_tree(H, W, 2)
    assert np.array_equal(np.array(leaves, np.float64), g["r0.boxes0"]) and ma == g["r0.minarea"][0]
    for rnd in range(4):
        total = 0
        for i in range(n):
            boxes = [tuple(b) for b in g[f"r{rnd}.boxes{i}"]]
            out, new_min = O.refine(boxes, float(g[f"r{rnd}.minarea"][i]), g[f"r{rnd}.table{i}"], 0.005)
            assert np.array_equal(np.array(out, np.float64), g[f"r{rnd}.newboxes{i}"])      # bit-exact leaf lists
            assert new_min == g[f"r{rnd}.newminarea"][i]
            total += sum(O.leaf_ray_count(b, float(g[f"r{rnd}.minarea"][i]), 1.0) for b in boxes)
        assert total == int(g[f"r{rnd}.n_rays"])
    # threshold tie: stat == fp32(thres) does not split, one ulp above does, coarser-than-minArea never does
    boxes, ma = O.uniform_tree(H, W, 2)
    out, new_min = O.refine(boxes, ma, g["tie.stat"], float(g["tie.thres"]))
    assert np.array_equal(np.array(out, np.float64), g["tie.newboxes"]) and len(out) == 4 + 3 * 2


def test_leaf_pixel_ranges():
    # tree.py:598-599 with the -0.01 quirk: a leaf ending on an integer column bound keeps the bound exclusive
    assert O.leaf_pixel_range((0, 0, 12.5, 12.5)) == (0, 13, 0, 13)
    assert O.leaf_pixel_range((12.5, 12.5, 25.0, 25.0)) == (13, 25, 13, 25)
    assert O.leaf_ray_count((0, 0, 12.5, 12.5), 156.25, 1.0) == 156
    assert O.leaf_ray_count((0, 0, 25.0, 25.0), 156.25, 1.0) == 10


def test_prob_sampling(golden):
    """prob=True (image_process.py, tree.py:583-595): sharpness maps and the pixels np.random.choice picks for the
    fixture's uniforms -- index work, exact."""
    g = golden("prob_sampling")
    boxes, ma, rf = g["boxes"], float(g["min_area"]), float(g["rand_frac"])
    for i in range(int(g["n_img"])):
        np.testing.assert_allclose(O.sharp_img(g["images"][i]), g["sharp%d" % i], atol=1e-6, rtol=0)
        u, pix, k = g["u%d" % i], g["pix%d" % i], 0
        for j, b in enumerate(boxes):
            n = int(g["counts"][j])
            got = O.emit_leaf_prob(tuple(b), ma, 1.0, rf, g["sharp%d" % i], u[k:k + n])
            assert np.array_equal(got, pix[k:k + n])
            n1 = int(n * (1 - rf))
            # prob part inside the int() block, uniform part inside the ceil() bounds
            assert (got[:n1, 0] >= int(b[0])).all() and (got[:n1, 0] < int(b[2])).all()
            assert (got[n1:, 0] >= np.ceil(b[0])).all() and (got[n1:, 0] < np.ceil(b[2])).all()
            k += n
        assert k == len(pix)
    p = O.to_prob_v2(g["sharp1"][3:11, 2:12])
    assert abs(p.sum() - 1) < 1e-12 and p.min() > 0


def test_subpixel_gather_golden(golden):
    """gen_rays_v3 (tree.py:231-307): the fixture holds the UNMODIFIED reference's outputs for seeded positions (make_golden.py
    replays its RNG); the oracle's grid_sample restatement reproduces them bit for bit, per image."""
    import warnings
    g = golden("subpixel")
    H, W = int(g["H"]), int(g["W"])
    imgs, poses = torch.from_numpy(g["images"]), torch.from_numpy(g["poses"])
    off = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i, n in enumerate(g["counts"]):
            o, d = O.camera_rays(H, W, g["K"], poses[i])
            oo, dd, cc = O.subpixel_gather(imgs[i], d, o, torch.from_numpy(g["xy"][off:off + n]))
            assert np.array_equal(oo.numpy(), g["origins"][off:off + n]) and np.array_equal(dd.numpy(), g["dirs"][off:off + n])
            assert np.array_equal(cc.numpy(), g["rgb"][off:off + n])
            assert np.all(g["leaf_id"][off:off + n, 0] == i)
            off += n
    assert O.subpixel_range((0.0, 5.0, 10.0, 12.5)) == (0, 9990, 5000, 12490)
