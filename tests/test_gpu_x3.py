"""GPU (-m gpu): FLNERF_MODE_BF16X3, the split-precision tcgen05 MLP (csrc/mlp_tc_x3.cuh), against the reference
goldens and the CPU oracle at the north-star tolerance (BASELINE.json: 1e-4 relative RGB / loss).

Protocol ("parity gates"):
  * teacher-forced per stage: each network is fed the ORACLE's inputs of that stage (the fine network gets the oracle's
    merged z_vals), raw within 1e-4 of the output scale, max-rel reported;
  * rgb0 within 1e-4 absolute (colours live in [0,1]);
  * end to end: rgb relative L2 <= 1e-4 (the fine pass amplifies coarse round-off through the inverse-CDF resampling,
    so the per-pixel max is reported, and bounded at 1e-3);
  * one full training step at BASELINE configs[1] size -- 4096 rays x (64 + 128) samples -- against the oracle's
    restatement of run_nerf.py:479-494: loss within 1e-4 relative, gradient relative L2 <= 2e-3 per network.
"""
import numpy as np
import pytest
import torch

import nerf_oracle as O

pytestmark = pytest.mark.gpu


def T(a, dev="cuda"):
    return torch.from_numpy(np.asarray(a)).to(dev)


def make_net(seed, precision="bf16x3"):
    import model
    net = model.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True,
                     precision=precision)
    net.load_state_dict(O.init_params(seed))
    return net.cuda()


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30))


def max_rel(a, b, floor=1e-3):
    return float(((a.double() - b.double()).abs() / b.double().abs().clamp(min=floor)).max())


def lego_rays(B, seed=0):
    """Rays of a lego-like camera ring (radius 4, looking at the origin) through random pixels of an 800x800 frame."""
    from flnerf_b200 import synthetic
    g = torch.Generator().manual_seed(seed)
    H = W = 800
    K = synthetic.intrinsics(H, W, 1111.11)
    poses = synthetic.lego_like_poses(8)
    ro, rd = [], []
    per = B // 8
    for i in range(8):
        o, d = O.camera_rays(H, W, K, torch.as_tensor(poses[i][:3, :4]).float())
        sel = torch.randint(0, H * W, (per,), generator=g)
        ro.append(o.reshape(-1, 3)[sel]); rd.append(d.reshape(-1, 3)[sel])
    ro, rd = torch.cat(ro), torch.cat(rd)
    tgt = torch.rand(ro.shape[0], 3, generator=g)
    return H, W, K, ro, rd, tgt


def test_x3_mlp_golden_forward_backward(golden):
    """The reference's own NeRF.forward / autograd on 90-channel rows (tests/golden/mlp.npz, written by the unmodified model.py)."""
    g = golden("mlp")
    net = make_net(int(g["seed"]))
    x = T(g["x"])
    y = net(x)
    scale = float(np.abs(g["y"]).max())
    np.testing.assert_allclose(y.detach().cpu().numpy(), g["y"], atol=1e-4 * scale, rtol=1e-4)
    net.zero_grad()
    (y * T(g["gout"])).sum().backward()
    grads = {n: p.grad.detach().cpu() for n, p in net.named_parameters()}
    for k in g.files:
        if k.startswith("grad."):
            ref = g[k]
            np.testing.assert_allclose(grads[k[5:]].numpy(), ref, rtol=2e-3, atol=2e-3 * float(np.abs(ref).max()) + 1e-12)
        elif k.startswith("gradnorm."):
            np.testing.assert_allclose(float(grads[k[9:]].double().norm()), float(g[k]), rtol=1e-3)


def test_x3_teacher_forced_stages_and_end_to_end():
    from flnerf_b200 import ops
    H, W, K, ro, rd, tgt = lego_rays(512, seed=1)
    pc, pf = O.init_params(31), O.init_params(32)
    rays11 = O.pack_rays(H, W, K, ro, rd, 2.0, 6.0, ndc=False)
    with torch.no_grad():
        ref = O.render_rays(rays11, pc, pf, 64, 128, white_bkgd=True)
    nc, nf = make_net(31), make_net(32)
    r11 = rays11.cuda()
    with torch.no_grad():
        raw0 = nc.query_rays(r11, ref["z0"].cuda()).cpu()                 # stage 1: coarse net on the oracle's depths
        raw1 = nf.query_rays(r11, ref["z_vals"].cuda()).cpu()             # stage 2: fine net on the oracle's merged depths
    for name, got, want in (("coarse raw", raw0, ref["raw0"]), ("fine raw", raw1, ref["raw"])):
        scale = float(want.abs().max())
        err = float((got - want).abs().max())
        print("x3 teacher-forced %s: max abs err %.3e (scale %.3f), rel-L2 %.3e, max-rel %.3e"
              % (name, err, scale, rel_l2(got, want), max_rel(got, want)))
        assert err <= 1e-4 * scale and rel_l2(got, want) <= 2e-5
    with torch.no_grad():
        rgb0 = ops.composite_forward(raw0.cuda().contiguous(), ref["z0"].cuda(), r11[:, 3:6].contiguous(), None, True)[0].cpu()
        rgb1 = ops.composite_forward(raw1.cuda().contiguous(), ref["z_vals"].cuda(), r11[:, 3:6].contiguous(), None, True)[0].cpu()
    assert float((rgb0 - ref["rgb0"]).abs().max()) <= 1e-4 and float((rgb1 - ref["rgb_map"]).abs().max()) <= 1e-4
    # end to end through the public render() API (own depths, own resampling)
    import render as R, run_nerf, run_nerf_helpers as Hh
    q = run_nerf.NetworkQuery(Hh.get_embedder(10)[0], Hh.get_embedder(4)[0], 65536)
    with torch.no_grad():
        rgb, _, _, ex = R.render(H, W, K, rays=torch.stack([ro, rd], 0).cuda(), ndc=False, near=2.0, far=6.0, use_viewdirs=True,
                                 network_query_fn=q, network_fn=nc, network_fine=nf, N_samples=64, N_importance=128,
                                 white_bkgd=True, perturb=0.0)
    e0 = float((ex["rgb0"].cpu() - ref["rgb0"]).abs().max())
    print("x3 end to end: rgb0 max abs %.3e | rgb rel-L2 %.3e max abs %.3e max-rel %.3e"
          % (e0, rel_l2(rgb.cpu(), ref["rgb_map"]), float((rgb.cpu() - ref["rgb_map"]).abs().max()), max_rel(rgb.cpu(), ref["rgb_map"])))
    assert e0 <= 1e-4
    assert rel_l2(rgb.cpu(), ref["rgb_map"]) <= 1e-4
    assert float((rgb.cpu() - ref["rgb_map"]).abs().max()) <= 1e-3


def test_x3_render_api_matches_reference_goldens(golden):
    """tests/golden/render_rays.npz was written by the UNMODIFIED reference render() (oracle/make_golden.py)."""
    import render as R, run_nerf, run_nerf_helpers as H
    g = golden("render_rays")
    nc, nf = make_net(int(g["seed_c"])), make_net(int(g["seed_f"]))
    q = run_nerf.NetworkQuery(H.get_embedder(10)[0], H.get_embedder(4)[0], 65536)
    tgt = T(g["target"])
    for name, kw in (("det", dict(perturb=0.0, pytest=False)), ("jit", dict(perturb=1.0, pytest=True))):
        nc.zero_grad(); nf.zero_grad()
        rgb, disp, acc, ex = R.render(int(g["H"]), int(g["W"]), g["K"], chunk=32768,
                                      rays=torch.stack([T(g["rays_o"]), T(g["rays_d"])], 0), ndc=False, near=2.0, far=6.0,
                                      use_viewdirs=True, network_query_fn=q, network_fn=nc, network_fine=nf,
                                      N_samples=64, N_importance=128, white_bkgd=True, raw_noise_std=0.0, retraw=True, **kw)
        np.testing.assert_allclose(ex["rgb0"].detach().cpu().numpy(), g[f"{name}.rgb0"], atol=1e-4, rtol=0)
        np.testing.assert_allclose(rgb.detach().cpu().numpy(), g[f"{name}.rgb"], atol=1e-4, rtol=0)
        loss = H.img2mse(rgb, tgt) + H.img2mse(ex["rgb0"], tgt)
        np.testing.assert_allclose(float(loss), float(g[f"{name}.loss"]), rtol=1e-4)
        np.testing.assert_allclose(acc.detach().cpu().numpy(), g[f"{name}.acc"], atol=1e-4)
        loss.backward()
        for tag, net in (("c", nc), ("f", nf)):
            for n_, p_ in net.named_parameters():
                ref = float(g[f"{name}.gnorm.{tag}.{n_}"])
                np.testing.assert_allclose(float(p_.grad.double().norm()), ref, rtol=5e-3, atol=1e-9)
                key = f"{name}.grad.{tag}.{n_}"
                if key in g.files:
                    np.testing.assert_allclose(p_.grad.cpu().numpy(), g[key], rtol=5e-3, atol=2e-3 * float(np.abs(g[key]).max()) + 1e-12)


def oracle_step_chunked(rays11, tgt, pc, pf, n_coarse, n_fine, chunk=512, **kw):
    """O.train_step's loss and gradients accumulated over ray chunks (bounds the autograd memory of the CPU oracle:
    the loss is a mean over rays, so chunk gradients add)."""
    params = list(pc.values()) + list(pf.values())
    for q in params:
        q.requires_grad_(True)
        q.grad = None
    B = rays11.shape[0]
    tot_f = tot_c = 0.0
    rgbs, rgb0s = [], []
    for i in range(0, B, chunk):
        out = O.render_rays(rays11[i:i + chunk], pc, pf, n_coarse, n_fine, **kw)
        t = tgt[i:i + chunk]
        lf = ((out["rgb_map"] - t) ** 2).sum() / (3.0 * B)
        lc = ((out["rgb0"] - t) ** 2).sum() / (3.0 * B)
        (lf + lc).backward()
        tot_f += float(lf.detach().double()); tot_c += float(lc.detach().double())
        rgbs.append(out["rgb_map"].detach()); rgb0s.append(out["rgb0"].detach())
    grads = [q.grad.clone() for q in params]
    for q in params:
        q.requires_grad_(False)
        q.grad = None
    return {"loss": tot_f + tot_c, "rgb": torch.cat(rgbs), "rgb0": torch.cat(rgb0s), "grads": grads}


@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
def test_parity_training_step_at_config2_size(precision):
    """BASELINE.json configs[1]: N_rand = 4096 rays, 64 coarse + 128 fine samples, lego-like cameras, white background.
    One fused training step against one reference iteration of the oracle on the same rays."""
    from flnerf_b200.engine import FusedAdam, Trainer
    H, W, K, ro, rd, tgt = lego_rays(4096, seed=2)
    nc, nf = make_net(41, precision), make_net(42, precision)
    opt = FusedAdam(list(nc.parameters()) + list(nf.parameters()), [nc, nf], lr=5e-4)
    tr = Trainer(nc, nf, opt, H, W, K, 2.0, 6.0, 64, 128, white_bkgd=True, perturb=0.0)
    loss = tr.step(ro.cuda(), rd.cuda(), tgt.cuda())
    torch.cuda.synchronize()
    pc, pf = O.init_params(41), O.init_params(42)
    rays11 = O.pack_rays(H, W, K, ro, rd, 2.0, 6.0, ndc=False)
    res = oracle_step_chunked(rays11, tgt, pc, pf, 64, 128, white_bkgd=True)
    rgb, rgb0 = tr.last["rgb"].cpu(), tr.last["rgb0"].cpu()
    g = tr.bucket.cpu()
    g_or = torch.cat([x.reshape(-1) for x in res["grads"]])
    half = g.numel() // 2
    print("%s @4096x(64+128): loss %.8f oracle %.8f (rel %.2e) | rgb0 max abs %.2e | rgb rel-L2 %.2e max abs %.2e max-rel %.2e | "
          "grad rel-L2 coarse %.2e fine %.2e"
          % (precision, float(loss.sum()), res["loss"], abs(float(loss.sum()) - res["loss"]) / res["loss"],
             float((rgb0 - res["rgb0"]).abs().max()), rel_l2(rgb, res["rgb"]), float((rgb - res["rgb"]).abs().max()),
             max_rel(rgb, res["rgb"]), rel_l2(g[:half], g_or[:half]), rel_l2(g[half:], g_or[half:])))
    np.testing.assert_allclose(float(loss.sum()), res["loss"], rtol=1e-4)
    assert float((rgb0 - res["rgb0"]).abs().max()) <= 1e-4
    assert rel_l2(rgb, res["rgb"]) <= 1e-4
    assert rel_l2(g[:half], g_or[:half]) <= 2e-3 and rel_l2(g[half:], g_or[half:]) <= 2e-3


def test_x3_inference_equals_training_forward_and_row_permutation():
    from flnerf_b200 import ops
    torch.manual_seed(0)
    B, S = 300, 40                                    # 12000 rows: 93.75 tiles -> padded last pair
    rays = torch.cat([torch.randn(B, 3) * 0.5, torch.nn.functional.normalize(torch.randn(B, 3), dim=-1),
                      2 * torch.ones(B, 1), 6 * torch.ones(B, 1), torch.nn.functional.normalize(torch.randn(B, 3), dim=-1)], -1).cuda()
    z = torch.sort(torch.rand(B, S, device="cuda") * 4 + 2, -1)[0]
    net = make_net(9)
    raw_t = net.query_rays(rays, z)                   # training forward (stash written)
    with torch.no_grad():
        raw_i = net.query_rays(rays, z)               # inference
        perm = torch.randperm(B, device="cuda")
        raw_p = net.query_rays(rays[perm].contiguous(), z[perm].contiguous())
    assert torch.equal(raw_t.detach(), raw_i)
    assert torch.equal(raw_p, raw_i[perm])            # a row's result does not depend on its tile / CTA
    # the bf16 path must stay within its own (looser) band of the x3 path: the two share everything but operand precision
    with torch.no_grad():
        raw16 = make_net(9, "bf16").query_rays(rays, z)
    assert rel_l2(raw16, raw_i) < 2e-2


def test_psnr_gate_cut_down():
    """A cut-down run of tools/psnr_check.py's protocol (the full one: 106 paired seeds, 200x200, 4000 iterations ->
    profiles/r02_psnr_gate_pooled.json: bf16 - bf16x3 = -0.01 dB, 95 % CI [-0.07, +0.06]): 48x48 views, 400 iterations, three
    seeds, both arms from the same (torch default) initial weights on the same batches.  At this size the check is a guard
    against a broken arm (a diverged mode shows up as many dB), not a 0.1 dB measurement; a seed whose initialisation is dead
    (white image, ~5.5 dB: one in ten NeRF initialisations) collapses in BOTH arms alike and contributes a zero."""
    import render as R, run_nerf, run_nerf_helpers as Hh, tree
    from flnerf_b200 import synthetic
    from flnerf_b200.engine import FusedAdam, Trainer
    H = W = 48
    K = synthetic.intrinsics(H, W, 0.5 * W / np.tan(0.5 * 0.6911112070083618))
    poses, test_poses = synthetic.lego_like_poses(8), synthetic.lego_like_poses(2, phi=-25.0)
    imgs, test_imgs = synthetic.render_scene(H, W, K, poses, n_samples=64), synthetic.render_scene(H, W, K, test_poses, n_samples=64)
    q = run_nerf.NetworkQuery(Hh.get_embedder(10)[0], Hh.get_embedder(4)[0], 65536)
    psnr = {}
    import model
    iters = 400
    for seed in (0, 1, 2):
        for prec in ("bf16x3", "bf16"):
            torch.manual_seed(seed)
            nc = model.NeRF(8, 256, 63, 27, 5, [4], True, precision=prec).cuda()
            nf = model.NeRF(8, 256, 63, 27, 5, [4], True, precision=prec).cuda()
            opt = FusedAdam(list(nc.parameters()) + list(nf.parameters()), [nc, nf], lr=5e-4)
            tr = Trainer(nc, nf, opt, H, W, K, 2.0, 6.0, 64, 128, white_bkgd=True, perturb=1.0, seed=seed, graph=True)
            mgr = tree.QuadTreeManager(H, W, K, imgs, torch.as_tensor(poses[:, :3, :4]), mseThres=0.0, max_depth=2, max_level=4, seed=seed)
            it = 0
            while it < iters:
                n = mgr.emit_epoch()
                for first in range(0, n - 512, 512):
                    tr.step_from_tree(mgr, first, 512)
                    for g in opt.param_groups:
                        g["lr"] = 5e-4 * (0.1 ** (it / float(iters)))
                    it += 1
                    if it >= iters:
                        break
            vals = []
            with torch.no_grad():
                for c2w, gt in zip(test_poses, test_imgs):
                    rgb = R.render(H, W, K, chunk=32768, c2w=torch.as_tensor(c2w[:3, :4]).cuda(), ndc=False, near=2.0, far=6.0,
                                   use_viewdirs=True, network_query_fn=q, network_fn=nc, network_fine=nf, N_samples=64,
                                   N_importance=128, white_bkgd=True, perturb=0.0)[0]
                    vals.append(float(-10 * torch.log10(torch.mean((rgb - gt) ** 2))))
            psnr[(seed, prec)] = float(np.mean(vals))
    d = [psnr[(s, "bf16")] - psnr[(s, "bf16x3")] for s in (0, 1, 2)]
    print("cut-down PSNR gate: %s, bf16 - bf16x3 = %s dB" % ({k: round(v, 2) for k, v in psnr.items()}, [round(x, 2) for x in d]))
    assert all(np.isfinite(v) for v in psnr.values()), psnr
    assert abs(float(np.mean(d))) < 2.0 and max(abs(x) for x in d) < 4.0, d
