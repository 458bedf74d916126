"""GPU (-m gpu): the eval path (render.py:94-146) -- the fused frame front end, the on-device
PSNR / SSIM kernel and render_path."""
import os

import numpy as np
import pytest
import torch

import nerf_oracle as O

pytestmark = pytest.mark.gpu


def make_net(seed, precision):
    import model
    net = model.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True, precision=precision)
    net.load_state_dict(O.init_params(seed))
    return net.cuda()


def scene(H=40, W=56):
    from flnerf_b200 import synthetic
    K = synthetic.intrinsics(H, W, 70.0)
    poses = synthetic.lego_like_poses(3)
    return H, W, K, poses


@pytest.mark.parametrize("precision,ndc", [("bf16", False), ("bf16x3", False), ("bf16x3", True)])
def test_fused_frame_front_end_equals_the_generic_path(precision, ndc):
    """render(c2w=pose) through ops.encode_frame_tc (one kernel: get_rays + packing + depths + PE) against the same frame
    rendered from explicit get_rays() rays: same rays11, same depths, same tiles -> the same image, bit for bit."""
    import render as R, run_nerf, run_nerf_helpers as Hh
    from flnerf_b200 import ops
    H, W, K, poses = scene()
    c2w = torch.as_tensor(poses[1][:3, :4]).cuda()
    if ndc:      # a forward-facing pose for the NDC warp
        c2w = torch.tensor([[1., 0, 0, 0.1], [0, 1., 0, -0.05], [0, 0, 1., 0.2]]).cuda()
    near, far = (0.0, 1.0) if ndc else (2.0, 6.0)
    nc, nf = make_net(3, precision), make_net(4, precision)
    q = run_nerf.NetworkQuery(Hh.get_embedder(10)[0], Hh.get_embedder(4)[0], 65536)
    kw = dict(ndc=ndc, near=near, far=far, use_viewdirs=True, network_query_fn=q, network_fn=nc, network_fine=nf, N_samples=64,
              N_importance=128, white_bkgd=not ndc, perturb=0.0, raw_noise_std=0.0)
    with torch.no_grad():
        rgb_f, disp_f, acc_f, ex_f = R.render(H, W, K, chunk=1000, c2w=c2w, **kw)        # 1000: ragged last chunk
        ro, rd = Hh.get_rays(H, W, K, c2w)
        rgb_g, disp_g, acc_g, ex_g = R.render(H, W, K, chunk=1000, rays=torch.stack([ro, rd], 0), **kw)
        # the front end's own outputs against the separate kernels
        r11, z, tiles, dirpe = ops.encode_frame_tc(nc.mode, H, W, K, c2w, near, far, ndc, False, 17, 999, 64)
        r11_ref = ops.pack_rays(ro.reshape(-1, 3)[17:17 + 999], rd.reshape(-1, 3)[17:17 + 999], near, far, ndc, H, W, float(K[0][0]))
        z_ref = ops.coarse_depths(r11_ref, 64, False, False)
        tiles_ref, dirpe_ref = ops.encode_tc(r11_ref, z_ref, nc.mode)
    assert rgb_f.shape == (H, W, 3) and disp_f.shape == (H, W) and ex_f["rgb0"].shape == (H, W, 3)
    assert torch.equal(r11, r11_ref) and torch.equal(z, z_ref) and torch.equal(dirpe, dirpe_ref) and torch.equal(tiles, tiles_ref)
    assert torch.equal(rgb_f, rgb_g) and torch.equal(ex_f["rgb0"], ex_g["rgb0"]) and torch.equal(acc_f, acc_g)


def test_ssim_psnr_kernel_matches_the_reference_formula():
    """csrc/eval.cu against compute_ssim's torch formula (pinned to the reference in tests/test_oracle_vs_reference.py)."""
    import run_nerf_helpers as Hh
    from flnerf_b200 import ops
    g = torch.Generator().manual_seed(0)
    for (H, W) in ((40, 56), (33, 17), (100, 100)):
        a = torch.rand(H, W, 3, generator=g)
        yy, xx = torch.meshgrid(torch.linspace(0, 6, H), torch.linspace(0, 9, W), indexing="ij")
        b = (0.5 + 0.4 * torch.sin(yy + xx)[..., None] * torch.tensor([1.0, 0.5, -0.7])).clamp(0, 1)
        for x, y in ((a, b), (b, (b + 0.05 * torch.randn(H, W, 3, generator=g)).clamp(0, 1)), (a, a)):
            want_ssim = float(Hh.compute_ssim(x, y))                       # CPU tensors -> the torch formula
            want_psnr = float(-10 * torch.log10(torch.mean((x - y) ** 2)))
            got = ops.ssim_psnr(x.cuda(), y.cuda()).tolist()
            assert abs(got[0] - want_ssim) < 2e-5, (H, W, got, want_ssim)      # fp32 summation order of the 121-tap window
            if torch.equal(x, y):
                assert got[1] == float("inf")
            else:
                assert abs(got[1] - want_psnr) < 1e-4
            assert abs(float(Hh.compute_ssim(x.cuda(), y.cuda())) - want_ssim) < 2e-5      # the public helper takes the kernel


def test_render_path_reports_oracle_psnr_and_ssim(tmp_path, capsys):
    """render_path (render.py:94-146) on two poses: images, files and the printed metrics; frame 0 against the CPU oracle's
    rendering of the same pose (fp32-grade bf16x3 networks)."""
    import render as R, run_nerf, run_nerf_helpers as Hh
    H, W, K, poses = scene(24, 32)
    nc, nf = make_net(7, "bf16x3"), make_net(8, "bf16x3")
    q = run_nerf.NetworkQuery(Hh.get_embedder(10)[0], Hh.get_embedder(4)[0], 65536)
    kw = dict(ndc=False, near=2.0, far=6.0, use_viewdirs=True, network_query_fn=q, network_fn=nc, network_fine=nf, N_samples=64,
              N_importance=128, white_bkgd=True, perturb=False, raw_noise_std=0.0)
    gts = torch.rand(2, H, W, 3, generator=torch.Generator().manual_seed(1)).numpy()
    rp = torch.as_tensor(poses[:2]).cuda()
    rgbs, disps = R.render_path(rp, [H, W, 70.0], K, 500, kw, gt_imgs=gts, savedir=str(tmp_path))
    out = capsys.readouterr().out
    assert rgbs.shape == (2, H, W, 3) and disps.shape == (2, H, W)
    assert os.path.isfile(tmp_path / "000.png") and os.path.isfile(tmp_path / "results.txt")
    # oracle: the same frame on the CPU
    o, d = O.camera_rays(H, W, K, torch.as_tensor(poses[0][:3, :4]).float())
    rays11 = O.pack_rays(H, W, K, o.reshape(-1, 3), d.reshape(-1, 3), 2.0, 6.0, ndc=False)
    with torch.no_grad():
        ref = O.render_rays(rays11, O.init_params(7), O.init_params(8), 64, 128, white_bkgd=True)["rgb_map"].reshape(H, W, 3)
    assert float((torch.from_numpy(rgbs[0]) - ref).abs().max()) <= 1e-4
    psnr_ref = float(-10 * torch.log10(torch.mean((ref - torch.from_numpy(gts[0])) ** 2)))
    ssim_ref = float(Hh.compute_ssim(torch.from_numpy(gts[0]), ref))
    line = [l for l in out.splitlines() if l.startswith("img-0:")][0]
    psnr = float(line.split("psnr=")[1].split(",")[0]); ssim = float(line.split("ssim=")[1].split(",")[0])
    assert abs(psnr - psnr_ref) < 1e-3 and abs(ssim - ssim_ref) < 1e-4, (line, psnr_ref, ssim_ref)
    assert "mean PSNR" in open(tmp_path / "results.txt").read()
