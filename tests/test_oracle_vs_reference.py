"""CPU, build container only: the oracle against the live, unmodified reference (skipped where /root/reference is
absent, e.g. on the GPU box).  Complements the golden fixtures with fresh random inputs."""
import numpy as np
import pytest
import torch

import nerf_oracle as O
import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load()


def test_composite_and_sample_pdf_random(ref):
    torch.manual_seed(7)
    raw = torch.randn(9, 40, 4) * 3
    z = torch.sort(torch.rand(9, 40) * 4 + 2, -1)[0]
    d = torch.randn(9, 3)
    for wb in (False, True):
        a = O.composite(raw, z, d, None, wb)
        b = ref.render.raw2outputs(raw, z, d, 0, wb)
        for x, y in zip(a, b):
            assert torch.equal(torch.nan_to_num(x, nan=-7), torch.nan_to_num(y, nan=-7))
    w = a[3]
    mid = 0.5 * (z[:, 1:] + z[:, :-1])
    assert torch.equal(O.inverse_cdf(mid, w[:, 1:-1], 33, None), ref.helpers.sample_pdf(mid, w[:, 1:-1], 33, det=True))


def test_mlp_matches_reference_module(ref):
    p = O.init_params(5)
    m = ref.model.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    m.load_state_dict(p)
    assert [k for k, _ in m.named_parameters()] == list(p.keys())            # parameters() order == flat layout
    x = torch.randn(50, 90)
    np.testing.assert_allclose(O.mlp_forward(p, x).numpy(), m(x).detach().numpy(), atol=2e-6, rtol=0)


def test_adam_matches_torch():
    torch.manual_seed(0)
    w = [torch.randn(7, 5), torch.randn(5)]
    ref_w = [t.clone().requires_grad_(True) for t in w]
    opt = torch.optim.Adam(ref_w, lr=5e-4, betas=(0.9, 0.999))
    mine = O.AdamState([t.clone() for t in w])
    for _ in range(5):
        gs = [torch.randn_like(t) for t in w]
        for t, g in zip(ref_w, gs):
            t.grad = g.clone()
        opt.step()
        mine.step(gs)
    for a, b in zip(mine.params, ref_w):
        np.testing.assert_allclose(a.numpy(), b.detach().numpy(), rtol=1e-6, atol=1e-8)


def test_prob_sampling_matches_image_processor(ref):
    """image_process.py (real cv2.blur / cvtColor, np.random.choice) against the oracle's restatement."""
    IP = ref.image_process.ImageProcessor
    rs = np.random.RandomState(3)
    imgs = rs.uniform(0, 1, (2, 37, 29, 3)).astype(np.float32)
    imgs[1, 5:20, 3:17] = 1.0                                   # a flat patch: variance 0 -> the 0.01*mean floor matters
    proc = IP(torch.from_numpy(imgs), scale=0)
    for i in range(2):
        np.testing.assert_allclose(O.sharp_img(imgs[i]), proc.sharp_imgs[i], atol=1e-6, rtol=0)
        for (a, b, c, d) in [(0, 0, 37, 29), (5, 3, 20, 17), (18, 14, 37, 29)]:
            block = proc.sharp_imgs[i][a:c, b:d]
            np.testing.assert_allclose(O.to_prob_v2(block), proc.to_prob_v2(block), rtol=1e-12, atol=0)
            np.random.seed(11 + i)
            want = proc.sample_pixels(block, 500).numpy()
            np.random.seed(11 + i)
            u = np.random.random_sample(500)
            assert np.array_equal(O.sample_pixels(block, u), want)


def test_compute_ssim_matches_reference(ref):
    """run_nerf_helpers.compute_ssim (run_nerf_helpers.py:158-234, the metric render_path reports) -- pure torch on both
    sides, so the product function is compared directly (it needs no device)."""
    import importlib
    ours = importlib.import_module("run_nerf_helpers")          # fast-learning-nerf_b200/run_nerf_helpers.py (conftest path)
    rs = np.random.RandomState(2)
    a = torch.from_numpy(rs.uniform(0, 1, (40, 36, 3)).astype(np.float32))
    b = (a + torch.from_numpy(rs.normal(0, 0.05, (40, 36, 3)).astype(np.float32))).clamp(0, 1)
    want = ref.helpers.compute_ssim(a, b)
    got = ours.compute_ssim(a, b)
    np.testing.assert_allclose(float(got), float(want), rtol=1e-5)
    np.testing.assert_allclose(ours.compute_ssim(a, b, return_map=True).numpy(), ref.helpers.compute_ssim(a, b, return_map=True).numpy(),
                               atol=2e-5)
    assert abs(float(ours.compute_ssim(a, a)) - 1.0) < 1e-6
