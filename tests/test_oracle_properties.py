"""CPU: size-independent properties of the oracle (hypothesis) -- the same invariants the GPU tests assert at full size
(tests/test_gpu_mlp.py::test_full_size_properties) hold for the checker itself on arbitrary small inputs."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

import nerf_oracle as O
import nerfpp_oracle as P

SET = dict(max_examples=25, deadline=None)


@settings(**SET)
@given(st.integers(1, 6), st.integers(2, 40), st.booleans(), st.integers(0, 2 ** 16))
def test_composite_weights_are_a_sub_probability(B, S, white, seed):
    g = torch.Generator().manual_seed(seed)
    raw = torch.randn(B, S, 4, generator=g) * 4
    z = torch.sort(torch.rand(B, S, generator=g) * 4 + 2, -1)[0]
    d = torch.randn(B, 3, generator=g)
    rgb, disp, acc, w, depth = O.composite(raw, z, d, None, white)
    assert float(w.min()) >= 0 and float(acc.max()) <= 1 + 1e-5
    np.testing.assert_allclose(w.sum(-1).numpy(), acc.numpy(), atol=1e-5)
    assert float(rgb.min()) >= -1e-6 and float(rgb.max()) <= 1 + 1e-5
    ok = acc > 1e-6
    assert bool(((depth[ok] / acc[ok]) >= z[ok].min(-1)[0] - 1e-4).all()) and bool(((depth[ok] / acc[ok]) <= z[ok].max(-1)[0] + 1e-4).all())


@settings(**SET)
@given(st.integers(1, 5), st.integers(3, 30), st.integers(2, 40), st.integers(0, 2 ** 16))
def test_fine_depths_sorted_merge_and_inside_the_bins(B, Nc, Nf, seed):
    g = torch.Generator().manual_seed(seed)
    z = torch.sort(torch.rand(B, Nc, generator=g) * 4 + 2, -1)[0]
    w = torch.rand(B, Nc, generator=g)
    u = torch.rand(B, Nf, generator=g)
    merged, zs = O.fine_depths(z, w, Nf, u)
    assert merged.shape == (B, Nc + Nf) and bool((merged[:, 1:] >= merged[:, :-1]).all())
    assert torch.equal(merged, torch.sort(torch.cat([z, zs], -1), -1)[0])
    mid = 0.5 * (z[:, 1:] + z[:, :-1])
    assert bool((zs >= mid[:, :1] - 1e-5).all()) and bool((zs <= mid[:, -1:] + 1e-5).all())
    # monotone in u: sorting the uniforms sorts the samples
    _, zs_sorted = O.fine_depths(z, w, Nf, torch.sort(u, -1)[0])
    assert bool((zs_sorted[:, 1:] >= zs_sorted[:, :-1] - 1e-6).all())


@settings(**SET)
@given(st.integers(1, 4), st.integers(0, 2 ** 16), st.floats(0.0, 0.02))
def test_quadtree_refine_keeps_a_partition_in_dfs_order(depth, seed, thres):
    H = W = 50
    leaves, min_area = O.uniform_tree(H, W, depth)
    rs = np.random.RandomState(seed)
    for _ in range(3):
        stat = rs.uniform(0, 0.03, len(leaves)).astype(np.float32)
        new, new_min = O.refine(leaves, min_area, stat, thres)
        assert abs(sum(O.area(b) for b in new) - H * W) < 1e-6                          # still tiles the image
        split = [i for i, b in enumerate(leaves) if stat[i] > np.float32(thres) and O.area(b) == min_area]
        assert len(new) == len(leaves) + 3 * len(split)
        assert new_min == (min_area / 4 if split else min_area)
        # children replace their parent in place: removing every split group restores the old order
        it, k = iter(new), 0
        for i, b in enumerate(leaves):
            if i in split:
                kids = [next(it) for _ in range(4)]
                assert kids == O.split4(b)
            else:
                assert next(it) == b
        # a leaf coarser than min_area is frozen at 10 rays, the finest get int(area)
        for b in new:
            n = O.leaf_ray_count(b, new_min, 1.0)
            assert n == (10 if O.area(b) > new_min + 0.01 else int(O.area(b)))
        leaves, min_area = new, new_min


@settings(**SET)
@given(st.integers(1, 5), st.integers(3, 24), st.integers(2, 24), st.integers(0, 2 ** 16))
def test_nerfpp_sampling_properties(B, M1, N, seed):
    g = torch.Generator().manual_seed(seed)
    bins = torch.sort(torch.rand(B, M1, generator=g), -1)[0]
    w = torch.rand(B, M1 - 1, generator=g)
    s = P.sample_pdf(bins, w, N, torch.sort(torch.rand(B, N, generator=g), -1)[0])
    assert bool((s >= bins[:, :1] - 1e-6).all()) and bool((s <= bins[:, -1:] + 2e-6).all())
    assert bool((s[:, 1:] >= s[:, :-1] - 1e-6).all())
    o = torch.randn(B, 3, generator=g) * 0.2
    d = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1)
    far = P.intersect_sphere(o, d)
    hit = o + far[:, None] * d
    np.testing.assert_allclose(hit.norm(dim=-1).numpy(), 1.0, atol=1e-5)                 # the exit point lies on the sphere
    depth = torch.rand(B, 7, generator=g) * 0.98 + 0.01
    pts, real = P.depth2pts_outside(o[:, None].expand(-1, 7, -1), d[:, None].expand(-1, 7, -1), depth)
    np.testing.assert_allclose(pts[..., :3].norm(dim=-1).numpy(), 1.0, atol=1e-5)       # unit direction + 1/r
    world = o[:, None] + real[..., None] * d[:, None]                                    # the same point, conventional depth
    np.testing.assert_allclose((world.norm(dim=-1) * depth).numpy(), 1.0, atol=2e-3)
