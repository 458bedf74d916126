"""CPU: host-side logic that needs no device -- CLI/config surface, tree (de)serialisation mirror, parameter
layout, checkpoint key compatibility."""
import io
import os
import pickle

import numpy as np
import torch

import nerf_oracle as O
from conftest import PKG


def test_config_parser_defaults_match_reference_flags():
    import argument_parser
    a = argument_parser.config_parser().parse_args([])
    exp = dict(basedir='./logs/', netdepth=8, netwidth=256, N_rand=4096, lrate=5e-4, lrate_decay=250, chunk=32768,
               netchunk=65536, N_samples=64, N_importance=0, perturb=1.0, multires=10, multires_views=4,
               raw_noise_std=0.0, n_epoch=12, init_level=3, subdivide_every=1, subdivide_thres=0.015,
               randSamp_perc=0.5, dataset_type='llff', testskip=8, factor=8, llffhold=8, precrop_frac=0.5,
               i_embed=0, use_viewdirs=False, white_bkgd=False, no_ndc=False, lindisp=False, render_only=False)
    for k, v in exp.items():
        assert getattr(a, k) == v, k


def test_config_file_and_cli_override():
    import argument_parser
    cfg = os.path.join(PKG, "configs", "lego.txt")
    a = argument_parser.config_parser().parse_args(["--config", cfg, "--N_rand", "4096"])
    assert a.N_rand == 4096 and a.white_bkgd and a.use_viewdirs and a.no_batching
    assert (a.N_samples, a.N_importance, a.n_epoch, a.init_level, a.subdivide_every) == (64, 128, 18, 2, 3)
    assert a.subdivide_thres == 0.001 and a.lrate_decay == 500 and a.dataset_type == 'blender'


def test_tree_mirror_roundtrip_and_pickle(golden):
    import tree
    g = golden("quadtree")
    for rnd in (0, 3):
        boxes = g[f"r{rnd}.newboxes1"]
        t = tree.QuadTree((int(g["H"]), int(g["W"])), 0.0, 1, _boxes=boxes, _min_area=float(g[f"r{rnd}.newminarea"][1]))
        leaves = tree.get_children(t.root)
        assert np.array_equal(np.array([l.box() for l in leaves], np.float64), boxes)       # DFS order preserved
        t2 = pickle.load(io.BytesIO(pickle.dumps([t])))[0]                                   # pickled by qualified name
        assert type(t2).__module__ == "tree" and t2.minArea == t.minArea
        assert [l.box() for l in tree.get_children(t2.root)] == [l.box() for l in leaves]
    # reference-style construction still works (uniform tree when thres <= 0)
    t = tree.QuadTree(np.zeros((64, 64, 3), np.float32), 0.0, 3)
    assert len(tree.get_children(t.root)) == 16 and t.minArea == 64 * 64 / 16
    assert [l.box() for l in tree.get_children(t.root)] == O.uniform_tree(64, 64, 3)[0]
    import tree_utils
    assert len(tree.get_children(tree_utils.SimpleQuadTree(32, 32, 2).root)) == 4


def test_parameter_layout_matches_reference_order():
    import model
    net = model.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    names = [n for n, _ in net.named_parameters()]
    ref = list(O.init_params(0).keys())
    assert names == ref
    shapes = {n: tuple(p.shape) for n, p in net.named_parameters()}
    assert shapes["pts_linears.5.weight"] == (256, 319) and shapes["views_linears.0.weight"] == (128, 283)
    assert sum(p.numel() for p in net.parameters()) == 595844
    # offsets used by the kernels (csrc/mlp_layout.h) == running sum over parameters()
    off, table = 0, {}
    for n, p in net.named_parameters():
        table[n] = off
        off += p.numel()
    assert table["pts_linears.5.weight"] == 279552 and table["views_linears.0.weight"] == 493056
    assert table["feature_linear.weight"] == 529408 and table["alpha_linear.weight"] == 595200
    assert table["rgb_linear.weight"] == 595457 and table["rgb_linear.bias"] == 595841
    # DataParallel-style checkpoint keys
    import run_nerf
    sd = run_nerf.ModuleHolder(net).state_dict()
    assert all(k.startswith("module.") for k in sd) and "module.pts_linears.0.weight" in sd


def test_lr_schedule_and_feistel_free_logic():
    from flnerf_b200.engine import lr_at
    assert abs(lr_at(5e-4, 500, 0) - 5e-4) < 1e-12
    assert abs(lr_at(5e-4, 500, 500000) - 5e-5) < 1e-12


def test_variance_driven_initial_tree_matches_reference(golden):
    """mseThres > 0 (tree.py:28-56,655-676): the host recursion that seeds the GPU SoA -- leaf lists bit-exact."""
    import tree
    g = golden("variance_tree")
    for k, th in enumerate(g["thres"]):
        t = tree.QuadTree(g["image"], float(th), int(g["max_depth"]))
        boxes = np.array([n.box() for n in tree.get_children(t.root)], np.float64)
        assert np.array_equal(boxes, g["boxes%d" % k]) and t.minArea == float(g["minarea%d" % k])
        # DFS leaf order survives the SoA round trip used by QuadTreeManager.quadTrees
        t2 = tree.QuadTree((48, 48), 0.0, 1, _boxes=boxes, _min_area=t.minArea)
        assert [n.box() for n in tree.get_children(t2.root)] == [tuple(b) for b in boxes]


def test_bench_clock_sampler_window():
    """bench.py's nvidia-smi sampler: rows are windowed by arrival time; a timed region shorter than the sampling period
    falls back to every row taken under load and says so."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(PKG), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    cs = bench.ClockSampler(0)
    cs.proc = object()                       # pretend nvidia-smi is running
    row = lambda mhz, cap: [str(mhz), "1965", "700.0", "Not Active", "Not Active", "Not Active", cap]
    cs.rows = [(1.0, row(1900, "Not Active")), (2.0, row(1950, "Active")), (2.5, row(1930, "Active")), (9.0, row(600, "Not Active"))]
    w = cs.window(1.5, 3.0)
    assert w["samples"] == 2 and w["sm_mhz"] == 1950.0 and w["sm_max_mhz"] == 1965.0 and w["reasons"] == ["sw_power_cap"]
    assert w["window"] == "timed region"
    w = cs.window(3.1, 3.2)                  # nothing inside: fall back to all rows
    assert w["samples"] == 4 and "no sample fell inside" in w["window"]
    none = bench.ClockSampler(0)
    assert none.window(0, 1)["reasons"] == ["nvidia-smi unavailable"]


def _no_viewdirs_reference(sd, x):
    """nerf-ours/model.py:38-63 with use_viewdirs=False, restated on a reference-named state dict: eight relu(Linear) layers with
    the skip concatenation after layer 4, then output_linear (W -> output_ch)."""
    h = x
    for i in range(8):
        h = torch.relu(torch.nn.functional.linear(h, sd["pts_linears.%d.weight" % i], sd["pts_linears.%d.bias" % i]))
        if i == 4:
            h = torch.cat([x, h], -1)
    return torch.nn.functional.linear(h, sd["output_linear.weight"], sd["output_linear.bias"])


def test_no_viewdirs_model_maps_exactly_onto_the_viewdirs_kernels():
    """use_viewdirs=False (model.py:55-63) runs on an inner use_viewdirs=True net through a frozen +-1 adapter (model.NeRF.
    _build_inner).  CPU check of the mapping, with the ORACLE's use_viewdirs=True forward standing in for the kernels: same
    parameter names / shapes / initial values as the reference's constructor, forward equal to the reference formula, and the
    gradients of the inner net's live entries equal to the reference's output_linear / pts_linears gradients."""
    import model
    torch.manual_seed(0)
    m = model.NeRF(D=8, W=256, input_ch=63, input_ch_views=0, output_ch=5, skips=[4], use_viewdirs=False, precision="fp32")
    sd = m.state_dict()
    assert list(sd)[-4:] == ["views_linears.0.weight", "views_linears.0.bias", "output_linear.weight", "output_linear.bias"]
    assert sd["views_linears.0.weight"].shape == (128, 256) and sd["output_linear.weight"].shape == (5, 256) and len(sd) == 20
    import ref_shim
    if ref_shim.available():                                 # the unmodified constructor, same seed: same tensors
        torch.manual_seed(0)
        ref = ref_shim.load().model.NeRF(D=8, W=256, input_ch=63, input_ch_views=0, output_ch=5, skips=[4], use_viewdirs=False)
        assert list(ref.state_dict()) == list(sd) and all(torch.equal(sd[k], v) for k, v in ref.state_dict().items())
    x = torch.randn(48, 63, generator=torch.Generator().manual_seed(1))
    pub = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    y_ref = _no_viewdirs_reference(pub, x)
    inner = {k: v.detach().clone().requires_grad_(True) for k, v in m.kernel_net.state_dict().items()}
    y_in = O.mlp_forward(inner, torch.cat([x, torch.zeros(48, 27)], -1))
    assert float((y_in - y_ref[:, :4]).abs().max()) <= 1e-6
    g = torch.randn(48, 4, generator=torch.Generator().manual_seed(2))
    (y_in * g).sum().backward()
    (y_ref[:, :4] * g).sum().backward()
    tol = dict(rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(inner["feature_linear.weight"].grad[0:3], pub["output_linear.weight"].grad[0:3], **tol)
    np.testing.assert_allclose(inner["feature_linear.bias"].grad[0:3], pub["output_linear.bias"].grad[0:3], **tol)
    np.testing.assert_allclose(inner["alpha_linear.weight"].grad, pub["output_linear.weight"].grad[3:4], **tol)
    for i in range(8):
        np.testing.assert_allclose(inner["pts_linears.%d.weight" % i].grad, pub["pts_linears.%d.weight" % i].grad, **tol)
    assert float(inner["feature_linear.weight"].grad[3:].abs().max()) == 0.0
    # checkpoint round trip through the "module." holder: the inner net follows a load, state_dict() follows the inner net
    import run_nerf
    h = run_nerf.ModuleHolder(m)
    shifted = {k: v + 0.5 for k, v in h.state_dict().items()}
    h.load_state_dict(shifted)
    assert torch.equal(m._inner.alpha_linear.weight, shifted["module.output_linear.weight"][3:4])
    with torch.no_grad():
        m._inner.feature_linear.weight[1].fill_(7.0)          # what the fused optimiser does: update the inner net only
    assert torch.equal(h.state_dict()["module.output_linear.weight"][1], torch.full((256,), 7.0))
    assert torch.equal(h.state_dict()["module.output_linear.weight"][4], shifted["module.output_linear.weight"][4])
