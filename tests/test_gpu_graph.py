"""GPU (-m gpu): the training step replayed as ONE CUDA graph (engine.Trainer(graph=True); include/flnerf.h
flnerf_step_record) against the same steps launched kernel by kernel.  Both run the same kernels on the same rays with
the same Philox offsets and Adam scalars; only the summation order of the weight-gradient atomics differs."""
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


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_graph_step_equals_eager_step(precision):
    import tree
    from flnerf_b200 import lib, synthetic
    from flnerf_b200.engine import FusedAdam, Trainer
    H = W = 64
    K = synthetic.intrinsics(H, W, 88.0)
    poses = synthetic.lego_like_poses(4)
    imgs = synthetic.render_scene(H, W, K, poses, n_samples=32)
    runs = {}
    for graph in (False, True):
        nc, nf = make_net(5, precision), make_net(6, precision)
        opt = FusedAdam(list(nc.parameters()) + list(nf.parameters()), [nc, nf], lr=5e-4)
        tr = Trainer(nc, nf, opt, H, W, K, 2.0, 6.0, 64, 128, white_bkgd=True, perturb=1.0, seed=3, graph=graph)
        mgr = tree.QuadTreeManager(H, W, K, imgs, torch.as_tensor(poses[:, :3, :4]), mseThres=0.0, max_depth=2, max_level=4, seed=1)
        n = mgr.emit_epoch()
        losses, launches = [], []
        for i, first in enumerate(range(0, 8 * 512, 512)):
            for g in opt.param_groups:               # a per-step learning-rate schedule, as run_nerf.py:498-502 applies it
                g["lr"] = 5e-4 * (0.1 ** (i / 20.0))
            c0 = lib.launch_count()
            losses.append(tr.step_from_tree(mgr, first, 512).clone())
            launches.append(lib.launch_count() - c0)
        # a ragged tail batch falls back to the eager path
        tail = tr.step_from_tree(mgr, n - 100, 512).clone()
        runs[graph] = dict(loss=torch.stack(losses).cpu(), tail=tail.cpu(), w=torch.cat([nc.flat_parameters(), nf.flat_parameters()]).cpu(),
                           leaf=mgr.leaf_max.clone().cpu(), steps=opt.state_dict()["state"][0]["step"], launches=launches,
                           calls=tr.calls)
    a, b = runs[False], runs[True]
    np.testing.assert_allclose(b["loss"].numpy(), a["loss"].numpy(), rtol=2e-3 if precision == "bf16" else 1e-4)
    np.testing.assert_allclose(b["tail"].numpy(), a["tail"].numpy(), rtol=2e-3 if precision == "bf16" else 1e-4)
    assert float(a["steps"]) == float(b["steps"]) == 9.0 and a["calls"] == b["calls"]
    d = (a["w"] - b["w"]).abs()
    assert float(d.mean()) < 2e-5 and float(d.max()) <= 9 * 2.1 * 5e-4      # Adam: sign flips of ~zero gradients only
    np.testing.assert_allclose(b["leaf"].numpy(), a["leaf"].numpy(), atol=2e-3 if precision == "bf16" else 1e-5)
    # steps 3.. are replays: the same kernels as the eager step, accounted for, plus the one record-update kernel
    assert b["launches"][0] == a["launches"][0] and b["launches"][2] == a["launches"][2] + 1
    assert b["launches"][2:] == [b["launches"][2]] * 6
