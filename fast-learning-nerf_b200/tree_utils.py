"""Drop-in for nerf-ours/tree_utils.py.  In the reference this module is dead code on the nerf-ours path: its
only call site (RaySampler.pre_gen_rays_v3) is commented out (run_nerf.py:357-362) and the consumer
gen_rays_v4 (tree.py:430-490) is never called.  The names are kept importable."""
from tree import QuadTreeNode, get_children  # noqa: F401
from flnerf_b200.lib import FlnerfError


def _subdivide(node, depth, max_depth):
    if depth >= max_depth:
        return
    node.subdivide_once()
    for c in node.children:
        _subdivide(c, depth + 1, max_depth)


class SimpleQuadTree:
    """Uniform tree of depth ``max_depth`` over an h x w image (tree_utils.py:15-25)."""

    def __init__(self, h, w, max_depth: int):
        self.H, self.W = h, w
        self.root = QuadTreeNode(0, 0, h, w)
        _subdivide(self.root, 1, max_depth)


class RaySampler:
    def __init__(self, images, rays_dir, rays_origin, max_level):
        self.n_images, self.h, self.w = images.shape[:3]
        self.images, self.dirs, self.origins, self.max_level = images, rays_dir, rays_origin, max_level
        self.epoch_size = self.n_images * self.h * self.w

    def pre_gen_rays_v3(self, down_scale=1, rand_samp_prec=0.2, dset_name='lego'):
        raise FlnerfError("RaySampler.pre_gen_rays_v3 is unused by nerf-ours/run_nerf.py (its call is commented out); "
                          "per-epoch emission runs on the GPU: QuadTreeManager.emit_epoch / gen_rays_v3_multiThread")
