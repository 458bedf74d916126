"""Drop-in for nerf-ours/tree.py: ``QuadTreeNode``, ``QuadTree``, ``QuadTreeManager``, ``get_children``,
``recursive_subdivide`` keep their names/attributes (``treeDivide_*.pkl`` pickles them by qualified name),
but the per-image trees LIVE ON THE GPU as structure-of-arrays: DFS-ordered leaf boxes
``boxes[n_images, cap, 4]`` (float64, exact dyadic fractions of H and W), ``count[n_images]``,
``min_area[n_images]``.  Ray emission (tree.py:569-626), the per-leaf max-|gt-pred| statistic and the
refinement (tree.py:533-557, 629-652) are kernels in csrc/train_ops.cu; the Python node objects are a
lazily rebuilt mirror used only for (de)serialisation.
"""
import math
from typing import List

import numpy as np
import torch

from flnerf_b200 import ops
from flnerf_b200.lib import FlnerfError


class QuadTreeNode:
    """A box (x0, y0, x1, y1): x = image row in [0, H], y = column in [0, W] (tree.py:92)."""

    def __init__(self, x0, y0, x1, y1):
        self.x0, self.y0, self.x1, self.y1 = x0, y0, x1, y1
        self.children = []

    def box(self):
        return (self.x0, self.y0, self.x1, self.y1)

    def get_error(self, img):
        """Sum over channels of the per-channel pixel variance inside the box (tree.py:28-56)."""
        px = img[math.ceil(self.x0):math.floor(self.x1), math.ceil(self.y0):math.floor(self.y1), :]
        px = px.detach().cpu().numpy() if torch.is_tensor(px) else np.asarray(px)
        return float(sum(np.square(px[:, :, c] - px[:, :, c].mean()).mean() for c in range(3)))

    def subdivide_once(self):
        mx, my = (self.x0 + self.x1) / 2, (self.y0 + self.y1) / 2
        self.children = [QuadTreeNode(self.x0, self.y0, mx, my), QuadTreeNode(mx, self.y0, self.x1, my),
                         QuadTreeNode(self.x0, my, mx, self.y1), QuadTreeNode(mx, my, self.x1, self.y1)]

    @property
    def area(self):
        return (self.x1 - self.x0) * (self.y1 - self.y0)

    def __str__(self):
        return "({:.1f}, {:.1f}), ({:.1f}, {:.1f})".format(self.x0, self.y0, self.x1, self.y1)


def recursive_subdivide(node, thres, image, cur_depth, max_depth):
    """tree.py:655-676: split while depth < max_depth and the block variance is >= thres."""
    if cur_depth >= max_depth:
        return
    if thres > 0 and node.get_error(image) < thres:     # with thres <= 0 the variance test can never fire
        return
    node.subdivide_once()
    for c in node.children:
        recursive_subdivide(c, thres, image, cur_depth + 1, max_depth)


def get_children(node):
    """Leaves in depth-first order (tree.py:679-686); the index in this list is the leaf id."""
    if not node.children:
        return [node]
    out = []
    for c in node.children:
        out += get_children(c)
    return out


class QuadTree:
    def __init__(self, image, stdThres, max_depth, _boxes=None, _min_area=None):
        self.H, self.W = (image.shape[0], image.shape[1]) if not isinstance(image, tuple) else image
        self.threshold = stdThres
        self.image = None if isinstance(image, tuple) else image
        self.root = QuadTreeNode(0, 0, self.H, self.W)
        if _boxes is None:
            recursive_subdivide(self.root, self.threshold, self.image, 1, max_depth)
            self.minArea = self.H * self.W / (4 ** (max_depth - 1))
        else:
            it = iter([tuple(float(v) for v in b) for b in _boxes])
            self._rebuild(self.root, it, [next(it)])
            self.minArea = _min_area

    @staticmethod
    def _rebuild(node, it, head):
        """Inverse of get_children for a tree produced by midpoint splits: a node is a leaf iff the next DFS leaf is
        exactly its box."""
        if head[0] == node.box():
            head[0] = next(it, None)
            return
        node.subdivide_once()
        for c in node.children:
            QuadTree._rebuild(c, it, head)


class QuadTreeManager:
    """GPU-resident replacement of tree.py:159-566 (constructor and the two methods run_nerf.py calls)."""

    def __init__(self, H, W, K, images, poses, mseThres=0.1, max_depth=5, max_level=None, device=None, seed=0,
                 use_mean=False):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.n_images, self.h, self.w = int(poses.shape[0]), int(H), int(W)
        self.K = np.asarray(K, dtype=np.float64)
        self.images = images
        self.epoch_size = self.n_images * self.h * self.w
        self.processor = None           # the reference's ImageProcessor; here the sharpness maps live on the GPU and are
        self._sharp = None              # computed on first use (only prob=True reads them: tree.py:583-595)
        self._images_dev = torch.as_tensor(images, dtype=torch.float32).to(self.device).contiguous()
        # uint8 image store: images that are exactly the loaders' float32(u / 255.) (load_blender.py:37 on opaque pixels, every
        # LLFF image) are kept as the bytes they were read from -- a quarter of the memory, decoded per batch through a
        # 256-entry table of the same floats.  Anything else (alpha-blended backgrounds, synthetic images) stays fp32.
        self._lut = torch.from_numpy((np.arange(256) / 255.).astype(np.float32)).to(self.device)
        q = torch.round(self._images_dev * 255.0).clamp_(0, 255).to(torch.uint8)
        if self._images_dev.numel() > 0 and torch.equal(self._lut[q.long()], self._images_dev):
            self._images_dev = q
        del q
        self._poses_dev = torch.as_tensor(poses, dtype=torch.float32)[:, :3, :4].to(self.device).contiguous()
        self.max_level = int(max_level) if max_level is not None else int(max_depth) + 6
        # leaf capacity per image.  A leaf is split only if rays fell into it (leaf_max > thres), i.e. if it held at least
        # one pixel (int(area * rays_per_pixel) >= 1, tree.py:581), so a tree never has more than 4*H*W leaves however deep
        # the schedule goes (the reference's defaults n_epoch=12, init_level=3, subdivide_every=1 give max_level 13).
        self.cap = min(4 ** (self.max_level - 1), 4 * self.h * self.w)
        n = self.n_images
        self._boxes = [torch.zeros(n, self.cap, 4, dtype=torch.float64, device=self.device) for _ in range(2)]
        self._count = [torch.zeros(n, dtype=torch.int32, device=self.device) for _ in range(2)]
        self._min_area = torch.zeros(n, dtype=torch.float64, device=self.device)
        self._cur = 0
        self._mirror = None
        if mseThres > 0:
            # variance-driven initial trees (tree.py:84-100,655-676): a one-off host recursion per image, uploaded into
            # the SoA.  Leaves coarser than minArea stay frozen at 10 rays per epoch (tree.py:578-581, 642).
            imgs = images.detach().cpu().numpy() if torch.is_tensor(images) else np.asarray(images)
            self.quadTrees = [QuadTree(imgs[i], float(mseThres), int(max_depth)) for i in range(n)]
        else:
            ops.qt_init(n, self.cap, self.h, self.w, int(max_depth), self._boxes[0], self._count[0], self._min_area)
        self.cur_level = max_depth
        self.leaf_max = torch.full((n * self.cap,), -1.0, dtype=torch.float32, device=self.device)
        # use_mean: the nerf++ / plenoxels copies of the tree refine on the MEAN |gt - pred| of a leaf's rays
        # (nerf++-ours/tree.py:622) instead of the max (tree.py:642): two more tables, folded into leaf_max before refine()
        self.use_mean = bool(use_mean)
        self.leaf_sum = torch.zeros(n * self.cap, dtype=torch.float64, device=self.device) if use_mean else None
        self.leaf_cnt = torch.zeros(n * self.cap, dtype=torch.int32, device=self.device) if use_mean else None
        self._ray_offset = torch.zeros(n * self.cap + 1, dtype=torch.int64, device=self.device)
        self.ray_pix = self.ray_gid = None
        self.n_rays = 0
        self._epoch = 0
        self.seed = int(seed)

    # ------------------------------------------------------------------ GPU state
    def _images_f32(self, img_id=None):
        """The training images as fp32 (decoding the uint8 store)."""
        x = self._images_dev if img_id is None else self._images_dev[img_id]
        return self._lut[x.long()] if x.dtype == torch.uint8 else x

    @property
    def boxes(self):
        return self._boxes[self._cur]

    @property
    def counts(self):
        return self._count[self._cur]

    def leaf_lists(self):
        """[(boxes float64 [n_i,4] numpy, min_area)] per image -- DFS order."""
        cnt = self.counts.cpu().numpy()
        bx = self.boxes.cpu().numpy()
        ma = self._min_area.cpu().numpy()
        return [(bx[i, :cnt[i]].copy(), float(ma[i])) for i in range(self.n_images)]

    # ------------------------------------------------------------------ reference attributes (lazy mirrors)
    @property
    def quadTrees(self) -> List[QuadTree]:
        if self._mirror is None:
            self._mirror = [QuadTree((self.h, self.w), 0.0, 1, _boxes=b, _min_area=m) for b, m in self.leaf_lists()]
        return self._mirror

    @quadTrees.setter
    def quadTrees(self, trees):
        """Resume path (run_nerf.py:339-345): upload pickled trees into the SoA."""
        bx = np.zeros((self.n_images, self.cap, 4), np.float64)
        cnt = np.zeros(self.n_images, np.int32)
        ma = np.zeros(self.n_images, np.float64)
        for i, t in enumerate(trees):
            leaves = get_children(t.root)
            if len(leaves) > self.cap:
                raise FlnerfError("pickled tree %d has %d leaves > capacity %d" % (i, len(leaves), self.cap))
            bx[i, :len(leaves)] = np.array([l.box() for l in leaves], np.float64)
            cnt[i], ma[i] = len(leaves), t.minArea
        self._boxes[self._cur].copy_(torch.from_numpy(bx))
        self._count[self._cur].copy_(torch.from_numpy(cnt))
        self._min_area.copy_(torch.from_numpy(ma))
        self._mirror = list(trees)

    @property
    def childrens(self):
        return [get_children(t.root) for t in self.quadTrees]

    @childrens.setter
    def childrens(self, value):      # derived from quadTrees; the reference assigns it after unpickling
        pass

    @property
    def dirs(self):
        return torch.stack([ops.raygen(self.h, self.w, self.K, self._poses_dev[i])[1] for i in range(self.n_images)], 0)

    @property
    def origins(self):
        return torch.stack([ops.raygen(self.h, self.w, self.K, self._poses_dev[i])[0] for i in range(self.n_images)], 0)

    @property
    def result_leaf_id(self):
        """[N,2] float32 rows (image, leaf) in emission order (tree.py:606)."""
        g = self.ray_gid.long()
        return torch.stack([g // self.cap, g % self.cap], 1).float()

    # ------------------------------------------------------------------ emission
    @property
    def sharp_imgs(self):
        """ImageProcessor.sharp_imgs (image_process.py:24-39) as one [n,H,W] GPU tensor."""
        if self._sharp is None:
            self._sharp = ops.sharp_map(self._images_f32())
        return self._sharp

    def emit_epoch(self, down_scale=1, last_epoch=False, seed=None, prob=False, randSamp_proc=0.95, u=None, shuffle=True):
        """Builds the epoch's shuffled ray index buffer on the GPU; returns the number of rays.  prob=True draws
        int(ray_num*(1-randSamp_proc)) rays of every leaf from its sharpness distribution (tree.py:583-595)."""
        rpp = self.epoch_size / self.n_images / down_scale / self.h / self.w      # tree.py:381-382
        n = self.n_images
        if last_epoch:   # throw-away depth-1 trees: H*W uniform draws per image (tree.py:390-400)
            boxes = torch.zeros(n, self.cap, 4, dtype=torch.float64, device=self.device)
            count = torch.zeros(n, dtype=torch.int32, device=self.device)
            min_area = torch.zeros(n, dtype=torch.float64, device=self.device)
            ops.qt_init(n, self.cap, self.h, self.w, 1, boxes, count, min_area)
        else:
            boxes, count, min_area = self.boxes, self.counts, self._min_area
        ops.qt_count(n, self.cap, boxes, count, min_area, rpp, self._ray_offset)
        self.n_rays = int(self._ray_offset[-1].item())        # one 8-byte D2H per epoch
        # the index buffer keeps its ADDRESS from epoch to epoch (it only ever grows): a training step captured as a CUDA
        # graph has the two pointers baked in
        if getattr(self, "_ray_store", None) is None or self._ray_store.shape[1] < self.n_rays:
            self._ray_store = torch.empty(2, max(self.n_rays, self.epoch_size), dtype=torch.int32, device=self.device)
        self.ray_pix, self.ray_gid = self._ray_store[0, :self.n_rays], self._ray_store[1, :self.n_rays]
        self._epoch += 1
        s = self.seed * 1000003 + self._epoch if seed is None else int(seed)
        if prob:
            tables = ops.qt_prob_prepare(n, self.cap, self.h, self.w, boxes, count, self.sharp_imgs)
            ops.qt_emit_prob(n, self.cap, self.h, self.w, boxes, count, self._ray_offset, self.n_rays, s, float(randSamp_proc),
                             self.sharp_imgs, tables, self.ray_pix, self.ray_gid, u=u, shuffle=shuffle)
        else:
            ops.qt_emit(n, self.cap, self.w, boxes, count, self._ray_offset, self.n_rays, s, self.ray_pix, self.ray_gid)
        self._emitted_last = bool(last_epoch)
        return self.n_rays

    def batch(self, first, B, stride=1):
        """rays_o, rays_d, target_rgb, leaf_gid for rows first, first+stride, ... of the index buffer."""
        return ops.gather_batch(B, first, stride, self.ray_pix, self.ray_gid, self.cap, self.h, self.w, self.K,
                                self._poses_dev, self._images_dev, lut=self._lut)

    def gen_rays_v3_multiThread(self, down_scale=16, prob=True, randSamp_proc=0.95, debug=False, last_epoch=False):
        """tree.py:377-428 -> (origins[N,3], dirs[N,3], rgb[N,3]) (GPU tensors, already shuffled)."""
        n = self.emit_epoch(down_scale, last_epoch, prob=bool(prob), randSamp_proc=randSamp_proc)
        o, d, rgb, _ = ops.gather_batch(n, 0, 1, self.ray_pix, self.ray_gid, self.cap, self.h, self.w, self.K,
                                        self._poses_dev, self._images_dev, want_gid=False, lut=self._lut)
        return o, d, rgb

    def gen_rays_v3(self, down_scale=16, debug=False, last_epoch=False):
        """tree.py:231-307: the sub-pixel variant -- positions on a 1/1000-pixel grid inside every leaf, colours / directions /
        origins interpolated bilinearly (F.grid_sample semantics, including the reference's transposed grid).  Returns
        (origins[N,3], dirs[N,3], rgb[N,3]); result_leaf_id follows.  The positions stay in ``ray_xy`` (emission is shuffled)."""
        rpp = self.epoch_size / self.n_images / down_scale / self.h / self.w
        n = self.n_images
        if last_epoch:
            boxes = torch.zeros(n, self.cap, 4, dtype=torch.float64, device=self.device)
            count = torch.zeros(n, dtype=torch.int32, device=self.device)
            min_area = torch.zeros(n, dtype=torch.float64, device=self.device)
            ops.qt_init(n, self.cap, self.h, self.w, 1, boxes, count, min_area)
        else:
            boxes, count, min_area = self.boxes, self.counts, self._min_area
        ops.qt_count(n, self.cap, boxes, count, min_area, rpp, self._ray_offset)
        self.n_rays = int(self._ray_offset[-1].item())
        self.ray_xy = torch.empty(self.n_rays, 2, dtype=torch.float32, device=self.device)
        self.ray_gid = torch.empty(self.n_rays, dtype=torch.int32, device=self.device)
        self.ray_pix = None
        self._epoch += 1
        ops.qt_emit_sub(n, self.cap, boxes, self._ray_offset, self.n_rays, self.seed * 1000003 + self._epoch, self.ray_xy, self.ray_gid)
        if debug:
            for i in range(n):
                self.visualize_split_and_sample_points(i, self.ray_xy[(self.ray_gid // self.cap) == i])
        return ops.gather_sub(self.ray_xy, self.ray_gid, self.cap, self.h, self.w, self.K, self._poses_dev, self._images_dev, self._lut)

    def gen_rays_v3_1(self, down_scale=16, debug=False, last_epoch=False):
        """tree.py:309-375, the single-thread integer-pixel version: same semantics as gen_rays_v3_multiThread(prob=False)."""
        return self.gen_rays_v3_multiThread(down_scale=down_scale, prob=False, debug=debug, last_epoch=last_epoch)

    def gen_rays_v4(self, sampler_ret, down_scale=1, debug=False):
        """tree.py:430-491 indexes rays pre-generated by tree_utils.RaySampler.pre_gen_rays_v3 -- an offline table the driver
        never builds (run_nerf.py:357-366 is commented out).  On-the-fly emission (emit_epoch) replaces it."""
        raise FlnerfError("gen_rays_v4 needs tree_utils.RaySampler's offline ray table (dead code in the reference driver); "
                          "use gen_rays_v3_multiThread / emit_epoch")

    # ------------------------------------------------------------------ debug pictures (tree.py:148-229)
    def _draw(self, img_id, points=None):
        import cv2
        img = self._images_f32(img_id).detach().cpu().numpy().copy() * 255.0
        boxes, _ = self.leaf_lists()[img_id]
        for x0, y0, x1, y1 in boxes:
            img = cv2.rectangle(img, (int(y0), int(x0)), (int(y1), int(x1)), (0, 0, 0), 1)
        if points is not None:
            for x, y in np.asarray(points.detach().cpu() if torch.is_tensor(points) else points):
                img = cv2.circle(img, (int(y), int(x)), 0, (255, 0, 0), -1)
        return img

    def _write(self, name, img):
        import os
        import cv2
        os.makedirs('debug', exist_ok=True)
        cv2.imwrite(os.path.join('debug', name + '.jpg'), img[:, :, [2, 1, 0]])
        return img

    def visualize_subdivide(self, tree_id=-1, filename_prefix='tree_subdivide'):
        for i in (range(self.n_images) if tree_id == -1 else [tree_id]):
            self._write('tree_subdivide_' + str(i), self._draw(i))

    def visualize_split(self, img_id):
        return self._write('tree_split_{}'.format(img_id), self._draw(img_id))

    def visualize_split_and_sample_points(self, img_id, selected_pixel):
        return self._write('tree_sample_points_{}'.format(img_id), self._draw(img_id, selected_pixel))

    # ------------------------------------------------------------------ refinement
    def reset_leaf_stats(self):
        self.leaf_max.fill_(-1.0)
        if self.use_mean:
            self.leaf_sum.zero_()
            self.leaf_cnt.zero_()

    def accumulate(self, pred, target, leaf_gid):
        """Adds one batch to the refinement statistic (use_mean trees; the max variant is fused into the loss kernel)."""
        if self.use_mean:
            ops.leaf_sum(pred, target, leaf_gid, self.leaf_sum, self.leaf_cnt)
        else:
            ops.mse_leafmax(pred, None, target, max(pred.shape[0], 1), leaf_gid, self.leaf_max, want_grads=False)

    def refine(self, thres):
        """adjust_tree on the accumulated per-leaf table (tree.py:629-652), then clears the table."""
        nxt = 1 - self._cur
        if self.use_mean:
            ops.leaf_mean(self.leaf_sum, self.leaf_cnt, self.leaf_max)
        ops.qt_refine(self.n_images, self.cap, self._boxes[self._cur], self._count[self._cur], self._min_area,
                      self.leaf_max, thres, self._boxes[nxt], self._count[nxt])
        self._cur = nxt
        self._mirror = None
        self.cur_level += 1
        self.reset_leaf_stats()

    def adjust_tree(self, rgb_gt, rgb_pred, thres=0.01, debug=False):
        """tree.py:493-531, the single-thread variant: it splits on the MEAN leaf loss (:515), unlike the multi-thread one the
        driver calls (max, :642)."""
        gt = torch.as_tensor(rgb_gt, dtype=torch.float32).to(self.device)
        pr = torch.as_tensor(rgb_pred, dtype=torch.float32).to(self.device)
        n = gt.shape[0]
        keep = self.use_mean, self.leaf_sum, self.leaf_cnt
        self.use_mean = True
        if self.leaf_sum is None:
            self.leaf_sum = torch.zeros(self.n_images * self.cap, dtype=torch.float64, device=self.device)
            self.leaf_cnt = torch.zeros(self.n_images * self.cap, dtype=torch.int32, device=self.device)
        try:
            self.reset_leaf_stats()
            self.accumulate(pr, gt, self.ray_gid[:n].contiguous())
            self.refine(thres)
        finally:
            self.use_mean = keep[0]
        print('After sudivide, there are {} child nodes'.format(int(self.counts.sum().item())))

    def adjust_tree_multiThread(self, rgb_gt, rgb_pred, thres=0.001, debug=False):
        """tree.py:533-557: rgb_gt / rgb_pred are the epoch's [N,3] targets and predictions in emission order."""
        gt = torch.as_tensor(rgb_gt, dtype=torch.float32).to(self.device)
        pr = torch.as_tensor(rgb_pred, dtype=torch.float32).to(self.device)
        n = gt.shape[0]
        self.reset_leaf_stats()
        self.accumulate(pr, gt, self.ray_gid[:n].contiguous())
        self.refine(thres)
        print('After sudivide, there are {} child nodes'.format(int(self.counts.sum().item())))
