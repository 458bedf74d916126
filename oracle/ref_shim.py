"""TEST INFRASTRUCTURE -- loads the UNMODIFIED reference modules from /root/reference, or from the untracked copy
``baseline/_ref/`` that ``tools/install_reference.sh`` makes of it (git-ignored, travels to the GPU box with the snapshot).

It is used by ``oracle/make_golden.py`` to generate the committed fixtures under
``tests/golden/``, by ``tests/test_oracle_vs_reference.py`` (skipped when no reference tree is
present) to pin the oracle restatement against the real code, and by ``bench.py --impl reference`` /
``tools/psnr_check.py`` to TIME and TRAIN the reference's own code beside ours.

The reference hot path (nerf-ours/{run_nerf_helpers,render,model,tree}.py) imports a
few modules that are not installed here (imageio, matplotlib, colour, threadpool,
``from cv2 import cv2``) and hard-codes ``.cuda()``; we stub those *modules* and make
``Tensor.cuda`` the identity on a GPU-less host.  No reference file is edited or copied.
"""
import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    for cand in (os.environ.get("FLNERF_REFERENCE_ROOT"), "/root/reference",
                 os.path.join(os.path.dirname(_HERE), "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "nerf-ours", "render.py")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()
REF_NERF = os.path.join(REF_ROOT, "nerf-ours")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_NERF, "render.py"))


class _SerialPool:
    """Serial stand-in for threadpool.ThreadPool (tree.py:410-414, 547-550)."""

    def __init__(self, n):
        self.q = []

    def putRequest(self, r):
        self.q.append(r)

    def wait(self):
        q, self.q = self.q, []
        for fn, kw in q:
            fn(**kw)


def load():
    """Returns a namespace with the reference modules: helpers, model, render, tree."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_NERF)
    import torch

    for name in ["imageio", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "colour"]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["colour"].Color = object
    try:
        import cv2
        cv2.cv2 = cv2
        sys.modules["cv2.cv2"] = cv2
    except Exception:  # pragma: no cover
        pass
    if "threadpool" not in sys.modules:
        tp = types.ModuleType("threadpool")
        tp.ThreadPool = _SerialPool
        tp.makeRequests = lambda fn, vs: [(fn, kw) for _, kw in vs]
        sys.modules["threadpool"] = tp
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self

    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in
                  ["run_nerf_helpers", "model", "render", "tree", "image_process", "tree_utils"]}
    for k in saved_mods:
        sys.modules.pop(k, None)
    sys.path.insert(0, REF_NERF)
    try:
        ns = types.SimpleNamespace()
        ns.helpers = importlib.import_module("run_nerf_helpers")
        ns.model = importlib.import_module("model")
        ns.render = importlib.import_module("render")
        ns.tree = importlib.import_module("tree")
        ns.image_process = importlib.import_module("image_process")
    finally:
        sys.path[:] = saved_path
        # keep the reference modules out of the global namespace so that the product's
        # same-named modules (fast-learning-nerf_b200/render.py ...) can be imported later
        for k, v in saved_mods.items():
            cur = sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
            setattr(ns, "_" + k, cur)
    return ns


def _configargparse_stub():
    """``configargparse`` is not installed: an argparse subclass covering what argument_parser.py:4-123 uses of it --
    ``add_argument(..., is_config_file=True)`` and ``key = value`` config files (command-line flags win)."""
    import argparse

    class ArgumentParser(argparse.ArgumentParser):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self._cfg_dests = []

        def add_argument(self, *names, **kw):
            is_cfg = kw.pop("is_config_file", False)
            act = super().add_argument(*names, **kw)
            if is_cfg:
                self._cfg_dests.append(act.dest)
            return act

        def parse_args(self, args=None, namespace=None):
            args = list(sys.argv[1:] if args is None else args)
            pre, _ = super().parse_known_args(args)
            extra = []
            for dest in self._cfg_dests:
                path = getattr(pre, dest, None)
                if not path:
                    continue
                for line in open(path):
                    line = line.split("#", 1)[0].strip()
                    if "=" not in line:
                        continue
                    k, v = [t.strip() for t in line.split("=", 1)]
                    act = next((a for a in self._actions if a.dest == k), None)
                    if act is None:
                        continue
                    if act.nargs == 0:                      # store_true style flag
                        if v.lower() in ("true", "1", "yes"):
                            extra.append("--" + k)
                    else:
                        extra += ["--" + k, v]
            return super().parse_args(extra + args, namespace)

    m = types.ModuleType("configargparse")
    m.ArgumentParser = ArgumentParser
    return m


def load_run_nerf():
    """The reference's driver module nerf-ours/run_nerf.py, imported UNMODIFIED (its create_nerf / run_network / render are
    the stock code path ``bench.py --impl reference`` and tools/psnr_check.py time and train).  Returns (ns, run_nerf)."""
    ns = load()
    if "configargparse" not in sys.modules:
        sys.modules["configargparse"] = _configargparse_stub()
    names = ["run_nerf_helpers", "model", "render", "tree", "image_process", "tree_utils", "argument_parser", "run_nerf",
             "load_llff", "load_deepvoxels", "load_blender", "load_LINEMOD"]
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in names}
    for k in names:
        sys.modules.pop(k, None)
    # the same module objects load() imported, so that run_nerf's star-imports see them
    for k in ("run_nerf_helpers", "model", "render", "tree", "image_process"):
        m = getattr(ns, "_" + k, None)
        if m is not None:
            sys.modules[k] = m
    sys.path.insert(0, REF_NERF)
    try:
        rn = importlib.import_module("run_nerf")
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
    return ns, rn


def ref_run_network(ns, inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """run_network is defined in run_nerf.py (50-64), which cannot be imported without
    configargparse/imageio/loaders; this is its 12-line call sequence expressed over the
    reference's own embedders and model (no new arithmetic)."""
    import torch
    flat = inputs.reshape(-1, inputs.shape[-1])
    emb = embed_fn(flat)
    if viewdirs is not None:
        dirs = viewdirs[:, None].expand(inputs.shape).reshape(-1, inputs.shape[-1])
        emb = torch.cat([emb, embeddirs_fn(dirs)], -1)
    out = torch.cat([fn(emb[i:i + netchunk]) for i in range(0, emb.shape[0], netchunk)], 0)
    return out.reshape(list(inputs.shape[:-1]) + [out.shape[-1]])


REF_NERFPP = os.path.join(os.path.dirname(REF_NERF), "nerf++-ours")


def load_nerfpp():
    """The reference's nerf++ fork: nerf_network (Embedder, MLPNet), ddp_model (depth2pts_outside,
    NerfNet) and the sampling helpers of ddp_train_nerf (intersect_sphere, perturb_samples, sample_pdf), imported
    unmodified behind the same kind of stubs as load()."""
    if not os.path.isdir(REF_NERFPP):
        raise RuntimeError("reference tree not present at %s" % REF_NERFPP)
    load()      # installs the imageio / matplotlib / colour / cv2.cv2 / threadpool stubs
    mpl = sys.modules["matplotlib"]
    mpl.__path__ = []           # make the stub a package so that its sub-modules can be stubbed too
    for name, attrs in [("matplotlib.backends", {}), ("matplotlib.backends.backend_agg", {"FigureCanvasAgg": object}),
                        ("matplotlib.figure", {"Figure": object}), ("matplotlib.cm", {})]:
        m = sys.modules.get(name) or types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
    mpl.cm = sys.modules["matplotlib.cm"]
    names = ["utils", "nerf_network", "ddp_model", "nerf_sample_ray_split", "data_loader_split", "image_process", "tree",
             "tree_utils", "ddp_train_nerf"]
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in names}
    for k in names:
        sys.modules.pop(k, None)
    sys.path.insert(0, REF_NERFPP)
    try:
        ns = types.SimpleNamespace()
        ns.network = importlib.import_module("nerf_network")
        ns.model = importlib.import_module("ddp_model")
        ns.train = importlib.import_module("ddp_train_nerf")
        ns.tree = sys.modules["tree"]              # nerf++-ours/tree.py (the .mean() refinement variant)
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v
    return ns
