"""Drop-in for nerf-ours/argument_parser.py: ``config_parser()`` with every flag and default of the reference
(argument_parser.py:7-121).  configargparse is not a dependency here: a small argparse subclass reads the
same ``key = value`` config files (``--config``); command-line flags override the file, as in configargparse.
"""
import argparse

# (flag, kind, default, help).  kind: a type for valued flags, 'flag' for store_true switches.
_FLAGS = [
    ("expname", str, None, "experiment name"),
    ("basedir", str, "./logs/", "where to store ckpts and logs"),
    ("datadir", str, "./data/llff/fern", "input data directory"),
    # network / optimisation
    ("netdepth", int, 8, "layers in network"),
    ("netwidth", int, 256, "channels per layer"),
    ("netdepth_fine", int, 8, "layers in fine network"),
    ("netwidth_fine", int, 256, "channels per layer in fine network"),
    ("N_rand", int, 32 * 32 * 4, "batch size (number of random rays per gradient step)"),
    ("lrate", float, 5e-4, "learning rate"),
    ("lrate_decay", int, 250, "exponential learning rate decay (in 1000 steps)"),
    ("chunk", int, 1024 * 32, "number of rays processed in parallel"),
    ("netchunk", int, 1024 * 64, "number of pts sent through network in parallel"),
    ("no_batching", "flag", False, "only take random rays from 1 image at a time"),
    ("no_reload", "flag", False, "do not reload weights from saved ckpt"),
    ("ft_path", str, None, "specific weights file to reload"),
    # rendering
    ("N_samples", int, 64, "number of coarse samples per ray"),
    ("N_importance", int, 0, "number of additional fine samples per ray"),
    ("perturb", float, 1., "set to 0. for no jitter, 1. for jitter"),
    ("use_viewdirs", "flag", False, "use full 5D input instead of 3D"),
    ("i_embed", int, 0, "set 0 for default positional encoding, -1 for none"),
    ("multires", int, 10, "log2 of max freq for positional encoding (3D location)"),
    ("multires_views", int, 4, "log2 of max freq for positional encoding (2D direction)"),
    ("raw_noise_std", float, 0., "std dev of noise added to regularize sigma_a output"),
    ("render_only", "flag", False, "do not optimize, reload weights and render out render_poses path"),
    ("render_test", "flag", False, "render the test set instead of render_poses path"),
    ("render_factor", int, 0, "downsampling factor to speed up rendering"),
    ("precrop_iters", int, 0, "number of steps to train on central crops"),
    ("precrop_frac", float, .5, "fraction of img taken for central crops"),
    # quadtree ray selector ("ours")
    ("n_epoch", int, 12, "number of total epoch"),
    ("init_level", int, 3, "init quadtree subdivide level"),
    ("rays_downscale", int, 1, ""),
    ("subdivide_every", int, 1, "subdivide quadtrees every x epochs"),
    ("subdivide_thres", float, 0.015, ""),
    ("randSamp_perc", float, 0.5, ""),
    ("dset_name", str, "Truck", ""),
    ("end_rand", int, 11, "start to add color sampling from this epoch"),
    # datasets
    ("dataset_type", str, "llff", "options: llff / blender / deepvoxels"),
    ("testskip", int, 8, "will load 1/N images from test/val sets"),
    ("shape", str, "greek", "options : armchair / cube / greek / vase"),
    ("white_bkgd", "flag", False, "render synthetic data on a white bkgd"),
    ("half_res", "flag", False, "load blender synthetic data at 400x400 instead of 800x800"),
    ("factor", int, 8, "downsample factor for LLFF images"),
    ("no_ndc", "flag", False, "do not use normalized device coordinates"),
    ("lindisp", "flag", False, "sampling linearly in disparity rather than depth"),
    ("spherify", "flag", False, "set for spherical 360 scenes"),
    ("llffhold", int, 8, "will take every 1/N images as LLFF test set"),
    # logging
    ("i_print", int, 100, "frequency of console printout"),
    ("i_img", int, 500, "frequency of image logging"),
    ("i_weights", int, 10000, "frequency of weight ckpt saving"),
    ("i_testset", int, 50000, "frequency of testset saving"),
    ("i_video", int, 50000, "frequency of render_poses video saving"),
    # flnerf additions (not in the reference)
    ("precision", str, None, "MLP arithmetic: bf16 (tcgen05 throughput mode), bf16x3 (split-precision tcgen05, meets the "
                             "reference's 1e-4 tolerance) or fp32 (CUDA-core parity path); default $FLNERF_PRECISION or bf16"),
    ("no_graph", "flag", False, "launch every training step kernel by kernel instead of replaying one CUDA graph"),
]


class ConfigFileParser(argparse.ArgumentParser):
    """argparse + ``--config file`` of ``key = value`` lines ('#' comments; ``key = True`` sets a switch)."""

    def parse_known_args(self, args=None, namespace=None):
        import sys
        argv = list(sys.argv[1:] if args is None else args)
        file_args = []
        for i, a in enumerate(argv):
            path = argv[i + 1] if a == "--config" and i + 1 < len(argv) else (a.split("=", 1)[1] if a.startswith("--config=") else None)
            if path:
                file_args += self._read(path)
        return super().parse_known_args(file_args + argv, namespace)

    @staticmethod
    def _read(path):
        out = []
        with open(path) as f:
            for line in f:
                line = line.split("#", 1)[0].strip()
                if not line or "=" not in line:
                    continue
                k, v = [s.strip() for s in line.split("=", 1)]
                if v.lower() == "true":
                    out.append("--" + k)
                elif v.lower() not in ("false", ""):
                    out += ["--" + k, v]
        return out


def config_parser():
    p = ConfigFileParser()
    p.add_argument("--config", type=str, default=None, help="config file path")
    for name, kind, default, text in _FLAGS:
        if kind == "flag":
            p.add_argument("--" + name, action="store_true", help=text)
        else:
            p.add_argument("--" + name, type=kind, default=default, help=text)
    return p
