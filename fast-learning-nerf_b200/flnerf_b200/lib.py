"""ctypes binding of libflnerf.so (the C ABI declared in include/flnerf.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.  The product
path never imports anything from ``oracle/``.
"""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libflnerf.so")

MODE_FP32 = 0
MODE_BF16 = 1
MODE_BF16X3 = 2
MLP_PARAMS = 595844

_vp, _i, _i64, _u64, _f, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); mirrors include/flnerf.h one to one
SIGNATURES = {
    "flnerf_version": (_i, []),
    "flnerf_last_error": (C.c_char_p, []),
    "flnerf_create": (_vp, [_i]),
    "flnerf_destroy": (None, [_vp]),
    "flnerf_sm_count": (_i, [_vp]),
    "flnerf_raygen": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "flnerf_pack_rays": (_i, [_vp, _i64, _vp, _vp, _f, _f, _i, _i, _i, _d, _vp, _vp]),
    "flnerf_coarse_depths": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _i, _i, _u64, _u64, _vp, _vp]),
    "flnerf_posenc": (_i, [_vp, _i64, _i, _vp, _vp, _vp]),
    "flnerf_encode_f32": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _vp]),
    "flnerf_encode_tc": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _vp, _vp]),
    "flnerf_pack_x90": (_i, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "flnerf_encode_tc_x3": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _vp, _vp]),
    "flnerf_pack_x90_x3": (_i, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "flnerf_encode_frame_tc": (_i, [_vp, _i, _i, _i, _vp, _vp, _f, _f, _i, _i, _i64, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "flnerf_ssim_psnr": (_i, [_vp, _i, _i, _vp, _vp, _d, _vp, _vp]),
    "flnerf_padded_rows": (_i64, [_i64]),
    "flnerf_mlp_stash_bytes": (_sz, [_i, _i64, _i, _i]),
    "flnerf_mlp_bwd_workspace_bytes": (_sz, [_i, _i64]),
    "flnerf_mlp_packed_bytes": (_sz, []),
    "flnerf_mlp_pack_weights": (_i, [_vp, _vp, _vp, _vp]),
    "flnerf_mlp_forward": (_i, [_vp, _i, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "flnerf_mlp_backward": (_i, [_vp, _i, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "flnerf_mlp_backward_stages": (_i, [_vp, _i, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "flnerf_composite_forward": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _i64, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "flnerf_composite_backward": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _i64, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "flnerf_sample_pdf_merge": (_i, [_vp, _i64, _i, _i, _vp, _vp, _vp, _i, _u64, _u64, _vp, _vp, _vp, _vp]),
    "flnerf_sample_pdf": (_i, [_vp, _i64, _i, _i, _vp, _vp, _vp, _i, _u64, _u64, _vp, _vp]),
    "flnerf_mse_leafmax": (_i, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "flnerf_leaf_sum": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "flnerf_leaf_mean": (_i, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "flnerf_adam_step": (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _d, _d, _d, _d, _i64, _vp]),
    "flnerf_qt_init": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "flnerf_qt_refine": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp]),
    "flnerf_qt_count": (_i, [_vp, _i, _i, _vp, _vp, _vp, _d, _vp, _vp]),
    "flnerf_qt_emit": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _i64, _u64, _vp, _vp, _vp]),
    "flnerf_sharp_map": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "flnerf_qt_prob_rows": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp]),
    "flnerf_qt_prob_prepare": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "flnerf_qt_emit_prob": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _i64, _u64, _d, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp,
                                 _vp]),
    "flnerf_gather_batch": (_i, [_vp, _i64, _i64, _i64, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "flnerf_gather_batch_u8": (_i, [_vp, _i64, _i64, _i64, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "flnerf_qt_emit_sub": (_i, [_vp, _i, _i, _vp, _vp, _i64, _u64, _vp, _vp, _vp]),
    "flnerf_gather_sub": (_i, [_vp, _i64, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "flnerf_mlp_packed_bytes_g": (_sz, [_i]),
    "flnerf_mlp_pack_weights_g": (_i, [_vp, _i, _vp, _vp, _vp]),
    "flnerf_mlp_forward_g": (_i, [_vp, _i, _i, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "flnerf_mlp_backward_g": (_i, [_vp, _i, _i, _vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "flnerf_pack_xrows": (_i, [_vp, _i, _i, _i, _i64, _i, _vp, _vp, _vp, _vp]),
    "flnerf_mlp_param_count_g": (_i64, [_i, _i]),
    "flnerf_mlp_fp32_forward_g": (_i, [_vp, _i, _i, _vp, _i64, _vp, _vp, _vp, _vp]),
    "flnerf_mlp_fp32_backward_g": (_i, [_vp, _i, _i, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "flnerf_pp_depths0": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _vp, _i, _u64, _u64, _vp, _vp, _vp, _vp]),
    "flnerf_pp_bg_encode": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "flnerf_pp_composite_forward": (_i, [_vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "flnerf_pp_composite_backward": (_i, [_vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "flnerf_pp_sample_pdf_merge": (_i, [_vp, _i64, _i, _i, _vp, _vp, _vp, _i, _u64, _u64, _vp, _vp, _vp]),
    "flnerf_launch_count": (_i64, [_i]),
    "flnerf_launch_count_add": (None, [_i64]),
    "flnerf_set_step_record": (_i, [_vp, _vp]),
    "flnerf_step_record_write": (_i, [_vp, _vp, _i64, _u64, _d, _d, _d, _i64, _vp]),
}

_lib = None
_lock = threading.Lock()
_ctx = {}


class FlnerfError(RuntimeError):
    pass


def load():
    """Loads libflnerf.so and types every exported symbol.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise FlnerfError(
                "libflnerf.so not found at %s -- build it with fast-learning-nerf_b200/csrc/build.sh "
                "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().flnerf_last_error().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        raise FlnerfError("%s failed (%d): %s" % (what, rc, last_error()))


def context(device_index: int):
    """One flnerf_ctx per CUDA device (created on first use)."""
    lib = load()
    if device_index not in _ctx:
        h = lib.flnerf_create(int(device_index))
        if not h:
            raise FlnerfError("flnerf_create(%d) failed: %s" % (device_index, last_error()))
        _ctx[device_index] = h
    return _ctx[device_index]


def launch_count(reset: bool = False) -> int:
    return int(load().flnerf_launch_count(1 if reset else 0))
