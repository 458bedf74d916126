"""CPU: libflnerf.so loads and exports exactly the symbols include/flnerf.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "flnerf.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(flnerf_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_path():
    syms = header_symbols()
    for need in ["flnerf_raygen", "flnerf_pack_rays", "flnerf_coarse_depths", "flnerf_encode_tc", "flnerf_mlp_forward",
                 "flnerf_mlp_backward", "flnerf_composite_forward", "flnerf_composite_backward",
                 "flnerf_sample_pdf_merge", "flnerf_mse_leafmax", "flnerf_adam_step", "flnerf_qt_refine",
                 "flnerf_qt_emit"]:
        assert need in syms


def test_library_exports_every_declared_symbol():
    from flnerf_b200 import lib as L
    assert os.path.isfile(L.LIB_PATH), "libflnerf.so missing: run __graft_entry__.build()"
    so = ctypes.CDLL(L.LIB_PATH)
    for s in header_symbols():
        assert hasattr(so, s), "header declares %s but the library does not export it" % s
    assert set(L.SIGNATURES) == set(header_symbols())          # the ctypes table mirrors the header one to one
    lib = L.load()
    assert lib.flnerf_version() >= 100
    assert lib.flnerf_padded_rows(1) == 256 and lib.flnerf_padded_rows(512) == 512
    # hi part + lo part (bf16x3), each = 38 forward chunks + 34 transposed dgrad chunks
    assert lib.flnerf_mlp_packed_bytes() == 2 * ((34 * 32768 + 4 * 16384) + 34 * 32768)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from flnerf_b200 import ops
    from flnerf_b200.lib import FlnerfError
    with pytest.raises(FlnerfError):
        ops.posenc(torch.zeros(4, 3), 10)                # CPU tensor -> loud failure, not a torch fallback
    import model
    net = model.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    with pytest.raises(FlnerfError):
        net(torch.zeros(2, 90))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fast-learning-nerf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "nerf_oracle" not in src and "ref_shim" not in src and "import oracle" not in src, f
