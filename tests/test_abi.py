"""The C-ABI library loads without a GPU and exports every symbol include/wsdl_b200.h declares; host-side
argument validation works before anything touches the device.  CPU only."""
import ctypes
import os
import re

import pytest
import torch

import weaklysuperviseddl_b200 as W
from weaklysuperviseddl_b200 import _native
from weaklysuperviseddl_b200 import build as wbuild
from weaklysuperviseddl_b200 import functional as WF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    wbuild.build()  # no-op when up to date; nvcc cross-compiles sm_100a without a GPU
    return _native.lib()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "wsdl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wsdl_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound(lib):
    syms = declared_symbols()
    assert len(syms) >= 11
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/wsdl_b200.h but not exported"
        assert s in _native.SIGNATURES, f"{s} has no ctypes signature in _native.py"
    assert set(_native.SIGNATURES) == set(syms)


def test_version_and_strerror(lib):
    assert lib.wsdl_version() == 200
    assert lib.wsdl_strerror(0) == b"ok"
    for rc in range(-6, 0):
        assert lib.wsdl_strerror(rc).startswith(b"wsdl:")
    assert b"unknown" in lib.wsdl_strerror(-99)


def test_workspace_queries_are_host_only(lib):
    Int2 = ctypes.c_int * 2
    n = lib.wsdl_layercam_workspace_bytes(Int2(1024, 2048), Int2(32, 32), Int2(32, 32), 2, 16, 0)
    assert n >= 16 * 2 * 32 * 32 * 4
    assert lib.wsdl_layercam_workspace_bytes(Int2(1024, 2048), Int2(32, 0), Int2(32, 32), 2, 16, 0) == 0
    assert lib.wsdl_layercam_workspace_bytes(Int2(1024, 2048), Int2(32, 32), Int2(32, 32), 2, 16, 7) == 0
    assert lib.wsdl_pairwise_workspace_bytes(32, 224, 224) >= 32 * 49 * 4
    assert lib.wsdl_pairwise_workspace_bytes(0, 224, 224) == 0
    n_ccl = lib.wsdl_keep_largest_workspace_bytes(4, 512, 512)  # per-tile component tables: 4 KB + 256 B per 32x32 tile
    assert 4 * 256 * (4 * 1024 + 256) <= n_ccl < 4 * 512 * 512 * 8  # round 1 kept 8 bytes per PIXEL
    assert lib.wsdl_weak_loss_workspace_bytes(32, 224, 224) >= 512 and lib.wsdl_weak_loss_workspace_bytes(0, 1, 1) == 0
    assert lib.wsdl_refine_workspace_bytes(16, 256, 256) >= 16 * 4 and lib.wsdl_refine_workspace_bytes(1, 0, 5) == 0


def test_argument_errors_do_not_touch_the_device(lib):
    # NULL pointers / bad extents are rejected before any CUDA call
    assert lib.wsdl_pairwise_fwd_bwd(None, None, 1, 2, 8, 8, 5, 0.05, 0.0, 1, 1, 0, None, None, None, None, 0, None) == -1
    assert lib.wsdl_affinities(None, 1, 8, 8, 5, 0.1, 5.0, None, None) == -1
    assert lib.wsdl_keep_largest(None, 1, 8, 8, None, None, None, 0, None) == -1
    assert lib.wsdl_threshold_mask(None, 10, 0.3, 1e-6, None, None, None) == -1
    assert lib.wsdl_scale(None, None, 10, None, 0, None) == -1
    fake = ctypes.c_void_p(256)  # never dereferenced: shape checks come first
    assert lib.wsdl_pairwise_fwd_bwd(fake, fake, 1, 9, 8, 8, 5, 0.05, 0.0, 1, 1, 0, None, fake, None, fake, 1 << 20, None) == -2
    assert lib.wsdl_pairwise_fwd_bwd(fake, fake, 1, 2, 8, 8, 4, 0.05, 0.0, 1, 1, 0, None, fake, None, fake, 1 << 20, None) == -2
    assert lib.wsdl_pairwise_fwd_bwd(fake, fake, 1, 2, 2, 8, 5, 0.05, 0.0, 1, 1, 0, None, fake, None, fake, 1 << 20, None) == -2
    assert lib.wsdl_pairwise_fwd_bwd(fake, fake, 1, 2, 8, 8, 5, 0.0, 0.0, 1, 1, 0, None, fake, None, fake, 1 << 20, None) == -6
    assert lib.wsdl_pairwise_fwd_bwd(fake, fake, 1, 2, 8, 8, 5, 0.05, 0.0, 1, 1, 0, None, fake, None, fake, 8, None) == -4


def test_no_cpu_fallback():
    """CPU tensors are refused loudly: the product path never routes through torch eager or the oracle."""
    with pytest.raises(_native.WsdlError):
        WF.layercam_fused([torch.zeros(1, 4, 2, 2)], [torch.zeros(1, 4, 2, 2)])
    with pytest.raises(_native.WsdlError):
        W.LocalNormalizedCutLoss()(torch.zeros(1, 2, 8, 8), torch.zeros(1, 3, 8, 8))
    with pytest.raises(_native.WsdlError):
        W.ConstrainToBoundaryLossSingle()(torch.zeros(2, 8, 8), torch.zeros(3, 8, 8))
    with pytest.raises(_native.WsdlError):
        WF.keep_largest(torch.zeros(4, 4, dtype=torch.uint8))
    with pytest.raises(_native.WsdlError):
        WF.affinities(torch.zeros(1, 3, 8, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "weaklysuperviseddl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle/"
                assert "wsdl_oracle" not in src


def test_reference_signatures_are_kept():
    import inspect

    sig = inspect.signature(W.LayerCAMGenerator.__init__)
    assert list(sig.parameters)[:3] == ["self", "model", "target_layer_names"]
    assert sig.parameters["target_layer_names"].default == ["layer3", "layer4"]
    sig = inspect.signature(W.LayerCAMGenerator.generate)
    assert list(sig.parameters) == ["self", "images", "alpha", "class_idx"]
    from weaklysuperviseddl_b200.AlternatingDirectionCutLoss import LayerCAMGenerator as Variant

    assert list(inspect.signature(Variant.generate).parameters) == ["self", "images", "class_idx", "alpha"]
    sig = inspect.signature(W.generate_pseudo_masks)
    assert list(sig.parameters)[:6] == ["loader", "layercam_gen", "cam_thresh", "alpha", "keep_largest_masks", "run_id"]
    assert [sig.parameters[k].default for k in ("cam_thresh", "alpha", "keep_largest_masks", "run_id")] == [0.3, 1.0, True, "default"]
    sig = inspect.signature(W.LocalNormalizedCutLoss.__init__)
    assert [sig.parameters[k].default for k in ("sigma_color", "window_size")] == [0.05, 5]
    sig = inspect.signature(W.ConstrainToBoundaryLossSingle.__init__)
    assert [sig.parameters[k].default for k in ("sigma_color", "sigma_space", "window_size", "eps")] == [0.1, 5, 5, 1e-8]
    sig = inspect.signature(W.compute_affinities)
    assert [sig.parameters[k].default for k in ("sigma_color", "sigma_space", "window_size")] == [0.1, 5, 5]
    sig = inspect.signature(W.refine_pseudo_mask)
    assert [sig.parameters[k].default for k in ("lambda_boundary", "threshold", "lr", "num_steps", "sigma_color", "window_size")] == [0.1, 0.5, 1e-2, 20, 0.1, 5]
    sig = inspect.signature(W.evaluate_layercam_on_test_set)
    assert [sig.parameters[k].default for k in ("alpha", "cam_thresh")] == [1.0, 0.3]
