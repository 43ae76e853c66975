"""CPU tests of the drop-in boundary: librr_b200.so loads without a GPU, exports every symbol include/rgbd_recon_b200.h
declares (and nothing in the header is missing from the ctypes binding), and fails loudly — never falls back — when no
CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rgbd_recon_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rr_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("rr_create", "rr_destroy", "rr_calib_upload", "rr_calib_upload_inv", "rr_calib_invert", "rr_configure",
                 "rr_upload_frames", "rr_preprocess", "rr_bricks_clear", "rr_bricks_update", "rr_integrate", "rr_raymarch",
                 "rr_download_tsdf", "rr_last_error"):
        assert must in syms
    assert len(syms) >= 30


def test_library_exports_every_declared_symbol():
    from rrpy import capi
    assert os.path.exists(capi.LIB_PATH), "build first: make -C rgbd-recon_b200 (or __graft_entry__.build())"
    L = C.CDLL(capi.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert L.rr_version() >= 100


def test_binding_covers_the_header():
    from rrpy import capi
    L = capi.lib()
    unbound = [s for s in declared_symbols() if getattr(L, s).argtypes is None and s not in ("rr_version",)]
    assert not unbound, f"ctypes binding lacks argtypes for: {unbound}"


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rrpy import capi
    with pytest.raises(capi.RRError):
        capi.Fusion(1, 64, 53, 80, 68)
    h = C.c_void_p()
    assert capi.lib().rr_create(C.byref(h), 0, 1, 64, 53, 80, 68) == -3      # RR_ERR_NO_DEVICE
    assert not h.value


def test_product_does_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may touch oracle/."""
    pkg = os.path.join(ROOT, "rgbd-recon_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_py" not in txt and "librr_oracle" not in txt and "ro_" + "integrate" not in txt, f"{f} references the oracle"
