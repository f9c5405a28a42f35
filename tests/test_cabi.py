"""CPU-side checks of the drop-in boundary: the shared library loads, exports
every symbol include/nbmf_cuda.h declares, and refuses to compute without a GPU
(no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from nbmf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "nbmf_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(nbmf_[a-z0-9_]+)\s*\(", src, flags=re.I)
    return sorted({n for n in names if n not in ("nbmf_allreduce_fn",)})


def test_header_symbols_are_exported_and_bound():
    lib = _lib.load()
    decl = declared_symbols()
    assert len(decl) >= 25
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/nbmf_cuda.h but not exported"
    # the ctypes table covers the header exactly
    assert sorted(_lib.SYMBOLS) == decl


def test_version_and_no_fortran_names():
    lib = _lib.load()
    assert b"sm_100a" in lib.nbmf_version()


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.nbmf_device_count() == 0
    h = C.c_void_p()
    rc = lib.nbmf_create(C.byref(h), None)
    assert rc == -2 and not h  # NBMF_ERR_CUDA, no handle, no fallback
    from nbmf_b200 import Engine, NbmfError
    with pytest.raises(NbmfError):
        Engine()


def test_product_package_never_uses_oracle():
    """The oracle is test infrastructure: nothing shipped may import, include or link it."""
    pkg = os.path.join(ROOT, "nbmf_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if not fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                continue
            txt = open(os.path.join(dirpath, fn), errors="ignore").read()
            assert "import oracle" not in txt and "from oracle" not in txt, fn
            assert "oracle/" not in txt and "nbmf_oracle" not in txt and "libnbmf_ref" not in txt, fn
    # the development tools and the reference-side glue do not use it either (the checker tools live in tests/tools/)
    for sub in ("tools", "integration", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                    assert "import oracle" not in txt and "from oracle" not in txt and "nbmf_oracle" not in txt, (sub, fn)


def test_glue_sources_use_only_declared_entry_points():
    """integration/nbmf_glue.cpp (the bodies for the reference's Rcpp functions) and the snippets of
    INTEGRATION.md call nothing but what include/nbmf_cuda.h declares."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    declared = set(re.findall(r"\b(nbmf_[a-z0-9_]+)\s*\(", open(os.path.join(root, "include", "nbmf_cuda.h")).read()))
    types = {"nbmf_handle", "nbmf_opts", "nbmf_timing", "nbmf_allreduce_fn", "nbmf_precision", "nbmf_dirber_mode"}
    for rel in ("integration/nbmf_glue.cpp", "integration/nbmf_glue.h", "INTEGRATION.md"):
        text = open(os.path.join(root, rel)).read()
        used = set(re.findall(r"\b(nbmf_[a-z0-9_]+)\s*\(", text)) - types
        used = {u for u in used if not u.startswith("nbmf_glue") and not u.startswith("nbmf_b200")}
        assert used <= declared, (rel, sorted(used - declared))


def test_auto_precision_horizons_are_below_the_measured_ones():
    """Host logic of NBMF_PREC_AUTO (tensor_pass.cu:tp_horizon): the planned-update horizon of each f16 variant, as a
    function of the reduction length, stays below what profiles/r2_error_growth_*.log measured (hi only 16 / 23 / 31
    updates at 16384^2 / 65536^2 / 262144^2, hi + lo 64 / 39 / 57), shrinks for short reductions and never exceeds the
    reference's default iter = 100 (collapsed_gibbs_z_VB.cpp:376), so a default-length run is always FP64."""
    from nbmf_b200 import _lib
    lib = _lib.load()
    hz = lambda path, n: lib.nbmf_debug_model_error(path, float(n), 0)
    measured = {_lib.PATH_TENSOR_FAST: {16384: 16, 65536: 23, 262144: 31}, _lib.PATH_TENSOR_LO8: {16384: 64, 65536: 39, 262144: 57}}
    for path, pts in measured.items():
        for n, m in pts.items():
            assert 0 < hz(path, n) <= m, (path, n, hz(path, n), m)
        assert hz(path, 131072) <= max(hz(path, 65536), hz(path, 262144))
        assert hz(path, 1024) < hz(path, 16384)
        assert all(hz(path, n) < 100 for n in (512, 4096, 16384, 65536, 131072, 262144, 1 << 20))
    assert hz(_lib.PATH_TENSOR_FAST, 262144) == 25      # the plan bench.py's default run executes (warm-up 5 + 20 steps)
    assert hz(_lib.PATH_TENSOR, 65536) == hz(_lib.PATH_TENSOR_LO8, 65536)
    assert hz(_lib.PATH_FP64, 65536) > 1e6               # no horizon: exact
