"""The C-ABI library loads and exports every symbol include/visreps_b200.h declares.
No compute calls here (no GPU needed): only size queries and argument validation,
which the library performs before touching the device."""
import ctypes as C
import os
import re

import pytest

from visreps_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "visreps_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vr_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_typed(lib):
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_sizes(lib):
    assert lib.vr_abi_version() == _lib.ABI_VERSION
    assert lib.vr_mask_words(1000) == 32 and lib.vr_mask_words(1) == 1 and lib.vr_mask_words(33) == 2
    N = 1000 * 999 // 2
    assert lib.vr_dev_stride(1000) >= N and lib.vr_dev_stride(1000) % 32 == 0
    assert lib.vr_presort_bytes(1000) >= 4 * N
    assert lib.vr_presort_workspace_bytes(1000) >= 16 * N
    assert lib.vr_rdm_workspace_bytes(1000, 4096) >= 1000 * 4096 * 4
    small = lib.vr_bootstrap_spearman_workspace_bytes(1000, 1, 1, 1)
    big = lib.vr_bootstrap_spearman_workspace_bytes(1000, 8, 8, 1001)
    assert big > small > 4 * N
    assert isinstance(lib.vr_launch_count(), int)


def test_argument_validation_happens_before_any_device_work(lib):
    # null pointers / bad sizes -> VR_ERR_INVALID with a message, no CUDA call made
    rc = lib.vr_presort_triu_f32(None, 10, 10, None, None, 0, None)
    assert rc == _lib.VR_ERR_INVALID and b"null" in lib.vr_last_error_string()
    rc = lib.vr_presort_triu_f32(C.c_void_p(16), 1, 1, C.c_void_p(16), C.c_void_p(16), 1 << 20, None)
    assert rc == _lib.VR_ERR_INVALID
    rc = lib.vr_presort_triu_f32(C.c_void_p(16), 40000, 40000, C.c_void_p(16), C.c_void_p(16), 1 << 20, None)
    assert rc == _lib.VR_ERR_INVALID and b"40000" in lib.vr_last_error_string()
    rc = lib.vr_presort_triu_f32(C.c_void_p(16), 100, 100, C.c_void_p(16), C.c_void_p(16), 16, None)
    assert rc == _lib.VR_ERR_WORKSPACE
    rc = lib.vr_rdm_pearson_f32(C.c_void_p(16), 10, 0, 0, 1e-12, C.c_void_p(16), None, None, 0, None)
    assert rc == _lib.VR_ERR_INVALID
    rc = lib.vr_build_masks(C.c_void_p(16), 4, 20, 10, C.c_void_p(16), None)   # m > n
    assert rc == _lib.VR_ERR_INVALID
    ptrs = (C.c_void_p * 1)(16)
    rc = lib.vr_bootstrap_spearman(ptrs, 1, ptrs, 1, 100, C.c_void_p(16), 0, C.c_void_p(16), C.c_void_p(16), 1 << 30, None)
    assert rc == _lib.VR_ERR_INVALID
    with pytest.raises(ValueError):
        _lib.check(rc, "x")


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import visreps_b200 as vb
    with pytest.raises(_lib.VisrepsB200Error):
        vb.compute_rdm(torch.randn(4, 3))
    with pytest.raises(_lib.VisrepsB200Error):
        vb.compute_rdm_correlation(torch.zeros(4, 4), torch.zeros(4, 4), correlation="Spearman")
    # argument errors keep the reference's precedence even without a device
    with pytest.raises(ValueError):
        vb.compute_rdm(torch.randn(4, 3), correlation="cosine")
    with pytest.raises(ValueError):
        vb.compute_rdm_correlation(torch.zeros(5, 5), torch.zeros(6, 6))
    import math
    assert math.isnan(vb.compute_rdm_correlation(torch.zeros(1, 1), torch.zeros(1, 1), correlation="Spearman"))


def test_product_never_imports_the_oracle():
    """The oracle (and SciPy, the reference's CPU arithmetic) must stay out of the product path."""
    pkg = os.path.join(ROOT, "visreps_b200")
    bad = re.compile(r"^\s*(from|import)\s+(oracle|scipy)\b", re.M)
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                text = open(os.path.join(dirpath, fn)).read()
                assert not bad.search(text), f"{fn} imports the oracle or SciPy"
