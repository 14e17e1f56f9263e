"""CPU: the C-ABI library builds, loads and exports every symbol include/cglb_b200.h declares; without a
GPU every compute entry fails loudly (no fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from cglb_b200 import _ffi, build
    if not os.path.isfile(_ffi.LIB_PATH):
        build.build(verbose=False)
    return _ffi.load_library()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "cglb_b200.h")).read()
    return sorted(set(re.findall(r"CGLB_API\s+[\w\s\*]+?\b(cglb_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    from cglb_b200 import _ffi
    declared = _declared_symbols()
    assert len(declared) >= 31
    for must in ("cglb_kmv_sym", "cglb_kmv_bwd_sym", "cglb_kmv_sym_f32", "cglb_kmv_bwd_sym_f32", "cglb_pack_inputs_f32",
                 "cglb_kmv_sym_variant", "cglb_potrf", "cglb_precond_project", "cglb_cg_step"):
        assert must in declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cglb_b200.h but not exported"
        assert name in _ffi.SIGNATURES, f"{name} has no ctypes signature in cglb_b200/_ffi.py"
    assert sorted(_ffi.SIGNATURES) == declared


def test_abi_helpers_without_gpu(lib):
    assert lib.cglb_abi_version() == 1
    assert lib.cglb_packed_width(11) == 12 and lib.cglb_packed_width(8) == 10 and lib.cglb_packed_width(1) == 2
    assert lib.cglb_padded_rows(300) == 384 and lib.cglb_padded_rows(128) == 128 and lib.cglb_padded_rows(0) == 0
    # d > 32: wide DMMA layout (coordinates padded to a multiple of 4, row pitch = 4 or 12 mod 16)
    assert lib.cglb_packed_width(90) == 100 and lib.cglb_packed_width(33) == 44 and lib.cglb_packed_width(64) == 68
    # fp32-pair layout: d + 1 floats rounded up to 16 bytes
    assert [lib.cglb_packed_width_f32(d) for d in (1, 3, 4, 8, 11, 12, 32)] == [4, 4, 8, 12, 12, 16, 36]
    # queries with a null context fail with a status, never crash
    assert lib.cglb_kmv_sym_variant(None, 11, 1000, 1) < 0


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.cglb_create(C.byref(h), 0)
    assert rc != 0 and b"no CPU fallback" in lib.cglb_last_error()
    import cglb_b200 as cb
    from cglb_b200.engine import get_engine
    with pytest.raises(cb.CglbError):
        get_engine()
    from helpers import make_model
    from oracle import cglb_oracle as o
    x, y, z = o.synthetic_problem(50, 2, 5)
    model = make_model("matern32", x.numpy(), y.numpy(), z.numpy(), 0.1, 1.0, 1.0, device="cpu")
    with pytest.raises(cb.CglbError):
        cb.LowerBoundCG(model)((model.train_inputs[0], model.train_targets))
    with pytest.raises(cb.CglbError):
        cb.ConjugateGradient()(torch.eye(3, dtype=torch.float64), torch.ones(3, 1, dtype=torch.float64),
                               torch.zeros(3, 1, dtype=torch.float64), lambda r: (r, (r * r).sum()))


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cglb_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
