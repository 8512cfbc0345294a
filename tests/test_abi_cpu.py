"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/bpk.h declares, the host layer mirrors the reference's helpers, and -- with no GPU -- the
product path fails loudly instead of falling back to anything."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from tests._bpk import bpk, ROOT
from oracle import bls12_381 as O


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build()
    return bpk.load_library()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "bpk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bpk_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    syms = header_symbols()
    assert len(syms) >= 25
    assert tuple(syms) == bpk.EXPORTED_SYMBOLS
    for s in syms:
        assert hasattr(lib, s), s


def test_exports_in_dynamic_symbol_table(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", bpk.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (bpk_[a-z0-9_]+)", out))
    assert set(header_symbols()) <= exported


def test_library_is_sm100a_only(lib):
    out = subprocess.run(["cuobjdump", "-lelf", bpk.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_abi_version_and_strerror(lib):
    assert lib.bpk_abi_version() == 1
    assert b"power of two" in lib.bpk_strerror(-4)
    assert b"no CPU fallback" in lib.bpk_strerror(-1)


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    assert lib.bpk_init(ctypes.byref(h), 0) == -1  # BPK_ERR_NO_DEVICE
    with pytest.raises(bpk.BpkPanic):
        bpk.Context(0)
    with pytest.raises(bpk.BpkPanic):
        bpk.ntt_381(bpk.scalars_from_ints([1, 2, 3, 4]))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "baby-plonk-rust_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert not re.search(r"import_module\(.oracle", src), f
                assert "liboracle" not in src, f


def test_host_value_codecs_match_oracle():
    vals = [0, 1, 2, O.Q - 1, O.Q + 5] + O.random_fr(3, 5)
    arr = bpk.scalars_from_ints(vals)
    for row, v in zip(arr, vals):
        assert [int(x) for x in row] == O.fr_to_mont(v)
    assert bpk.scalars_to_ints(arr) == [v % O.Q for v in vals]
    pts = [None, O.G1_GEN, O.g1_mul(O.G1_GEN, 77)]
    arr = bpk.points_from_affine(pts)
    for row, p in zip(arr, pts):
        assert [int(x) for x in row] == O.g1_affine_to_proj_limbs(p)
        assert bpk.point_to_affine(row) == p
        assert bpk.point_to_compressed(row) == O.g1_to_compressed(p)
    arr = bpk.points_from_affine(pts, z_scale=[5, 6, 7])
    for row, p in zip(arr, pts):
        assert bpk.point_to_affine(row) == p


def test_utils_helpers_match_oracle():
    for n in (1, 2, 4, 8, 1024):
        assert bpk.root_of_unity(n) == O.root_of_unity(n)
    assert bpk.roots_of_unity(8) == O.roots_of_unity(8)
    for a, b in ((0, 0), (1, 1), (3, 4), (7, 8), (100, 28)):
        assert bpk.find_next_power_of_two(a, b) == O.find_next_power_of_two(a, b)
    assert bpk.is_power_of_two(8) and not bpk.is_power_of_two(0) and not bpk.is_power_of_two(12)


def test_reference_panics_are_mirrored_on_host():
    with pytest.raises(bpk.BpkPanic):
        bpk.Polynomial.from_ints([1, 2], bpk.Basis.Lagrange).ntt()
    with pytest.raises(bpk.BpkPanic):
        bpk.Polynomial.from_ints([1, 2], bpk.Basis.Monomial).i_ntt()
    a = bpk.Polynomial.from_ints([1, 2], bpk.Basis.Lagrange)
    with pytest.raises(bpk.BpkPanic):
        a * a  # todo!() in the reference
