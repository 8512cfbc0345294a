"""Pin the C restatement (oracle/ref_cpu.c) against the reference KATs and the Python oracle."""
import json
import os
import random

import numpy as np
import pytest

from oracle import bls12_381 as O
from oracle import cref

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))
L = lambda xs: np.array([int(x, 16) for x in xs], dtype=np.uint64)


def test_fp_kats():
    t = KATS["fp"]["test_multiplication"]
    assert list(cref.binop("oracle_fp_mul", L(t["a"]), L(t["b"]), 6)) == list(L(t["a_times_b"]))
    t = KATS["fp"]["test_squaring"]
    assert list(cref.binop("oracle_fp_mul", L(t["a"]), L(t["a"]), 6)) == list(L(t["a_squared"]))
    t = KATS["fp"]["test_addition"]
    assert list(cref.binop("oracle_fp_add", L(t["a"]), L(t["b"]), 6)) == list(L(t["a_plus_b"]))
    t = KATS["fp"]["test_subtraction"]
    assert list(cref.binop("oracle_fp_sub", L(t["a"]), L(t["b"]), 6)) == list(L(t["a_minus_b"]))


def test_field_mul_random_vs_python():
    rng = random.Random(3)
    for _ in range(200):
        a, b = rng.randrange(O.P), rng.randrange(O.P)
        got = cref.binop("oracle_fp_mul", np.array(O.int_to_limbs(a, 6), dtype=np.uint64),
                         np.array(O.int_to_limbs(b, 6), dtype=np.uint64), 6)
        assert O.limbs_to_int(got) == a * b * O.FP_RINV % O.P
        a, b = rng.randrange(O.Q), rng.randrange(O.Q)
        got = cref.binop("oracle_fr_mul", np.array(O.int_to_limbs(a, 4), dtype=np.uint64),
                         np.array(O.int_to_limbs(b, 4), dtype=np.uint64), 4)
        assert O.limbs_to_int(got) == a * b * O.FR_RINV % O.Q


def P(pts, zs=None):
    return np.array([O.g1_scale_proj(p, 1 if zs is None else z) for p, z in zip(pts, zs or [1] * len(pts))],
                    dtype=np.uint64)


def test_complete_formulas_all_cases():
    G = O.G1_GEN
    pts = [None, G, O.g1_double(G), O.g1_neg(G), O.g1_mul(G, 12345)]
    zs = [7, 1, 11, 13, 17]
    arr = P(pts, zs)
    dg = KATS["g1"]["double_generator"]
    out = np.zeros(18, dtype=np.uint64)
    cref.lib().oracle_g1_double(arr[1].ctypes.data, out.ctypes.data)
    two_g = O.g1_proj_limbs_to_affine([int(v) for v in out])
    assert O.fp_to_mont(two_g[0]) == [int(x, 16) for x in dg["x"]]
    for i, a in enumerate(pts):
        cref.lib().oracle_g1_double(arr[i].ctypes.data, out.ctypes.data)
        assert O.g1_proj_limbs_to_affine([int(v) for v in out]) == O.g1_double(a)
        for j, b in enumerate(pts):
            got = cref.binop("oracle_g1_add", arr[i], arr[j], 18)
            assert O.g1_proj_limbs_to_affine([int(v) for v in got]) == O.g1_add(a, b), (i, j)


def S(ints):
    return np.array([O.fr_to_mont(v) for v in ints], dtype=np.uint64)


@pytest.mark.parametrize("b,c", [(256, 4), (256, 8), (256, 5), (200, 4), (256, 1)])
def test_bucket_msm_vs_python(b, c):
    pts = O.generate_srs_points(9, 101)
    sc = O.random_fr(b + c, 9) + [5]  # extra scalar: zip truncation
    got = cref.bucket_msm(P(pts, list(range(2, 11))), S(sc), b, c)
    assert O.g1_proj_limbs_to_affine([int(v) for v in got]) == O.bucket_msm(pts, sc, b, c)
    got = cref.bucket_msm(P(pts), S(sc), b, c, threads=3)
    assert O.g1_proj_limbs_to_affine([int(v) for v in got]) == O.bucket_msm(pts, sc, b, c)


def test_bucket_msm_panics_and_empty():
    with pytest.raises(IndexError):
        cref.bucket_msm(P([O.G1_GEN]), S([1]), 3, 4)
    with pytest.raises(IndexError):
        cref.bucket_msm(P([O.G1_GEN]), S([1]), 300, 4)
    got = cref.bucket_msm(np.zeros((0, 18), dtype=np.uint64), np.zeros((0, 4), dtype=np.uint64))
    assert O.g1_proj_limbs_to_affine([int(v) for v in got]) is None


def test_commit_pins_setup_rs():
    srs = O.generate_srs_points(2, 10)
    got = cref.bucket_msm(P(srs), S([2, 3]))
    assert O.g1_proj_limbs_to_affine([int(v) for v in got]) == O.g1_mul(O.G1_GEN, 32)


@pytest.mark.parametrize("n", [1, 2, 8, 32])
def test_naive_dft_vs_python(n):
    x = O.random_fr(n, n)
    f = cref.ntt_381(S(x))
    assert [O.fr_from_mont(r) for r in f] == O.ntt_381(x)
    g = cref.ntt_381(S(x), inverse=True)
    assert [O.fr_from_mont(r) for r in g] == O.i_ntt_381(x)
    with pytest.raises(AssertionError):
        cref.ntt_381(S([1, 2, 3]))


def test_fr_horner():
    sc = O.random_fr(5, 50)
    got = cref.fr_horner(S(sc), S([101])[0])
    assert O.fr_from_mont(got) == sum(s * pow(101, i, O.Q) for i, s in enumerate(sc)) % O.Q


def test_g1_iota():
    pts = cref.g1_iota(20)
    for i in (0, 1, 2, 19):
        assert O.g1_proj_limbs_to_affine([int(v) for v in pts[i]]) == O.g1_mul(O.G1_GEN, i + 1)
