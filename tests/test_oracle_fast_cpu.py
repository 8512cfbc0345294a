"""Pins for oracle/fast_cpu.c -- the OPTIMISED CPU baseline (radix-2 NTT, Pippenger) and the
O(n log n) helpers of the large-size verifier check -- against the reference-algorithm oracles."""
import random

import numpy as np
import pytest

from oracle import bls12_381 as O
from oracle import cref
from oracle import plonk as P


def mont(vals):
    return np.array([O.fr_to_mont(v) for v in vals], dtype=np.uint64)


def ints(arr):
    return [O.fr_from_mont([int(x) for x in row]) for row in arr]


@pytest.mark.parametrize("n", [1, 2, 8, 64, 1024, 8192])
def test_fast_ntt_equals_reference_dft(n):
    x = O.random_fr(n + 5, n)
    want = O.ntt_381(x) if n <= 64 else O.ntt_fast(x)
    for threads in (1, 3):
        assert ints(cref.ntt_fast(mont(x), threads=threads)) == want
        assert ints(cref.ntt_fast(cref.ntt_fast(mont(x), threads=threads), inverse=True, threads=threads)) == x
    if n <= 64:  # the C restatement of the reference's naive DFT agrees as well
        assert ints(cref.ntt_381(mont(x))) == want


def test_fast_ntt_rejects_non_power_of_two():
    with pytest.raises(AssertionError):
        cref.ntt_fast(mont([1, 2, 3]))


def test_poly_at_is_interpolate_then_evaluate():
    n = 256
    vals = O.random_fr(9, n)
    coeffs = O.ntt_fast(vals, inverse=True)
    for x in (0, 1, 101, O.Q - 1, O.random_fr(10, 1)[0]):
        got = O.fr_from_mont([int(v) for v in cref.poly_at(mont(vals), mont([x])[0], threads=2)])
        assert got == P.p_eval(coeffs, x)


@pytest.mark.parametrize("n,threads,window", [(1, 1, 0), (37, 1, 4), (300, 4, 0), (300, 3, 7), (2000, 8, 11)])
def test_pippenger_equals_reference_algorithm(n, threads, window):
    rng = random.Random(n)
    G = O.G1_GEN
    pts = [O.g1_mul(G, rng.randrange(1, 1 << 40)) for _ in range(min(n, 40))]
    pts = [pts[i % len(pts)] for i in range(n)]
    if n > 5:
        pts[3] = None                     # identity operand
        pts[4] = O.g1_neg(pts[5])         # cancelling pair with equal scalars below
    sc = [rng.randrange(O.Q) for _ in range(n)]
    if n > 5:
        sc[4] = sc[5]
        sc[0] = 0
        sc[1] = O.Q - 1
    arr = np.array([O.g1_scale_proj(p, 1) for p in pts], dtype=np.uint64)
    got = cref.msm_pippenger(arr, mont(sc), threads=threads, window=window)
    want = cref.bucket_msm(arr, mont(sc), 256, 4)
    assert O.g1_proj_limbs_to_affine([int(v) for v in got]) == O.g1_proj_limbs_to_affine([int(v) for v in want])
    if n <= 300:
        assert O.g1_proj_limbs_to_affine([int(v) for v in got]) == O.msm_naive(pts, sc)


def test_pippenger_rejects_unnormalised_points():
    arr = np.array([O.g1_scale_proj(O.G1_GEN, 5)], dtype=np.uint64)
    with pytest.raises(ValueError):
        cref.msm_pippenger(arr, mont([3]))


def test_verify_columns_agrees_with_verify():
    """the column verifier (closed-form commitments from the C transform) accepts / rejects exactly like verify()"""
    n = 32
    prog, wit, pub = P.synthetic_circuit(n, 20, seed=4)
    srs = O.generate_srs_points(n + 6, 101)
    backend = P.OracleBackend(srs, reference_msm=False)
    proof = P.prove(prog, wit, O.random_fr(42, 11), backend)
    sel = [mont(c) for c in prog.selectors()]
    sig = [mont(c) for c in prog.sigmas()]
    assert P.verify(prog, proof, pub, 101, backend.commit)
    assert P.verify_columns(n, sel, sig, proof, pub, 101)
    zeta = 0x1234567
    pi_vals = [(-v) % O.Q for v in pub] + [0] * (n - len(pub))
    assert P.lagrange_public_input_eval(n, pub, zeta) == P.p_eval(O.ntt_fast(pi_vals, inverse=True), zeta)
    proof.a_bar = (proof.a_bar + 1) % O.Q
    assert not P.verify_columns(n, sel, sig, proof, pub, 101)
    proof.a_bar = (proof.a_bar - 1) % O.Q
    sel[2] = sel[2].copy()
    sel[2][5] = mont([77])[0]             # a different circuit: the proof must not verify against it
    assert not P.verify_columns(n, sel, sig, proof, pub, 101)


def test_reference_algorithm_poly_mul_and_prover_backend():
    """oracle_poly_mul (coeffs_evaluate + naive inverse DFT, polynomial.rs:241-273) equals the Python restatement;
    the reference-algorithm prover backend (CRefBackend) yields the pinned n = 8 proof"""
    a, b = O.random_fr(1, 5), O.random_fr(2, 9)
    assert ints(cref.poly_mul(mont(a), mont(b))) == (O.Polynomial(a) * O.Polynomial(b)).values
    assert ints(cref.poly_mul(mont([1, 1]), mont([1, 1]))) == [1, 2, 1]          # polynomial.rs:437-451
    prog, wit, pub = P.reference_test_circuit()
    srs = O.generate_srs_points(14, 101)
    proof = P.prove(prog, wit, list(range(1, 12)), P.CRefBackend(srs))
    assert proof.sha256() == "479cc377c535fd831b5fcaf30af5c2756c535a3ddbc20589ab6759843e974967"


def test_g1_powers_small_is_the_srs():
    srs = O.generate_srs_points(9, 101)
    start = np.array(O.g1_scale_proj(srs[2], 1), dtype=np.uint64)
    got = cref.g1_powers_small(start, 101, 7)
    assert [O.g1_proj_limbs_to_affine([int(v) for v in row]) for row in got] == srs[2:9]
