"""Pin the CPU oracle (oracle/bls12_381.py) against the reference's own known-answer vectors.

Fixtures: tests/golden/reference_kats.json (made by tests/golden/make_golden.py from the
reference checkout).  These are the vectors SURVEY.md section 8c lists for the MSM / NTT path.
"""
import hashlib
import json
import os

import pytest

from oracle import bls12_381 as O

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))


def L(xs):
    return [int(x, 16) for x in xs]


# ------------------------------------------------------------------ Fp (fp.rs:700-941)
def test_fp_constants():
    fp = KATS["fp"]
    assert O.limbs_to_int(L(fp["MODULUS"]["limbs"])) == O.P
    assert O.limbs_to_int(L(fp["R"]["limbs"])) == O.FP_R
    assert O.limbs_to_int(L(fp["R2"]["limbs"])) == pow(2, 768, O.P)
    assert O.limbs_to_int(L(fp["R3"]["limbs"])) == pow(2, 1152, O.P)
    inv = L(fp["INV"]["limbs"])[0]
    assert (inv * O.P + 1) % (1 << 64) == 0
    # 32-bit Montgomery constant used by the CUDA kernels
    assert (inv & 0xFFFFFFFF) == 0xFFFCFFFD


def test_fp_mul_square_kat():
    t = KATS["fp"]["test_multiplication"]
    assert O.fp_mont_mul_limbs(L(t["a"]), L(t["b"])) == L(t["a_times_b"])
    t = KATS["fp"]["test_squaring"]
    assert O.fp_mont_mul_limbs(L(t["a"]), L(t["a"])) == L(t["a_squared"])


def test_fp_add_sub_neg_inv_kat():
    t = KATS["fp"]["test_addition"]
    a, b, c = (O.limbs_to_int(L(t[k])) for k in ("a", "b", "a_plus_b"))
    assert (a + b) % O.P == c
    t = KATS["fp"]["test_subtraction"]
    a, b, c = (O.limbs_to_int(L(t[k])) for k in ("a", "b", "a_minus_b"))
    assert (a - b) % O.P == c
    t = KATS["fp"]["test_negation"]
    a, b = (O.limbs_to_int(L(t[k])) for k in ("a", "neg_a"))
    assert (-a) % O.P == b
    t = KATS["fp"]["test_inversion"]
    a, b = L(t["a"]), L(t["a_inv"])
    # Montgomery: inv(aR) = a^-1 R
    assert O.fp_to_mont(pow(O.fp_from_mont(a), -1, O.P)) == b


# ------------------------------------------------------------------ Fr (scalar.rs:80-221,795-1046)
def test_fr_constants():
    fr = KATS["fr"]
    assert O.limbs_to_int(L(fr["MODULUS"]["limbs"])) == O.Q
    assert O.limbs_to_int(L(fr["R"]["limbs"])) == O.FR_R
    assert O.limbs_to_int(L(fr["R2"]["limbs"])) == pow(2, 512, O.Q)
    assert O.limbs_to_int(L(fr["R3"]["limbs"])) == pow(2, 768, O.Q)
    assert fr["S"]["value"] == O.FR_S == 32
    assert O.fr_from_mont(L(fr["GENERATOR"]["limbs"])) == O.FR_GENERATOR
    assert O.fr_from_mont(L(fr["ROOT_OF_UNITY"]["limbs"])) == O.ROOT_OF_UNITY
    assert O.fr_from_mont(L(fr["ROOT_OF_UNITY_INV"]["limbs"])) == O.ROOT_OF_UNITY_INV
    assert O.fr_from_mont(L(fr["TWO_INV"]["limbs"])) == pow(2, -1, O.Q)
    assert O.fr_from_mont(L(fr["DELTA"]["limbs"])) == pow(O.FR_GENERATOR, 1 << 32, O.Q)
    inv = L(fr["INV"]["limbs"])[0]
    assert (inv * O.Q + 1) % (1 << 64) == 0
    assert (inv & 0xFFFFFFFF) == 0xFFFFFFFF
    # survey-time model value of ROOT_OF_UNITY
    assert O.ROOT_OF_UNITY == 0x16A2A19EDFE81F20D09B681922C813B4B63683508C2280B93829971F439F0D2B
    assert O.root_of_unity(8) == 0x345766F603FA66E78C0625CD70D77CE2B38B21C28713B7007228FD3397743F7A


def test_fr_from_bytes_wide_kat():
    exp = L(KATS["fr"]["from_bytes_wide_all_ff"]["limbs"])
    assert O.fr_to_mont(O.fr_from_bytes_wide(b"\xff" * 64)) == exp
    # scalar.rs:1011-1033
    assert O.fr_from_bytes_wide(O.FR_R.to_bytes(32, "little") + bytes(32)) == O.FR_R
    assert O.fr_from_bytes_wide((O.Q - 1).to_bytes(32, "little") + bytes(32)) == O.Q - 1


def test_fr_mul_vs_double_and_add():
    """scalar.rs:1113-1140: multiplication equals double-and-add, on LARGEST multiples"""
    largest = O.Q - 1
    cur = largest
    for _ in range(20):
        cur_m = O.fr_to_mont(cur)
        prod = O.fr_mont_mul_limbs(cur_m, cur_m)
        acc = 0
        for bit in bin(cur)[2:]:
            acc = (acc + acc) % O.Q
            if bit == "1":
                acc = (acc + cur) % O.Q
        assert O.fr_from_mont(prod) == acc
        cur = (cur + largest) % O.Q


# ------------------------------------------------------------------ G1 (g1.rs, tests/mod.rs)
def test_g1_constants():
    g = KATS["g1"]
    assert O.fp_from_mont(L(g["B"]["limbs"])) == O.CURVE_B
    assert O.fp_from_mont(L(g["generator"]["x"])) == O.G1_X
    assert O.fp_from_mont(L(g["generator"]["y"])) == O.G1_Y
    assert O.g1_is_on_curve(O.G1_GEN)
    assert O.g1_mul(O.G1_GEN, O.Q - 1) == O.g1_neg(O.G1_GEN)
    assert O.g1_add(O.g1_mul(O.G1_GEN, O.Q - 1), O.G1_GEN) is None


def test_g1_double_generator_kat():
    g = KATS["g1"]["double_generator"]
    two_g = O.g1_double(O.G1_GEN)
    assert O.fp_to_mont(two_g[0]) == L(g["x"])
    assert O.fp_to_mont(two_g[1]) == L(g["y"])
    assert O.g1_proj_limbs_to_affine(O.g1_scale_proj(two_g, 0x1234567)) == two_g


def test_g1_encoding_vectors_all_1000():
    """[i]G for i = 0..999, compressed and uncompressed, byte for byte (digest of the .dat files)"""
    comp = bytearray()
    unc = bytearray()
    e = None
    pts = []
    for _ in range(1000):
        pts.append(e)
        comp += O.g1_to_compressed(e)
        unc += O.g1_to_uncompressed(e)
        e = O.g1_add(e, O.G1_GEN)
    dc, du = KATS["g1"]["dat_compressed"], KATS["g1"]["dat_uncompressed"]
    for i, hx in dc["sample"].items():
        assert comp[int(i) * 48:(int(i) + 1) * 48].hex() == hx
    for i, hx in du["sample"].items():
        assert unc[int(i) * 96:(int(i) + 1) * 96].hex() == hx
        assert O.g1_from_uncompressed(bytes.fromhex(hx)) == pts[int(i)]
    assert hashlib.sha256(comp).hexdigest() == dc["sha256"]
    assert hashlib.sha256(unc).hexdigest() == du["sha256"]
    # scalar multiplication agrees with the addition chain
    for i in (0, 1, 2, 3, 17, 255, 999):
        assert O.g1_mul(O.G1_GEN, i) == pts[i]


# ------------------------------------------------------------------ MSM pins (src/setup.rs:45-116)
def test_generate_srs_tau2():
    srs = O.generate_srs_points(8, 2)
    for i in range(8):
        assert srs[i] == O.g1_mul(O.G1_GEN, pow(2, i, O.Q))


def test_monomial_commit_pins():
    # setup.rs:59-72: commit([2,3]) with tau = 10 == [2]G + [30]G
    srs = O.generate_srs_points(2, 10)
    assert O.bucket_msm(srs, [2, 3], 256, 4) == O.g1_mul(O.G1_GEN, 32)
    # setup.rs:74-90: commit([0,1]) with tau = 2 == [2]G
    srs = O.generate_srs_points(8, 2)
    assert O.bucket_msm(srs, [0, 1], 256, 4) == O.g1_mul(O.G1_GEN, 2)
    # setup.rs:92-116: commit((3x^2+2x+1)(x-1)) == [(tau-1)] commit(3x^2+2x+1)   (pairing identity in G1)
    p1 = O.Polynomial([1, 2, 3])
    p2 = p1 * O.Polynomial([O.Q - 1, 1])
    assert O.bucket_msm(srs, p2.values) == O.g1_mul(O.bucket_msm(srs, p1.values), 2 - 1)


def test_bucket_msm_semantics():
    pts = O.generate_srs_points(6, 101)
    sc = O.random_fr(7, 6) + [O.Q - 1]          # one more scalar than points: zip truncation
    exp = O.msm_naive(pts, sc)
    assert O.bucket_msm(pts, sc, 256, 4) == exp
    assert O.bucket_msm(pts, sc, 256, 8) == exp
    assert O.bucket_msm(pts, sc, 256, 2) == exp
    # c not dividing 256: the low 256 - k*c bits are dropped
    sh = 256 - (256 // 5) * 5
    assert O.bucket_msm(pts, sc, 256, 5) == O.msm_naive(pts, [s >> sh for s in sc])
    # empty input -> identity
    assert O.bucket_msm([], [], 256, 4) is None
    # identical points / cancelling points (tau = 1 SRS, prover.rs:684)
    same = O.generate_srs_points(4, 1)
    assert O.bucket_msm(same, [1, 2, 3, 4]) == O.g1_mul(O.G1_GEN, 10)
    assert O.bucket_msm(same, [5, O.Q - 5, 0, 0]) is None


# ------------------------------------------------------------------ NTT pins
def test_omega4_pow4_is_one():
    # utils.rs:238-242
    assert pow(O.root_of_unity(4), 4, O.Q) == 1
    assert pow(O.root_of_unity(4), 2, O.Q) != 1


def test_poly_mul_pin():
    # polynomial.rs:437-451: (1+x)^2 = 1 + 2x + x^2 through the evaluate / i_ntt path
    assert (O.Polynomial([1, 1]) * O.Polynomial([1, 1])).values == [1, 2, 1]


@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 64])
def test_fast_ntt_equals_naive_dft(n):
    x = O.random_fr(n + 5, n)
    f = O.ntt_381(x)
    assert O.ntt_fast(x) == f
    assert O.i_ntt_381(f) == x
    assert O.ntt_fast(f, inverse=True) == x
    # definition: out[i] = poly(omega^i)
    w = O.root_of_unity(n)
    p = O.Polynomial(x)
    for i in range(min(n, 4)):
        assert f[i] == p.coeffs_evaluate(pow(w, i, O.Q))
    # coset
    c = O.ntt_fast(x, coset_shift=7)
    for i in range(min(n, 4)):
        assert c[i] == p.coeffs_evaluate(7 * pow(w, i, O.Q) % O.Q)
    assert O.ntt_fast(c, inverse=True, coset_shift=7) == x


def test_survey_model_intt_value():
    """i_ntt(A)[0..3] for the n=8 test circuit wire column A = [80,3,16,0,...]"""
    a = O.i_ntt_381([80, 3, 16, 0, 0, 0, 0, 0])
    assert a[0] == 0x48748893FA026E4D200427050605270354568681DFFEF97F5FFFFFFF6000000D
    assert a[1] == 0x15C807036792F9372C243A452A6FE499677DB2E75C37790E9BE9008F1E8681F9
    assert a[2] == 0x73EDA753299D7D47FE3B2B3A9D60B636FB3C840213BD3BFEFFFF9FFF00000009
