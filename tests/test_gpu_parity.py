"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI (ctypes -> libbpk.so),
against the CPU oracle on the same seeded inputs; full-size cases through size-independent
properties (round trips, closed-form MSM answers with known discrete logs, spot evaluations).

Bit-exact bar: every comparison is on integers (canonical field values / affine coordinates /
raw Montgomery limbs); there is no tolerance anywhere.
"""
import random

import numpy as np
import pytest

from oracle import bls12_381 as O
from tests._bpk import bpk

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = bpk.Context(0)
    yield c
    c.close()


def S(ints):
    return bpk.scalars_from_ints(ints)


def I(arr):
    return bpk.scalars_to_ints(arr)


# =============================================================================================
# NTT   (src/utils.rs:63-129, src/polynomial.rs:47-55, 189-273)
# =============================================================================================
@pytest.mark.parametrize("logn", list(range(0, 7)))
def test_ntt_equals_reference_naive_dft(ctx, logn):
    n = 1 << logn
    x = O.random_fr(100 + logn, n)
    f = bpk.ntt_381(S(x), ctx)
    assert I(f) == O.ntt_381(x)
    assert I(bpk.i_ntt_381(f, ctx)) == x
    assert I(bpk.i_ntt_381(S(x), ctx)) == O.i_ntt_381(x)


@pytest.mark.parametrize("logn", [7, 8, 9, 10, 11, 12, 13, 14, 16])
def test_ntt_medium_vs_oracle_fft(ctx, logn):
    n = 1 << logn
    x = O.random_fr(200 + logn, n)
    f = bpk.ntt_381(S(x), ctx)
    assert I(f) == O.ntt_fast(x)
    g = bpk.i_ntt_381(S(x), ctx)
    assert I(g) == O.ntt_fast(x, inverse=True)


def test_ntt_raw_limbs_are_canonical_montgomery(ctx):
    x = O.random_fr(7, 256)
    f = bpk.ntt_381(S(x), ctx)
    exp = np.array([O.fr_to_mont(v) for v in O.ntt_fast(x)], dtype=np.uint64)
    assert np.array_equal(f, exp)


def test_ntt_edge_inputs(ctx):
    n = 512
    for x in ([0] * n, [1] + [0] * (n - 1), [O.Q - 1] * n, [1] * n):
        assert I(bpk.ntt_381(S(x), ctx)) == O.ntt_fast(x)
        assert I(bpk.i_ntt_381(S(x), ctx)) == O.ntt_fast(x, inverse=True)


def test_ntt_batch(ctx):
    n, batch = 2048, 5
    xs = [O.random_fr(300 + b, n) for b in range(batch)]
    arr = np.stack([S(x) for x in xs])
    f = bpk.ntt_381(arr, ctx)
    g = bpk.i_ntt_381(arr, ctx)
    for b in range(batch):
        assert I(f[b]) == O.ntt_fast(xs[b])
        assert I(g[b]) == O.ntt_fast(xs[b], inverse=True)


@pytest.mark.parametrize("tile", [4, 7, 9, 11, 12])
def test_ntt_tile_shapes(ctx, tile):
    """every pass shape (R, C) the planner can emit, including 3-pass plans at small n"""
    ctx.set_option("ntt.tile_log2", tile)
    try:
        for logn in (3, 6, 10, 13):
            n = 1 << logn
            x = O.random_fr(400 + logn + tile, n)
            assert I(bpk.ntt_381(S(x), ctx)) == O.ntt_fast(x), (tile, logn)
            assert I(bpk.i_ntt_381(S(x), ctx)) == O.ntt_fast(x, inverse=True), (tile, logn)
            assert I(bpk.coset_ntt(S(x), 7, ctx)) == O.ntt_fast(x, coset_shift=7), (tile, logn)
    finally:
        ctx.set_option("ntt.tile_log2", 10)


@pytest.mark.parametrize("kernel", [1, 2, 3])
@pytest.mark.parametrize("tile,maxr", [(5, 0), (8, 3), (9, 4), (10, 5), (10, 0), (11, 7), (11, 11)])
def test_ntt_both_pass_kernels_every_step_shape(ctx, kernel, tile, maxr):
    """ntt.kernel 1 = one radix-2 stage per barrier, 2 = register-blocked radix-8 steps (falls back to 1 for
    R < 8); radices 3..11 exercise full steps, 1- and 2-stage remainder steps, first-pass and later-pass stores"""
    ctx.set_option("ntt.kernel", kernel)
    ctx.set_option("ntt.tile_log2", tile)
    ctx.set_option("ntt.max_radix_log2", maxr)
    try:
        for logn, batch in ((3, 1), (4, 3), (7, 2), (11, 1), (13, 2)):
            n = 1 << logn
            xs = [O.random_fr(900 + logn + tile + b, n) for b in range(batch)]
            arr = np.stack([S(x) for x in xs])
            f, g = bpk.ntt_381(arr, ctx), bpk.i_ntt_381(arr, ctx)
            cs, ci = bpk.coset_ntt(arr, 7, ctx), bpk.coset_intt(arr, 7, ctx)
            for b in range(batch):
                assert I(f[b]) == O.ntt_fast(xs[b]), (kernel, tile, maxr, logn)
                assert I(g[b]) == O.ntt_fast(xs[b], inverse=True), (kernel, tile, maxr, logn)
                assert I(cs[b]) == O.ntt_fast(xs[b], coset_shift=7), (kernel, tile, maxr, logn)
                assert I(ci[b]) == O.ntt_fast(xs[b], inverse=True, coset_shift=7), (kernel, tile, maxr, logn)
    finally:
        ctx.set_option("ntt.kernel", 0)
        ctx.set_option("ntt.tile_log2", 10)
        ctx.set_option("ntt.max_radix_log2", 0)


def test_ntt_large_sizes_linear_and_round_trip(ctx):
    """sizes the oracle cannot transform: the two pass kernels must agree bit for bit, the transform must be
    linear and the inverse must undo it (2^20 x 3 and 2^22, the bench shapes)"""
    import torch

    rng = np.random.default_rng(77)
    for logn, batch in ((20, 3), (22, 1)):
        n = 1 << logn
        a = rng.integers(0, 1 << 64, size=(batch * n, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 62) - 1)
        x = torch.from_numpy(a.view(np.int64)).cuda()
        outs = []
        for kernel in (1, 2):
            ctx.set_option("ntt.kernel", kernel)
            y = torch.empty_like(x)
            ctx.check(ctx.lib.bpk_ntt_fr_dev(ctx.handle, x.data_ptr(), y.data_ptr(), n, batch, 0, None), "ntt")
            z = torch.empty_like(x)
            ctx.check(ctx.lib.bpk_ntt_fr_dev(ctx.handle, y.data_ptr(), z.data_ptr(), n, batch, 1, None), "intt")
            assert torch.equal(z, x), (logn, kernel)
            outs.append(y)
        ctx.set_option("ntt.kernel", 0)
        assert torch.equal(outs[0], outs[1]), logn
        # X(0) = sum of the inputs (row 0 of the DFT matrix is all ones)
        first = bpk.scalars_to_ints(outs[0][0:1].cpu().numpy().view(np.uint64))[0]
        col = a[:n].astype(object)
        total = sum(int(col[:, k].sum()) << (64 * k) for k in range(4))
        assert first == total * pow(1 << 256, -1, O.Q) % O.Q


def test_ntt_direct_and_two_level_twiddles_agree(ctx):
    """inter-pass twiddles come either from a per-size direct table or from the two-level table"""
    for logn in (11, 14, 17):
        x = S(O.random_fr(700 + logn, 1 << logn))
        ctx.set_option("ntt.direct_max_log2", 0)
        try:
            a = bpk.ntt_381(x, ctx)
            ai = bpk.i_ntt_381(x, ctx)
        finally:
            ctx.set_option("ntt.direct_max_log2", 25)
        assert np.array_equal(a, bpk.ntt_381(x, ctx))
        assert np.array_equal(ai, bpk.i_ntt_381(x, ctx))
        if logn <= 14:
            assert I(a) == O.ntt_fast(I(x))


@pytest.mark.parametrize("logn", [0, 1, 5, 10, 12])
@pytest.mark.parametrize("shift", [7, 2, 3, 0x123456789ABCDEF0123])
def test_coset_ntt(ctx, logn, shift):
    n = 1 << logn
    x = O.random_fr(500 + logn, n)
    c = bpk.coset_ntt(S(x), shift, ctx)
    assert I(c) == O.ntt_fast(x, coset_shift=shift)
    assert I(bpk.coset_intt(c, shift, ctx)) == x
    y = O.random_fr(600 + logn, n)
    assert I(bpk.coset_intt(S(y), shift, ctx)) == O.ntt_fast(y, inverse=True, coset_shift=shift)


def test_coset_inverse_table_is_keyed_by_n(ctx):
    for n in (64, 4096, 64, 256):
        y = O.random_fr(n, n)
        assert I(bpk.coset_intt(S(y), 7, ctx)) == O.ntt_fast(y, inverse=True, coset_shift=7)


def test_ntt_not_power_of_two_panics(ctx):
    # assert!(is_power_of_two(n))  utils.rs:65,108
    for n in (3, 6, 12, 1000):
        with pytest.raises(bpk.BpkPanic):
            bpk.ntt_381(S([1] * n), ctx)
        with pytest.raises(bpk.BpkPanic):
            bpk.i_ntt_381(S([1] * n), ctx)
    out = np.empty((3, 4), dtype=np.uint64)
    a = S([1, 2, 3])
    assert ctx.lib.bpk_ntt_fr(ctx.handle, a.ctypes.data, out.ctypes.data, 3, 1) == -4
    assert ctx.lib.bpk_ntt_fr(ctx.handle, a.ctypes.data, out.ctypes.data, 0, 1) == -4


def test_polynomial_wrappers_and_mul_pin(ctx):
    # polynomial.rs:437-451: (1 + x)^2 = 1 + 2x + x^2 through the evaluate / i_ntt path
    p = bpk.Polynomial.from_ints([1, 1], ctx=ctx)
    assert (p * p).to_ints() == [1, 2, 1]
    x = O.random_fr(9, 64)
    pl = bpk.Polynomial.from_ints(x, ctx=ctx).ntt()
    assert pl.basis == bpk.Basis.Lagrange and pl.to_ints() == O.ntt_381(x)
    assert pl.i_ntt().to_ints() == x


@pytest.mark.parametrize("la,lb", [(1, 1), (1, 5), (2, 2), (3, 9), (9, 3), (17, 16), (100, 29), (1000, 1025), (4097, 5)])
def test_poly_mul_vs_oracle(ctx, la, lb):
    a = O.random_fr(la, la)
    b = O.random_fr(1000 + lb, lb)
    if la > 2:
        a[-1] = 0  # trailing zero coefficient is kept (polynomial.rs:272)
    got = bpk.Polynomial.from_ints(a, ctx=ctx) * bpk.Polynomial.from_ints(b, ctx=ctx)
    exp = O.Polynomial(a) * O.Polynomial(b)
    assert got.to_ints() == exp.values
    assert len(got.to_ints()) == la + lb - 1


def _mont_sum(rows):
    """exact sum of Montgomery residues (uint64[m, 4]) as a Python int mod q, vectorised"""
    total = 0
    for limb in range(4):
        col = rows[:, limb]
        lo = int(np.sum(col & np.uint64(0xFFFFFFFF), dtype=np.uint64))
        hi = int(np.sum(col >> np.uint64(32), dtype=np.uint64))
        total += (lo + (hi << 32)) << (64 * limb)
    return total % O.Q


def _mont_val(row):
    return sum(int(v) << (64 * i) for i, v in enumerate(row))


@pytest.mark.parametrize("logn", [18, 20, 22, 24])
def test_ntt_large_properties(ctx, logn):
    """sizes the oracle cannot transform in seconds: inverse(forward(x)) == x bit for bit, and the
    outputs at k = 0, n/2, n/4, 3n/4 against exact closed forms (sums of coefficient classes mod 4;
    the transform is linear, so these hold on the Montgomery residues themselves)"""
    n = 1 << logn
    rng = np.random.default_rng(logn)
    raw = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    raw[:, 3] &= np.uint64((1 << 62) - 1)  # any residue < 2^254 < q is a valid Montgomery value
    f = bpk.ntt_381(raw, ctx)
    assert np.array_equal(bpk.i_ntt_381(f, ctx), raw)
    s = [_mont_sum(raw[r::4]) for r in range(4)]
    i4 = O.root_of_unity(4)  # w^(n/4)
    assert _mont_val(f[0]) == (s[0] + s[1] + s[2] + s[3]) % O.Q
    assert _mont_val(f[n // 2]) == (s[0] - s[1] + s[2] - s[3]) % O.Q
    assert _mont_val(f[n // 4]) == (s[0] + i4 * s[1] - s[2] - i4 * s[3]) % O.Q
    assert _mont_val(f[3 * n // 4]) == (s[0] - i4 * s[1] - s[2] + i4 * s[3]) % O.Q
    if logn <= 18:  # one generic output by Horner
        coeffs = I(raw)
        k = 12345
        pt = pow(O.root_of_unity(n), k, O.Q)
        acc = 0
        for cf in reversed(coeffs):
            acc = (acc * pt + cf) % O.Q
        assert I(f[k:k + 1])[0] == acc


# =============================================================================================
# SRS   (src/setup.rs:12-31)
# =============================================================================================
@pytest.mark.parametrize("tau,powers", [(101, 14), (2, 8), (1, 8), (10, 2), (O.Q - 1, 5), (0, 3)])
def test_generate_srs(ctx, tau, powers):
    s = bpk.Setup.generate_srs(powers, tau, ctx)
    got = [bpk.point_to_affine(p) for p in s.powers_of_x()]
    assert got == O.generate_srs_points(powers, tau)
    raw = s.powers_of_x()
    for row, pt in zip(raw, got):
        assert [int(v) for v in row] == O.g1_affine_to_proj_limbs(pt)
    s.free()


def test_generate_srs_spot_checks_large(ctx):
    n = 1 << 14
    s = bpk.Setup.generate_srs(n, 101, ctx)
    for i in (0, 1, 255, 256, 4095, n - 1):
        got = bpk.point_to_affine(s.powers_of_x(i, 1)[0])
        assert got == O.g1_mul(O.G1_GEN, pow(101, i, O.Q))
    s.free()


def test_srs_load_normalises_projective_points(ctx):
    rng = random.Random(5)
    pts = [O.g1_mul(O.G1_GEN, k) for k in (1, 2, 3, 77, O.Q - 1)] + [None, O.G1_GEN]
    zs = [rng.randrange(1, O.P) for _ in pts]
    zs[0] = 1
    arr = bpk.points_from_affine(pts, zs)
    s = bpk.Setup.from_points(arr, ctx)
    back = s.powers_of_x()
    for row, pt in zip(back, pts):
        assert [int(v) for v in row] == O.g1_affine_to_proj_limbs(pt)
    s.free()


# =============================================================================================
# MSM   (src/msm.rs:76-139, src/setup.rs:32-37)
# =============================================================================================
def rand_points(seed, n):
    rng = random.Random(seed)
    return [O.g1_mul(O.G1_GEN, rng.randrange(1, O.Q)) for _ in range(n)]


@pytest.mark.parametrize("n", [0, 1, 2, 3, 7, 33, 150])
def test_bucket_msm_small_vs_oracle(ctx, n):
    pts = rand_points(n, n)
    sc = O.random_fr(n + 1, n)
    got = bpk.BucketMSM.bucket_msm(bpk.points_from_affine(pts), S(sc), 256, 4, ctx)
    assert bpk.point_to_affine(got) == O.msm_naive(pts, sc)
    if n <= 7:
        assert bpk.point_to_affine(got) == O.bucket_msm(pts, sc, 256, 4)
    # the returned representative is normalised: raw limbs equal the oracle's
    assert [int(v) for v in got] == O.g1_affine_to_proj_limbs(O.msm_naive(pts, sc))


def test_bucket_msm_reference_semantics(ctx):
    pts = O.generate_srs_points(6, 101)
    sc = O.random_fr(7, 6) + [O.Q - 1]  # one more scalar than points: zip truncation (msm.rs:29)
    P = bpk.points_from_affine(pts, [3, 5, 7, 11, 13, 17])
    for c in (1, 2, 4, 8, 16, 32):
        got = bpk.BucketMSM.bucket_msm(P, S(sc), 256, c, ctx)
        assert bpk.point_to_affine(got) == O.msm_naive(pts, sc)
    # more points than scalars
    got = bpk.BucketMSM.bucket_msm(P, S(sc[:4]), 256, 4, ctx)
    assert bpk.point_to_affine(got) == O.msm_naive(pts[:4], sc[:4])
    # c does not divide 256: low bits dropped (oracle restates msm.rs:119-139)
    for b, c in ((256, 5), (256, 7), (256, 3), (200, 4), (13, 6), (256, 10)):
        got = bpk.BucketMSM.bucket_msm(P, S(sc), b, c, ctx)
        assert bpk.point_to_affine(got) == O.bucket_msm(pts, sc, b, c), (b, c)
    # parameters the reference panics on
    for b, c in ((3, 4), (300, 4), (256, 0), (512, 2)):
        with pytest.raises(bpk.BpkPanic):
            bpk.BucketMSM.bucket_msm(P, S(sc), b, c, ctx)


def test_commit_pins_from_setup_rs(ctx):
    # setup.rs:59-72: commit([2,3]), tau = 10 == [32]G
    s = bpk.Setup.generate_srs(2, 10, ctx)
    assert bpk.point_to_affine(s.commit(bpk.Polynomial.from_ints([2, 3]))) == O.g1_mul(O.G1_GEN, 32)
    # setup.rs:74-90: commit([0,1]), tau = 2 == [2]G
    s2 = bpk.Setup.generate_srs(8, 2, ctx)
    assert bpk.point_to_affine(s2.commit(bpk.Polynomial.from_ints([0, 1]))) == O.g1_mul(O.G1_GEN, 2)
    # setup.rs:92-116: commit(p1 * (x - 1)) == [tau - 1] commit(p1)
    p1 = bpk.Polynomial.from_ints([1, 2, 3], ctx=ctx)
    p2 = p1 * bpk.Polynomial.from_ints([O.Q - 1, 1], ctx=ctx)
    assert bpk.point_to_affine(s2.commit(p2)) == O.g1_mul(bpk.point_to_affine(s2.commit(p1)), 2 - 1)
    with pytest.raises(bpk.BpkPanic):
        s2.commit(bpk.Polynomial.from_ints([1, 2], bpk.Basis.Lagrange))


def test_msm_degenerate_inputs(ctx):
    G = O.G1_GEN
    same = [G] * 40                      # tau = 1 SRS: every point equal (prover.rs:684)
    sc = O.random_fr(11, 40)
    got = bpk.BucketMSM.bucket_msm(bpk.points_from_affine(same), S(sc), 256, 4, ctx)
    assert bpk.point_to_affine(got) == O.g1_mul(G, sum(sc) % O.Q)
    # equal scalars on equal points: P + P inside a bucket (doubling path)
    got = bpk.BucketMSM.bucket_msm(bpk.points_from_affine(same), S([5] * 40), 256, 4, ctx)
    assert bpk.point_to_affine(got) == O.g1_mul(G, 200)
    # P + (-P): cancelling pairs, total identity
    pts = [G, O.g1_neg(G), O.g1_mul(G, 9), O.g1_neg(O.g1_mul(G, 9))]
    got = bpk.BucketMSM.bucket_msm(bpk.points_from_affine(pts), S([7, 7, 1234567, 1234567]), 256, 4, ctx)
    assert bpk.point_to_affine(got) is None
    assert [int(v) for v in got] == O.g1_affine_to_proj_limbs(None)
    # s and q - s on the same point
    got = bpk.BucketMSM.bucket_msm(bpk.points_from_affine([G, G]), S([5, O.Q - 5]), 256, 4, ctx)
    assert bpk.point_to_affine(got) is None
    # all-zero scalars, identity points, scalar q - 1
    got = bpk.BucketMSM.bucket_msm(bpk.points_from_affine(rand_points(3, 10)), S([0] * 10), 256, 4, ctx)
    assert bpk.point_to_affine(got) is None
    pts = [None, G, None, O.g1_mul(G, 3)]
    sc = [5, O.Q - 1, 7, O.Q - 1]
    got = bpk.BucketMSM.bucket_msm(bpk.points_from_affine(pts), S(sc), 256, 4, ctx)
    assert bpk.point_to_affine(got) == O.msm_naive(pts, sc)


def horner_expected(scalars, tau):
    acc = 0
    for s in reversed(scalars):
        acc = (acc * tau + s) % O.Q
    return O.g1_mul(O.G1_GEN, acc)


@pytest.mark.parametrize("window", [0, 2, 5, 8, 11, 13, 16])
@pytest.mark.parametrize("chunk", [0, 1, 7, 64])
def test_msm_window_and_chunk_shapes(ctx, window, chunk):
    """closed form: P_i = [tau^i]G  =>  sum s_i P_i = [sum s_i tau^i] G  (valid for any N)"""
    n = 3000
    setup = bpk.Setup.generate_srs(n, 101, ctx)
    sc = O.random_fr(window * 100 + chunk, n)
    ctx.set_option("msm.window", window)
    ctx.set_option("msm.chunk", chunk)
    try:
        got = setup.commit_scalars(S(sc))
    finally:
        ctx.set_option("msm.window", 0)
        ctx.set_option("msm.chunk", 0)
        setup.free()
    assert bpk.point_to_affine(got) == horner_expected(sc, 101)


@pytest.mark.parametrize("c", [0, 2, 5, 8, 13, 16, 20])
def test_msm_precomputed_srs_levels(ctx, c):
    """bpk_srs_precompute stores [2^(c w)]P_i per window; every result must stay bit-identical"""
    import torch
    n = 2500
    plain = bpk.Setup.generate_srs(n, 101, ctx)
    pre = bpk.Setup.generate_srs(n, 101, ctx).precompute(c)
    rng = random.Random(c)
    for sc in (O.random_fr(900 + c, n), [rng.choice([0, 1, 3, O.Q - 1, 80]) for _ in range(n)], [O.Q - 1] * 7):
        a = plain.commit_scalars(S(sc))
        b = pre.commit_scalars(S(sc))
        assert np.array_equal(a, b)
        assert bpk.point_to_affine(b) == horner_expected(sc, 101)
    # the SRS itself is unchanged by the precomputation
    assert np.array_equal(plain.powers_of_x(0, 50), pre.powers_of_x(0, 50))
    # the (b, c) quirk and zip truncation go through the same path
    sc = O.random_fr(5, n + 3)
    assert np.array_equal(plain.commit_scalars(S(sc), 256, 5), pre.commit_scalars(S(sc), 256, 5))
    # device-resident shard with an offset into every level
    sc = O.random_fr(6, n)
    d_sc = torch.from_numpy(S(sc).view(np.int64)).cuda()
    outs = []
    for setup in (plain, pre):
        d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
        ctx.check(ctx.lib.bpk_msm_g1_dev(ctx.handle, setup.handle, 700, d_sc.data_ptr() + 32 * 700, 1500, 1,
                                         d_out.data_ptr()))
        torch.cuda.synchronize()
        outs.append(d_out.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    plain.free()
    pre.free()


def test_msm_precomputed_degenerate_points(ctx):
    G = O.G1_GEN
    pts = [G] * 20 + [None, O.g1_neg(G), G, None]
    sc = O.random_fr(3, len(pts))
    s = bpk.Setup.from_points(bpk.points_from_affine(pts), ctx).precompute(6)
    assert bpk.point_to_affine(s.commit_scalars(S(sc))) == O.msm_naive(pts, sc)
    s.free()


class options:
    """set library tunables for a block, restore the defaults afterwards"""
    DEFAULTS = {"msm.affine_levels": -1, "msm.batch": 256, "msm.min_pairs": 1 << 18, "msm.window": 0, "msm.chunk": 0,
                "msm.tree_top": 1, "msm.level_mib": 48 << 10, "msm.scatter_l2_mib": 400, "msm.cta_shape": 0}

    def __init__(self, ctx, **kw):
        self.ctx, self.kw = ctx, {k.replace("_", ".", 1): v for k, v in kw.items()}

    def __enter__(self):
        for k, v in self.kw.items():
            self.ctx.set_option(k, v)

    def __exit__(self, *exc):
        for k in self.kw:
            self.ctx.set_option(k, self.DEFAULTS[k])


@pytest.mark.parametrize("levels", [0, 1, 2, 3, 5, 9, 14, 29])
@pytest.mark.parametrize("batch", [1, 3, 256])
def test_msm_affine_tree_levels_and_batches(ctx, levels, batch):
    """the batched-affine pairwise tree with every depth (0 = XYZZ chunks only; 1..: buckets finish inside the tree or
    leave a tail of every length; more levels than any bucket needs), batches of 1 / 3 / many additions per inversion,
    own windows and precomputed levels (few buckets, long runs)"""
    n = 3000
    setup = bpk.Setup.generate_srs(n, 101, ctx)
    sc = O.random_fr(1000 + levels * 10 + batch, n)
    want = horner_expected(sc, 101)
    with options(ctx, msm_affine_levels=levels, msm_batch=batch):
        for window in (0, 5, 13):
            with options(ctx, msm_window=window):
                assert bpk.point_to_affine(setup.commit_scalars(S(sc))) == want, window
        stats = ctx.msm_last_stats()
        assert stats["tree_levels"] == levels and stats["entries"] > 0
        assert stats["affine_adds"] + stats["xyzz_adds_bound"] + stats["nonempty_buckets"] == stats["entries"]
        if levels == 0:
            assert stats["affine_adds"] == 0
        setup.precompute(8)
        assert bpk.point_to_affine(setup.commit_scalars(S(sc))) == want
        assert bpk.point_to_affine(setup.commit_scalars(S(sc[:1]))) == O.g1_mul(O.G1_GEN, sc[0])
        if levels >= 12:   # 128 buckets x ~750 entries: the tree finishes every bucket, nothing is left to the tail
            setup.commit_scalars(S(sc))
            assert ctx.msm_last_stats()["xyzz_adds_bound"] == 0
    setup.free()


@pytest.mark.parametrize("levels", [1, 4, 20])
def test_msm_affine_tree_degenerate_points(ctx, levels):
    """every degenerate addition inside the batched-affine levels: all points equal (tau = 1: each pair is a doubling at
    every level, src/prover.rs:684), tau = 0 (identity operands), P + (-P) pairs, opposite scalars on equal points"""
    n = 2048
    with options(ctx, msm_affine_levels=levels, msm_batch=5):
        for tau in (1, 0, 2):
            setup = bpk.Setup.generate_srs(n, tau, ctx)
            for sc in (O.random_fr(levels + tau, n), [7] * n, [(1 if i % 2 else O.Q - 1) for i in range(n)]):
                assert bpk.point_to_affine(setup.commit_scalars(S(sc))) == horner_expected(sc, tau), (tau, sc[:2])
            setup.precompute(6)
            sc = O.random_fr(77, n)
            assert bpk.point_to_affine(setup.commit_scalars(S(sc))) == horner_expected(sc, tau)
            setup.free()
        G = O.G1_GEN
        pts = ([G] * 20 + [None, O.g1_neg(G), G, None]) * 30
        sc = O.random_fr(3, len(pts))
        s = bpk.Setup.from_points(bpk.points_from_affine(pts), ctx)
        assert bpk.point_to_affine(s.commit_scalars(S(sc))) == O.msm_naive(pts, sc)
        sc = [5] * len(pts)    # one bucket holds 630 copies of G, 30 of -G and 60 identities
        assert bpk.point_to_affine(s.commit_scalars(S(sc))) == O.msm_naive(pts, sc)
        s.precompute(6)
        assert bpk.point_to_affine(s.commit_scalars(S(sc))) == O.msm_naive(pts, sc)
        s.free()
        # x = 0 without being the identity: (0, 2) and (0, -2) lie on y^2 = x^3 + 4; the kernels test x == 0 first and must
        # then tell them from the identity encoding (0, 0)
        Z2, Z2n = (0, 2), (0, O.P - 2)
        pts = ([G, Z2, O.g1_double(G), Z2n, Z2, None, Z2, G] * 40)[:300]
        s = bpk.Setup.from_points(bpk.points_from_affine(pts), ctx)
        for sc in (O.random_fr(9, len(pts)), [3] * len(pts)):
            assert bpk.point_to_affine(s.commit_scalars(S(sc))) == O.msm_naive(pts, sc)
        s.free()


@pytest.mark.parametrize("levels", [0, 2, 12])
def test_msm_phased_scatter(ctx, levels):
    """the scatter in phases over the bucket range that large inputs take (forced here: 1 MiB slices -> ~10 phases), with
    few buckets and many, uniform and skewed scalars (the skew flag switches to aggregated cursor updates)"""
    n = 40000
    setup = bpk.Setup.generate_srs(n, 101, ctx)
    rng = random.Random(levels)
    cases = [O.random_fr(50 + levels, n), [0xABCDEF0123456789] * n, [rng.choice([1, 2, O.Q - 1]) for _ in range(n)]]
    with options(ctx, msm_scatter_l2_mib=1, msm_affine_levels=levels):
        for sc in cases:
            want = horner_expected(sc, 101)
            for window in (0, 14):          # 29 x 2^8 buckets / 19 x 2^13 buckets (many partitions)
                with options(ctx, msm_window=window):
                    assert bpk.point_to_affine(setup.commit_scalars(S(sc))) == want
        setup.precompute(8)                 # 128 buckets
        for sc in cases:
            assert bpk.point_to_affine(setup.commit_scalars(S(sc))) == horner_expected(sc, 101)
    setup.free()


@pytest.mark.parametrize("shape", [1, 2])
def test_msm_level_kernel_cta_shapes(ctx, shape):
    """the level kernel as four CTAs of 4 warps per SM and as one CTA of 16 (by default chosen per level by its size)"""
    n = 6000
    setup = bpk.Setup.generate_srs(n, 101, ctx).precompute(10)
    sc = O.random_fr(40 + shape, n)
    want = horner_expected(sc, 101)
    with options(ctx, msm_cta_shape=shape, msm_affine_levels=5, msm_batch=7):
        assert bpk.point_to_affine(setup.commit_scalars(S(sc))) == want
    setup.free()


def test_msm_tree_top_and_per_level_reduction_agree(ctx):
    n = 5000
    setup = bpk.Setup.generate_srs(n, 101, ctx).precompute(13)
    sc = O.random_fr(5, n)
    a = setup.commit_scalars(S(sc))
    with options(ctx, msm_tree_top=0):
        b = setup.commit_scalars(S(sc))
    setup.free()
    assert np.array_equal(a, b) and bpk.point_to_affine(a) == horner_expected(sc, 101)


@pytest.mark.parametrize("window", [2, 3, 5, 11, 12, 13, 16])
def test_msm_bit_plane_reduction_all_window_sizes(ctx, window):
    """default reduction: every plane count, with one chunk per plane (window <= 11), exactly one (12) and
    several (13, 16), own windows and precomputed levels"""
    n = 3000
    setup = bpk.Setup.generate_srs(n, 101, ctx)
    sc = O.random_fr(100 + window, n)
    want = horner_expected(sc, 101)
    ctx.set_option("msm.window", window)
    try:
        assert bpk.point_to_affine(setup.commit_scalars(S(sc))) == want
    finally:
        ctx.set_option("msm.window", 0)
    setup.precompute(window)
    assert bpk.point_to_affine(setup.commit_scalars(S(sc))) == want
    setup.free()


@pytest.mark.parametrize("dist", ["witness", "all_equal", "tiny", "qminus1", "two_values"])
def test_msm_skewed_distributions(ctx, dist):
    """heavy buckets: runs that span many chunks exercise the partial-merge path"""
    n = 1 << 13
    rng = random.Random(42)
    if dist == "witness":
        sc = [rng.randrange(1 << 16) if rng.random() < 0.9 else (0 if rng.random() < 0.5 else rng.randrange(O.Q))
              for _ in range(n)]
    elif dist == "all_equal":
        sc = [0xDEADBEEFCAFEBABE1234567] * n
    elif dist == "tiny":
        sc = [rng.choice([0, 1, 3, 4, 16, 80]) for _ in range(n)]
    elif dist == "qminus1":
        sc = [O.Q - 1] * n
    else:
        sc = [rng.choice([12345, O.Q - 12345]) for _ in range(n)]
    setup = bpk.Setup.generate_srs(n, 101, ctx)
    got = setup.commit_scalars(S(sc))
    want = horner_expected(sc, 101)
    assert bpk.point_to_affine(got) == want
    for levels in (3, 16):     # heavy buckets inside the batched-affine tree, with and without a tail
        with options(ctx, msm_affine_levels=levels):
            assert bpk.point_to_affine(setup.commit_scalars(S(sc))) == want
    setup.free()


@pytest.mark.parametrize("pre", [None, 0, 21])
def test_msm_heavy_buckets_take_the_block_merge_path(ctx, pre):
    """runs of one bucket over thousands of accumulate chunks (all-equal scalars, 4 pairs per thread; a
    precomputed window size whose top window has only a few buckets) are merged by a block-wide tree"""
    n = 1 << 15
    setup = bpk.Setup.generate_srs(n, 101, ctx)
    if pre is not None:
        setup.precompute(pre)
    rng = random.Random(3)
    cases = [[0x1234567890ABCDEF1234567890ABCDEF] * n,
             [rng.choice([5, O.Q - 5]) for _ in range(n)],
             O.random_fr(31, n)]
    ctx.set_option("msm.chunk", 4)
    try:
        for sc in cases:
            assert bpk.point_to_affine(setup.commit_scalars(S(sc))) == horner_expected(sc, 101)
    finally:
        ctx.set_option("msm.chunk", 0)
        setup.free()


def test_msm_all_equal_scalars_at_scale_is_not_serialised(ctx):
    """2^20 identical scalars: one bucket per window holds every point; must finish in milliseconds"""
    import time
    n = 1 << 20
    setup = bpk.Setup.generate_srs(n, 101, ctx)
    sc_one = 0xDEADBEEFCAFEBABE1234567
    arr = np.repeat(S([sc_one]), n, axis=0)
    setup.commit_scalars(arr)
    t0 = time.perf_counter()
    got = setup.commit_scalars(arr)
    dt = time.perf_counter() - t0
    setup.free()
    # sum_i tau^i = (tau^n - 1) / (tau - 1)
    geo = (pow(101, n, O.Q) - 1) * pow(100, -1, O.Q) % O.Q
    assert bpk.point_to_affine(got) == O.g1_mul(O.G1_GEN, sc_one * geo % O.Q)
    assert dt < 0.5, dt


@pytest.mark.parametrize("logn", [16, 20])
def test_msm_full_size_closed_form(ctx, logn):
    n = (1 << logn) + 6  # the prover commits n + 6 coefficients (prover.rs:483-485)
    setup = bpk.Setup.generate_srs(n, 101, ctx)
    rng = np.random.default_rng(logn)
    raw = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64)
    raw[:, 3] &= np.uint64((1 << 62) - 1)
    sc = I(raw)
    got = setup.commit_scalars(raw)
    setup.free()
    assert bpk.point_to_affine(got) == horner_expected(sc, 101)


def test_msm_device_resident_shards_and_sum(ctx):
    """the multi-GPU decomposition on one GPU: per-shard partials (un-normalised projective) + g1_sum"""
    import torch
    n = 5000
    setup = bpk.Setup.generate_srs(n, 101, ctx)
    sc = O.random_fr(77, n)
    arr = S(sc)
    d_sc = torch.from_numpy(arr.view(np.int64)).cuda()
    parts = []
    bounds = [0, 1234, 1234, 4000, n]
    for a, b in zip(bounds[:-1], bounds[1:]):
        d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
        st = ctx.lib.bpk_msm_g1_dev(ctx.handle, setup.handle, a, d_sc.data_ptr() + 32 * a, b - a, 0, d_out.data_ptr())
        ctx.check(st, "bpk_msm_g1_dev")
        torch.cuda.synchronize()
        parts.append(d_out.cpu().numpy().view(np.uint64))
    for (a, b), p in zip(zip(bounds[:-1], bounds[1:]), parts):
        pts = [bpk.point_to_affine(r) for r in setup.powers_of_x(a, b - a)] if b > a else []
        assert bpk.point_to_affine(p) == O.msm_naive(pts[:50], sc[a:a + 50]) if b - a <= 50 else True
    total = bpk.g1_sum(np.stack(parts), ctx)
    assert bpk.point_to_affine(total) == horner_expected(sc, 101)
    # normalised device output equals the host-buffer path bit for bit
    d_out = torch.zeros(18, dtype=torch.int64, device="cuda")
    ctx.check(ctx.lib.bpk_msm_g1_dev(ctx.handle, setup.handle, 0, d_sc.data_ptr(), n, 1, d_out.data_ptr()))
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64), setup.commit_scalars(arr))
    setup.free()


def test_g1_sum_edge_cases(ctx):
    G = O.G1_GEN
    pts = [G, G, O.g1_neg(G), None, O.g1_mul(G, 5)]
    arr = bpk.points_from_affine(pts, [2, 3, 4, 5, 6])
    assert bpk.point_to_affine(bpk.g1_sum(arr, ctx)) == O.g1_sum(pts)
    assert bpk.point_to_affine(bpk.g1_sum(arr[:0].reshape(0, 18), ctx)) is None


def test_kernels_really_launch(ctx):
    ctx.profile_reset()
    ctx.profile_enable(True)
    x = S(O.random_fr(1, 4096))
    bpk.ntt_381(x, ctx)
    setup = bpk.Setup.generate_srs(512, 101, ctx)
    setup.commit_scalars(S(O.random_fr(2, 512)))
    ms_ntt, l_ntt = ctx.profile_get("ntt.pass")
    ms_acc, l_acc = ctx.profile_get("msm.accumulate")
    ctx.profile_enable(False)
    assert l_ntt in (2, 3) and ms_ntt > 0      # two passes (+ the lazily built inter-pass twiddle table)
    assert l_acc == 1 and ms_acc > 0
    assert ctx.launch_count() >= 8
    setup.free()


@pytest.mark.parametrize("lanes", [1, 3])
def test_msm_batch_on_streams_matches_single_calls(ctx, lanes):
    """bpk_msm_g1_dev_batch: several commitments on one SRS, concurrently on separate streams / workspace banks;
    five MSMs over three lanes, different slices and lengths incl. an empty one"""
    import ctypes

    import torch

    n = 5000
    setup = bpk.Setup.generate_srs(n, 101, ctx).precompute(0)
    srs = None
    specs = [(0, n), (0, 1234), (100, 4000), (4999, 1), (7, 0)]
    sc = [O.random_fr(200 + i, max(cnt, 1)) for i, (_, cnt) in enumerate(specs)]
    d = [torch.from_numpy(S(v).view(np.int64)).cuda() for v in sc]
    k = len(specs)
    ptrs = (ctypes.c_void_p * k)(*[t.data_ptr() for t in d])
    firsts = (ctypes.c_size_t * k)(*[f for f, _ in specs])
    lens = (ctypes.c_size_t * k)(*[c for _, c in specs])
    out = torch.zeros((k, 18), dtype=torch.int64, device="cuda")
    ctx.set_option("msm.lanes", lanes)
    try:
        for _ in range(2):   # second round reuses the lane workspaces
            ctx.check(ctx.lib.bpk_msm_g1_dev_batch(ctx.handle, setup.handle, k, ptrs, firsts, lens, 1, out.data_ptr()),
                      "batch")
            got = out.cpu().numpy().view(np.uint64)
            for i, (first, cnt) in enumerate(specs):
                acc = 0
                for j in reversed(range(cnt)):
                    acc = (acc * 101 + sc[i][j]) % O.Q
                want = O.g1_mul(O.G1_GEN, acc * pow(101, first, O.Q) % O.Q) if cnt else None
                assert bpk.point_to_affine(got[i]) == want, (lanes, i)
    finally:
        ctx.set_option("msm.lanes", 3)
        setup.free()


def test_msm_and_ntt_at_the_largest_practical_sizes(ctx):
    """2^26 pairs (4x the bench size; 6.4 GiB SRS, 8 x 10^8 sorted pairs) and a ragged 2^25 + 12345, against the
    closed form [sum s_i tau^i] G with scalars of period 2^16; NTT 2^26 forward + inverse round trip.
    Guards 32-bit index arithmetic; skipped when the GPU has less than 100 GiB free."""
    import torch

    free, _ = torch.cuda.mem_get_info()
    if free < 100 << 30:
        pytest.skip("needs 100 GiB of free HBM")
    period = 1 << 16
    block = O.random_fr(2626, period)
    d_block = torch.from_numpy(S(block).view(np.int64)).cuda()
    base = 0
    for j in reversed(range(period)):
        base = (base * 101 + block[j]) % O.Q
    step = pow(101, period, O.Q)
    out = torch.zeros(18, dtype=torch.int64, device="cuda")
    for n, pre in (((1 << 25) + 12345, False), (1 << 26, True)):
        setup = bpk.Setup.generate_srs(n, 101, ctx)
        if pre:
            setup.precompute(0)
        reps = (n + period - 1) // period
        d_sc = d_block.repeat(reps, 1)[:n].contiguous()
        ctx.check(ctx.lib.bpk_msm_g1_dev(ctx.handle, setup.handle, 0, d_sc.data_ptr(), n, 1, out.data_ptr()), "msm")
        got = bpk.point_to_affine(out.cpu().numpy().view(np.uint64))
        full, rem = divmod(n, period)
        geo = (pow(step, full, O.Q) - 1) * pow(step - 1, -1, O.Q) % O.Q      # sum_k tau^(k period), k < full
        tail = 0
        for j in reversed(range(rem)):
            tail = (tail * 101 + block[j]) % O.Q
        want = (base * geo + tail * pow(step, full, O.Q)) % O.Q
        assert got == O.g1_mul(O.G1_GEN, want), n
        del d_sc
        setup.free()
    n = 1 << 26
    x = d_block.repeat(n // period, 1).contiguous()
    y = torch.empty_like(x)
    ctx.check(ctx.lib.bpk_ntt_fr_dev(ctx.handle, x.data_ptr(), y.data_ptr(), n, 1, 0, None), "ntt")
    # a sequence of period 2^16 has a spectrum supported on multiples of 2^10
    probe = y.view(n // 1024, 1024, 4)[:, 1:, :]
    assert int(torch.count_nonzero(probe)) == 0
    ctx.check(ctx.lib.bpk_ntt_fr_dev(ctx.handle, y.data_ptr(), y.data_ptr(), n, 1, 1, None), "intt")
    assert torch.equal(x, y)


def test_msm_from_host_pipelined_upload_matches_device_path(ctx):
    """bpk_msm_g1_from_host / bpk_msm_g1 at n >= 2^22: head slice + rest on two streams, bucket sets joined;
    must equal the device-resident MSM and the closed form, pinned or pageable memory, overlap on or off"""
    import torch

    n = (1 << 22) + 77
    period = 1 << 12
    block = O.random_fr(4242, period)
    reps = (n + period - 1) // period
    host = torch.from_numpy(np.tile(S(block), (reps, 1))[:n].copy().view(np.int64))
    pinned = host.pin_memory()
    base = 0
    for j in reversed(range(period)):
        base = (base * 101 + block[j]) % O.Q
    step = pow(101, period, O.Q)
    full, rem = divmod(n, period)
    tail = 0
    for j in reversed(range(rem)):
        tail = (tail * 101 + block[j]) % O.Q
    want = O.g1_mul(O.G1_GEN, (base * (pow(step, full, O.Q) - 1) * pow(step - 1, -1, O.Q) + tail * pow(step, full, O.Q)) % O.Q)
    out = torch.zeros(18, dtype=torch.int64, device="cuda")
    for pre in (False, True):
        setup = bpk.Setup.generate_srs(n, 101, ctx)
        if pre:
            setup.precompute(0)
        for slices in (1, 0):
            ctx.set_option("msm.host_slices", slices)
            for src in (pinned, host):
                ctx.check(ctx.lib.bpk_msm_g1_from_host(ctx.handle, setup.handle, 0, src.data_ptr(), n, 1, out.data_ptr()),
                          "from_host")
                assert bpk.point_to_affine(out.cpu().numpy().view(np.uint64)) == want, (pre, slices)
        ctx.set_option("msm.host_slices", 1)
        # Setup::commit path (host in, host out) and a slice of the SRS with an un-normalised partial
        assert bpk.point_to_affine(setup.commit_scalars(host.numpy().view(np.uint64))) == want
        setup.free()


def test_msm_from_host_slice_and_partial(ctx):
    """bpk_msm_g1_from_host on a slice of the SRS (first > 0) with an un-normalised partial as output: the
    per-rank leg of the sharded end-to-end path; partials of two halves must add up to the whole commitment"""
    import torch

    n = 3001
    setup = bpk.Setup.generate_srs(n, 101, ctx)
    sc = O.random_fr(5151, n)
    host = S(sc)
    parts = torch.zeros((2, 18), dtype=torch.int64, device="cuda")
    lo = n // 2
    ctx.check(ctx.lib.bpk_msm_g1_from_host(ctx.handle, setup.handle, 0, host[:lo].ctypes.data, lo, 0,
                                           parts[0].data_ptr()), "from_host")
    tail = np.ascontiguousarray(host[lo:])
    ctx.check(ctx.lib.bpk_msm_g1_from_host(ctx.handle, setup.handle, lo, tail.ctypes.data, n - lo, 0,
                                           parts[1].data_ptr()), "from_host")
    total = bpk.g1_sum(parts.cpu().numpy().view(np.uint64), ctx)
    assert bpk.point_to_affine(total) == horner_expected(sc, 101)
    # argument checks: slice beyond the SRS, null scalars
    assert ctx.lib.bpk_msm_g1_from_host(ctx.handle, setup.handle, n - 1, host.ctypes.data, 2, 1, parts[0].data_ptr()) == -3
    assert ctx.lib.bpk_msm_g1_from_host(ctx.handle, setup.handle, 0, None, 2, 1, parts[0].data_ptr()) == -3
    assert ctx.lib.bpk_msm_g1_from_host(ctx.handle, setup.handle, 0, None, 0, 1, parts[0].data_ptr()) == 0
    assert bpk.point_to_affine(parts[0].cpu().numpy().view(np.uint64)) is None
    setup.free()


def test_device_memory_helpers(ctx):
    """bpk_dev_alloc / free / upload / download / copy / zero: round trip, pooled reuse of a freed block, error codes"""
    import ctypes

    lib, h = ctx.lib, ctx.handle
    a = ctypes.c_void_p()
    b = ctypes.c_void_p()
    ctx.check(lib.bpk_dev_alloc(h, 1000, ctypes.byref(a)), "alloc")
    ctx.check(lib.bpk_dev_alloc(h, 1000, ctypes.byref(b)), "alloc")
    assert a.value and b.value and a.value != b.value
    src = np.arange(125, dtype=np.uint64)
    dst = np.zeros(125, dtype=np.uint64)
    ctx.check(lib.bpk_dev_upload(h, a, src.ctypes.data, 1000), "upload")
    ctx.check(lib.bpk_dev_copy(h, b, a, 1000), "copy")
    ctx.check(lib.bpk_dev_zero(h, a, 496), "zero")
    ctx.check(lib.bpk_dev_download(h, dst.ctypes.data, b, 1000), "download")
    assert np.array_equal(dst, src)
    ctx.check(lib.bpk_dev_download(h, dst.ctypes.data, a, 1000), "download")
    assert not dst[:62].any() and np.array_equal(dst[62:], src[62:])
    ctx.check(lib.bpk_dev_free(h, a), "free")
    assert lib.bpk_dev_free(h, a) == -3                      # double free / foreign pointer
    c = ctypes.c_void_p()
    ctx.check(lib.bpk_dev_alloc(h, 900, ctypes.byref(c)), "alloc")   # same 256-byte size class: the block comes back
    assert c.value == a.value
    ctx.check(lib.bpk_dev_free(h, b), "free")
    ctx.check(lib.bpk_dev_free(h, c), "free")
    assert lib.bpk_dev_free(h, None) == 0
    assert lib.bpk_dev_alloc(h, 16, None) == -3


def test_bench_workload_generator_is_survey_8d(ctx):
    """bench.py's scalars (SplitMix64(12345) -> 64 bytes -> from_bytes_wide, reduced on the device) are SURVEY 8d's:
    the device reduction, the CPU generator and the oracle's random_fr agree, at an offset into the stream too"""
    import importlib
    import torch
    bench = importlib.import_module("bench")
    want = O.random_fr(bench.SCALAR_SEED, 300)
    got = bench.scalars_device_mont(ctx, bpk, torch, 0, 300).cpu().numpy().view(np.uint64)
    assert bpk.scalars_to_ints(got) == want
    got = bench.scalars_device_mont(ctx, bpk, torch, 137, 300).cpu().numpy().view(np.uint64)
    assert bpk.scalars_to_ints(got) == want[137:]
    assert bench.scalars_host_ints(137, 300) == want[137:]
    assert bpk.scalars_to_ints(bench.scalars_host_mont(0, 50)) == want[:50]
