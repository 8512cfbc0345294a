// BLS12-381 G1 group operations for the MSM kernels (y^2 = x^3 + 4, a = 0).
//
// The reference keeps points as homogeneous projective (X:Y:Z) and adds with the complete
// formulas of ePrint 2015/1060 (lib/bls12_381/src/g1.rs:638-752).  G1Projective is not a unique
// representation, so parity is defined on G1Affine::from (g1.rs:49-63).  The kernels therefore
// use their own coordinates -- extended Jacobian "XYZZ" (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2)
// accumulators with affine addends -- and handle every degenerate case (identity operands,
// P + P, P + (-P)) explicitly, because the reference's tests run on SRS with tau = 1 / 2 where
// all points coincide (src/prover.rs:684, src/setup.rs:47).
#pragma once
#include "ff.cuh"

namespace bpk {

// affine point, Montgomery coordinates; (0, 0) encodes the point at infinity (not on the curve)
struct alignas(16) affine_t {
    fp_t x, y;
    BPK_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    BPK_HD static affine_t inf() {
        affine_t r;
        r.x = fp_t::zero();
        r.y = fp_t::zero();
        return r;
    }
};

// extended Jacobian accumulator; ZZ == 0 encodes the identity
struct alignas(16) xyzz_t {
    fp_t X, Y, ZZ, ZZZ;
    BPK_HD bool is_inf() const { return ZZ.is_zero(); }
    BPK_HD static xyzz_t inf() {
        xyzz_t r;
        r.X = fp_t::zero();
        r.Y = fp_t::zero();
        r.ZZ = fp_t::zero();
        r.ZZZ = fp_t::zero();
        return r;
    }
    BPK_HD static xyzz_t from_affine(const affine_t& a) {
        if (a.is_inf()) return inf();
        xyzz_t r;
        r.X = a.x;
        r.Y = a.y;
        r.ZZ = fp_t::one();
        r.ZZZ = fp_t::one();
        return r;
    }
};

// p <- 2p   (dbl-2008-s-1, a = 0).  Y == 0 or identity input yields ZZ == 0 (identity).
BPK_HD void xyzz_dbl(xyzz_t& p) {
    fp_t U = dbl(p.Y);
    fp_t V = sqr(U);
    fp_t W = mul(U, V);
    fp_t S = mul(p.X, V);
    fp_t X2 = sqr(p.X);
    fp_t M = add(dbl(X2), X2);
    fp_t X3 = sub(sqr(M), dbl(S));
    fp_t Y3 = sub(mul(M, sub(S, X3)), mul(W, p.Y));
    p.ZZ = mul(V, p.ZZ);
    p.ZZZ = mul(W, p.ZZZ);
    p.X = X3;
    p.Y = Y3;
}

// 2 * (affine, not infinity) -> XYZZ   (mdbl-2008-s-1)
BPK_HD xyzz_t affine_dbl(const affine_t& a) {
    xyzz_t r;
    fp_t U = dbl(a.y);
    fp_t V = sqr(U);
    fp_t W = mul(U, V);
    fp_t S = mul(a.x, V);
    fp_t X2 = sqr(a.x);
    fp_t M = add(dbl(X2), X2);
    r.X = sub(sqr(M), dbl(S));
    r.Y = sub(mul(M, sub(S, r.X)), mul(W, a.y));
    r.ZZ = V;
    r.ZZZ = W;
    return r;
}

// p <- p + q, q affine   (madd-2008-s) with all degenerate cases
BPK_HD void xyzz_madd(xyzz_t& p, const affine_t& q) {
    if (q.is_inf()) return;
    if (p.is_inf()) {
        p.X = q.x;
        p.Y = q.y;
        p.ZZ = fp_t::one();
        p.ZZZ = fp_t::one();
        return;
    }
    fp_t U2 = mul(q.x, p.ZZ);
    fp_t S2 = mul(q.y, p.ZZZ);
    fp_t Pd = sub(U2, p.X);
    fp_t Rd = sub(S2, p.Y);
    if (Pd.is_zero()) {
        if (Rd.is_zero())
            p = affine_dbl(q);  // same point
        else
            p = xyzz_t::inf();  // opposite points
        return;
    }
    fp_t PP = sqr(Pd);
    fp_t PPP = mul(Pd, PP);
    fp_t Qv = mul(p.X, PP);
    fp_t X3 = sub(sub(sqr(Rd), PPP), dbl(Qv));
    fp_t Y3 = sub(mul(Rd, sub(Qv, X3)), mul(p.Y, PPP));
    p.ZZ = mul(p.ZZ, PP);
    p.ZZZ = mul(p.ZZZ, PPP);
    p.X = X3;
    p.Y = Y3;
}

// p <- p + q   (add-2008-s) with all degenerate cases
BPK_HD void xyzz_add(xyzz_t& p, const xyzz_t& q) {
    if (q.is_inf()) return;
    if (p.is_inf()) {
        p = q;
        return;
    }
    fp_t U1 = mul(p.X, q.ZZ);
    fp_t U2 = mul(q.X, p.ZZ);
    fp_t S1 = mul(p.Y, q.ZZZ);
    fp_t S2 = mul(q.Y, p.ZZZ);
    fp_t Pd = sub(U2, U1);
    fp_t Rd = sub(S2, S1);
    if (Pd.is_zero()) {
        if (Rd.is_zero())
            xyzz_dbl(p);
        else
            p = xyzz_t::inf();
        return;
    }
    fp_t PP = sqr(Pd);
    fp_t PPP = mul(Pd, PP);
    fp_t Qv = mul(U1, PP);
    fp_t X3 = sub(sub(sqr(Rd), PPP), dbl(Qv));
    fp_t Y3 = sub(mul(Rd, sub(Qv, X3)), mul(S1, PPP));
    p.ZZ = mul(mul(p.ZZ, q.ZZ), PP);
    p.ZZZ = mul(mul(p.ZZZ, q.ZZZ), PPP);
    p.X = X3;
    p.Y = Y3;
}

// ------------------------------------------------------------------------------------------
// affine + affine with the slope's denominator inverted OUTSIDE (batched: Montgomery's trick over many
// independent additions, msm.cu).  prepare() classifies the pair and yields the denominator that has to be
// inverted (1 where none is needed, so that a batch product is never poisoned); finish() completes the
// addition from its inverse: 1 M (slope) + 1 S + 1 M, against 8 M + 2 S of the XYZZ mixed addition.
// All degenerate cases are explicit, as in xyzz_madd: the reference's tests run on SRS where all points coincide.
// ------------------------------------------------------------------------------------------
enum : int { AFF_COPY_P = 0, AFF_COPY_Q = 1, AFF_ADD = 2, AFF_DBL = 3, AFF_INF = 4 };

BPK_HD int affine_add_prepare(const affine_t& p, const affine_t& q, fp_t& den) {
    den = fp_t::one();
    if (q.is_inf()) return AFF_COPY_P;
    if (p.is_inf()) return AFF_COPY_Q;
    fp_t dx = sub(q.x, p.x);
    if (!dx.is_zero()) {
        den = dx;
        return AFF_ADD;
    }
    if (p.y != q.y) return AFF_INF;  // opposite points
    fp_t dy = dbl(p.y);              // same point: slope 3 x^2 / 2 y
    if (dy.is_zero()) return AFF_INF;  // y == 0 cannot occur on E(Fp) (odd group order); malformed input stays harmless
    den = dy;
    return AFF_DBL;
}

BPK_HD affine_t affine_add_finish(int kind, const affine_t& p, const affine_t& q, const fp_t& den_inv) {
    if (kind == AFF_COPY_P) return p;
    if (kind == AFF_COPY_Q) return q;
    if (kind == AFF_INF) return affine_t::inf();
    fp_t lam;
    if (kind == AFF_ADD) {
        lam = mul(sub(q.y, p.y), den_inv);
    } else {
        fp_t xx = sqr(p.x);
        lam = mul(add(dbl(xx), xx), den_inv);
    }
    affine_t r;
    r.x = sub(sub(sqr(lam), p.x), q.x);
    r.y = sub(mul(lam, sub(p.x, r.x)), p.y);
    return r;
}

// the same with the products left in [0, 2p) ("lazy": five conditional subtractions fewer); den_inv may be anywhere in
// [0, 2p), the operands are canonical, the result is canonical
BPK_HD affine_t affine_add_finish_lazy(int kind, const affine_t& p, const affine_t& q, const fp_t& den_inv) {
    if (kind == AFF_COPY_P) return p;
    if (kind == AFF_COPY_Q) return q;
    if (kind == AFF_INF) return affine_t::inf();
    fp_t lam;
    if (kind == AFF_ADD) {
        lam = mul_lazy(sub(q.y, p.y), den_inv);
    } else {
        fp_t xx = sqr(p.x);
        lam = mul_lazy(add(dbl(xx), xx), den_inv);
    }
    const fp_t x3 = sub_lazy(sub_lazy(sqr_lazy(lam), p.x), q.x);
    const fp_t y3 = sub_lazy(mul_lazy(lam, sub_lazy(p.x, x3)), p.y);
    affine_t r;
    r.x = reduce_once(x3);
    r.y = reduce_once(y3);
    return r;
}

// a + b for two affine points (neither the identity, a.x != b.x) -> XYZZ: add-2008-s with both ZZ = ZZZ = 1,
// 4M + 2S instead of 12M + 2S.  Used where both operands are known to be affine (first level of the bucket tree).
BPK_HD xyzz_t xyzz_from_affine_sum(const fp_t& ax, const fp_t& ay, const fp_t& bx, const fp_t& by) {
    const fp_t Pd = sub(bx, ax), Rd = sub(by, ay);
    const fp_t PP = sqr(Pd);
    const fp_t PPP = mul(Pd, PP);
    const fp_t Qv = mul(ax, PP);
    xyzz_t r;
    r.X = sub(sub(sqr(Rd), PPP), dbl(Qv));
    r.Y = sub(mul(Rd, sub(Qv, r.X)), mul(ay, PPP));
    r.ZZ = PP;
    r.ZZZ = PPP;
    return r;
}

BPK_HD affine_t affine_neg(const affine_t& a) {
    affine_t r;
    r.x = a.x;
    r.y = neg(a.y);  // neg(0) == 0 keeps the infinity encoding
    return r;
}

BPK_HD xyzz_t xyzz_neg(const xyzz_t& a) {
    xyzz_t r = a;
    r.Y = neg(a.Y);
    return r;
}

// G1Affine::from(&G1Projective): x = X/Z, y = Y/Z (g1.rs:49-63)
BPK_HD affine_t proj_to_affine(const fp_t& X, const fp_t& Y, const fp_t& Z) {
    if (Z.is_zero()) return affine_t::inf();
    fp_t zi = inv(Z);
    affine_t r;
    r.x = mul(X, zi);
    r.y = mul(Y, zi);
    return r;
}

BPK_HD affine_t xyzz_to_affine(const xyzz_t& p) {
    if (p.is_inf()) return affine_t::inf();
    fp_t i = inv(mul(p.ZZ, p.ZZZ));
    affine_t r;
    r.x = mul(p.X, mul(i, p.ZZZ));
    r.y = mul(p.Y, mul(i, p.ZZ));
    return r;
}

// [k] p by double-and-add, MSB first (k as nbits-bit little-endian u32 limbs)
BPK_HD xyzz_t xyzz_mul_small(const xyzz_t& p, uint64_t k) {
    xyzz_t acc = xyzz_t::inf();
#pragma unroll 1
    for (int b = 63; b >= 0; b--) {
        xyzz_dbl(acc);
        if ((k >> b) & 1ull) xyzz_add(acc, p);
    }
    return acc;
}

}  // namespace bpk
