// baby_plonk.hpp -- C++ host-side mirror of the reference's Rust call surfaces for the MSM / NTT hot
// path, layered on the C ABI of libbpk.so (include/bpk.h).  The reference is compiled code (Rust); its
// toolchain is absent from this image, so this header is the compiled-language host layer: same names,
// argument meaning and error behaviour as
//     BucketMSM::bucket_msm            src/msm.rs:76-118
//     Setup::{generate_srs, commit}    src/setup.rs:12-37
//     ntt_381 / i_ntt_381              src/utils.rs:63-81, 106-129
//     root_of_unity / roots_of_unity / find_next_power_of_two   src/utils.rs:39-61
//     Polynomial {values, basis}, ntt, i_ntt, operator*          src/polynomial.rs:14-55, 189-276
// (the Rust extern "C" shim that replaces this header in the reference tree is in INTEGRATION.md).
// Rust panics (assert!, slice bounds, todo!()) become baby_plonk::Panic exceptions.
//
// Scalar / G1Projective have the reference's in-memory layout (Montgomery u64 limbs), so vectors of
// them are handed to the C ABI as they are.  The few host-side scalar helpers (Scalar::from, pow, the
// G1 scalar multiplication used to build expected values in tests) reuse the library's own field / curve
// templates in their host build (csrc/ff.cuh, csrc/ec.cuh); every bulk operation runs on the GPU.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/bpk.h"
#include "../csrc/ec.cuh"

namespace baby_plonk {

struct Panic : std::runtime_error {
    explicit Panic(const std::string& m) : std::runtime_error("panicked: " + m) {}
};

inline bpk_ctx* ctx() {  // one context per process (one process per GPU)
    static bpk_ctx* c = [] {
        bpk_ctx* p = nullptr;
        const char* lr = std::getenv("LOCAL_RANK");
        int st = bpk_init(&p, lr ? std::atoi(lr) : 0);
        if (st != 0) throw Panic(std::string("bpk_init: ") + bpk_strerror(st));
        return p;
    }();
    return c;
}
inline void check(int st, const char* what) {
    if (st != 0) throw Panic(std::string(what) + ": " + bpk_strerror(st));
}

// ---- bls12_381::Scalar (lib/bls12_381/src/scalar.rs:22) -------------------------------------------
struct Scalar {
    uint64_t l[4];
    static bpk::fr_t to_fe(const Scalar& s) {
        bpk::fr_t f;
        for (int i = 0; i < 4; i++) { f.l[2 * i] = (uint32_t)s.l[i]; f.l[2 * i + 1] = (uint32_t)(s.l[i] >> 32); }
        return f;
    }
    static Scalar from_fe(const bpk::fr_t& f) {
        Scalar s;
        for (int i = 0; i < 4; i++) s.l[i] = (uint64_t)f.l[2 * i] | ((uint64_t)f.l[2 * i + 1] << 32);
        return s;
    }
    static Scalar zero() { return from_fe(bpk::fr_t::zero()); }
    static Scalar one() { return from_fe(bpk::fr_t::one()); }
    static Scalar from(uint64_t v) {  // scalar.rs:282-284: v * R2 / R
        bpk::fr_t f = bpk::fr_t::zero();
        f.l[0] = (uint32_t)v;
        f.l[1] = (uint32_t)(v >> 32);
        return from_fe(bpk::to_mont(f));
    }
    Scalar operator*(const Scalar& o) const { return from_fe(bpk::mul(to_fe(*this), to_fe(o))); }
    Scalar operator+(const Scalar& o) const { return from_fe(bpk::add(to_fe(*this), to_fe(o))); }
    Scalar operator-(const Scalar& o) const { return from_fe(bpk::sub(to_fe(*this), to_fe(o))); }
    Scalar neg() const { return from_fe(bpk::neg(to_fe(*this))); }
    Scalar invert() const { return from_fe(bpk::inv(to_fe(*this))); }  // scalar.rs:394-..., zero maps to zero
    Scalar pow(uint64_t e) const { return from_fe(bpk::pow_u64(to_fe(*this), e)); }  // pow(&[e,0,0,0])
    bool operator==(const Scalar& o) const { return std::memcmp(l, o.l, sizeof l) == 0; }
    bool operator!=(const Scalar& o) const { return !(*this == o); }
    // scalar.rs:208-213
    static Scalar ROOT_OF_UNITY() {
        return Scalar{{0xb9b58d8c5f0e466aull, 0x5b1b4c801819d7ecull, 0x0af53ae352a31e64ull, 0x5bf3adda19e9b27bull}};
    }
};

// ---- bls12_381::G1Projective (lib/bls12_381/src/g1.rs:442-446) -------------------------------------
struct G1Projective {
    uint64_t x[6], y[6], z[6];
    static bpk::fp_t fe(const uint64_t* w) {
        bpk::fp_t f;
        for (int i = 0; i < 6; i++) { f.l[2 * i] = (uint32_t)w[i]; f.l[2 * i + 1] = (uint32_t)(w[i] >> 32); }
        return f;
    }
    static void put(uint64_t* w, const bpk::fp_t& f) {
        for (int i = 0; i < 6; i++) w[i] = (uint64_t)f.l[2 * i] | ((uint64_t)f.l[2 * i + 1] << 32);
    }
    static G1Projective identity() {  // g1.rs:605-611
        G1Projective p;
        put(p.x, bpk::fp_t::zero());
        put(p.y, bpk::fp_t::one());
        put(p.z, bpk::fp_t::zero());
        return p;
    }
    static G1Projective generator() {  // g1.rs:615-635
        G1Projective p;
        const uint64_t gx[6] = {0x5cb38790fd530c16ull, 0x7817fc679976fff5ull, 0x154f95c7143ba1c1ull,
                                0xf0ae6acdf3d0e747ull, 0xedce6ecc21dbf440ull, 0x120177419e0bfb75ull};
        const uint64_t gy[6] = {0xbaac93d50ce72271ull, 0x8c22631a7918fd8eull, 0xdd595f13570725ceull,
                                0x51ac582950405194ull, 0x0e1c8c3fad0059c0ull, 0x0bbc3efc5008a26aull};
        std::memcpy(p.x, gx, sizeof gx);
        std::memcpy(p.y, gy, sizeof gy);
        put(p.z, bpk::fp_t::one());
        return p;
    }
    bpk::xyzz_t to_xyzz() const {  // x = X/Z = XZ/Z^2, y = Y/Z = YZ^2/Z^3
        bpk::fp_t X = fe(x), Y = fe(y), Z = fe(z);
        if (Z.is_zero()) return bpk::xyzz_t::inf();
        bpk::xyzz_t q;
        q.ZZ = bpk::sqr(Z);
        q.ZZZ = bpk::mul(q.ZZ, Z);
        q.X = bpk::mul(X, Z);
        q.Y = bpk::mul(Y, q.ZZ);
        return q;
    }
    static G1Projective from_xyzz(const bpk::xyzz_t& q) {
        if (q.is_inf()) return identity();
        bpk::affine_t a = bpk::xyzz_to_affine(q);
        G1Projective p;
        put(p.x, a.x);
        put(p.y, a.y);
        put(p.z, bpk::fp_t::one());
        return p;
    }
    // host-side group law: only for building expected values, as the reference's tests do
    G1Projective operator+(const G1Projective& o) const {
        bpk::xyzz_t a = to_xyzz();
        bpk::xyzz_add(a, o.to_xyzz());
        return from_xyzz(a);
    }
    G1Projective operator*(const Scalar& s) const {  // g1.rs:754-774 double-and-add, MSB first
        bpk::fr_t k = bpk::from_mont(Scalar::to_fe(s));
        bpk::xyzz_t acc = bpk::xyzz_t::inf(), base = to_xyzz();
        for (int i = 255; i >= 0; i--) {
            bpk::xyzz_dbl(acc);
            if ((k.l[i >> 5] >> (i & 31)) & 1u) bpk::xyzz_add(acc, base);
        }
        return from_xyzz(acc);
    }
    // g1.rs:479-496: projective equivalence
    bool operator==(const G1Projective& o) const {
        bpk::fp_t X1 = fe(x), Y1 = fe(y), Z1 = fe(z), X2 = fe(o.x), Y2 = fe(o.y), Z2 = fe(o.z);
        bool i1 = Z1.is_zero(), i2 = Z2.is_zero();
        if (i1 || i2) return i1 && i2;
        return bpk::mul(X1, Z2) == bpk::mul(X2, Z1) && bpk::mul(Y1, Z2) == bpk::mul(Y2, Z1);
    }
    bool operator!=(const G1Projective& o) const { return !(*this == o); }
};
static_assert(sizeof(Scalar) == 32 && sizeof(G1Projective) == 144, "layouts must match the C ABI");

// ---- src/utils.rs ---------------------------------------------------------------------------------
inline bool is_power_of_two(uint64_t n) { return n != 0 && (n & (n - 1)) == 0; }  // utils.rs:82-84
inline Scalar root_of_unity(uint64_t group_order) {                                // utils.rs:39-43
    return Scalar::ROOT_OF_UNITY().pow(((uint64_t)1 << 32) / group_order);
}
inline std::vector<Scalar> roots_of_unity(uint64_t group_order) {                  // utils.rs:45-52
    std::vector<Scalar> res{Scalar::from(1)};
    Scalar g = root_of_unity(group_order);
    for (uint64_t i = 1; i < group_order; i++) res.push_back(res.back() * g);
    return res;
}
inline size_t find_next_power_of_two(size_t n, size_t m) {                         // utils.rs:54-61
    size_t power = 1, target = n + m + 1;
    while (power < target) power <<= 1;
    return power;
}
inline std::vector<Scalar> ntt_381(const std::vector<Scalar>& elements) {          // utils.rs:63-81
    if (!is_power_of_two(elements.size())) throw Panic("assertion failed: is_power_of_two(n)");
    std::vector<Scalar> out(elements.size());
    check(bpk_ntt_fr(ctx(), elements[0].l, out[0].l, elements.size(), 1), "ntt_381");
    return out;
}
inline std::vector<Scalar> i_ntt_381(const std::vector<Scalar>& elements) {        // utils.rs:106-129
    if (!is_power_of_two(elements.size())) throw Panic("assertion failed: is_power_of_two(n)");
    std::vector<Scalar> out(elements.size());
    check(bpk_intt_fr(ctx(), elements[0].l, out[0].l, elements.size(), 1), "i_ntt_381");
    return out;
}

// ---- src/polynomial.rs ----------------------------------------------------------------------------
enum class Basis { Lagrange, Monomial };
struct Polynomial {
    std::vector<Scalar> values;
    Basis basis;
    Polynomial(std::vector<Scalar> v, Basis b) : values(std::move(v)), basis(b) {}
    Polynomial ntt() const {    // polynomial.rs:47-51
        if (basis != Basis::Monomial) throw Panic("assertion `left == right` failed: basis == Monomial");
        return Polynomial(ntt_381(values), Basis::Lagrange);
    }
    Polynomial i_ntt() const {  // polynomial.rs:52-55
        if (basis != Basis::Lagrange) throw Panic("assertion `left == right` failed: basis == Lagrange");
        return Polynomial(i_ntt_381(values), Basis::Monomial);
    }
    Polynomial operator*(const Polynomial& rhs) const {  // polynomial.rs:189-276
        if (basis != rhs.basis) throw Panic("assertion `left == right` failed: self.basis == rhs.basis");
        if (basis == Basis::Lagrange) throw Panic("not yet implemented");  // todo!()
        std::vector<Scalar> out(values.size() + rhs.values.size() - 1);
        check(bpk_poly_mul_fr(ctx(), values[0].l, values.size(), rhs.values[0].l, rhs.values.size(), out[0].l),
              "Polynomial::mul");
        return Polynomial(std::move(out), Basis::Monomial);
    }
    bool operator==(const Polynomial& o) const { return basis == o.basis && values == o.values; }
};

// ---- src/msm.rs -----------------------------------------------------------------------------------
struct BucketMSM {
    static G1Projective bucket_msm(const std::vector<G1Projective>& points, const std::vector<Scalar>& scalars,
                                   size_t b, size_t c) {
        size_t n = points.size() < scalars.size() ? points.size() : scalars.size();
        uint64_t h = 0;
        check(bpk_srs_load(ctx(), n ? points[0].x : nullptr, n, &h), "bucket_msm: srs_load");
        G1Projective out = G1Projective::identity();
        int st = bpk_bucket_msm(ctx(), h, scalars.empty() ? nullptr : scalars[0].l, scalars.size(), b, c, out.x);
        bpk_srs_free(ctx(), h);
        check(st, "bucket_msm");
        return out;
    }
};

// ---- src/setup.rs ---------------------------------------------------------------------------------
struct Setup {
    std::vector<G1Projective> powers_of_x;  // host copy, as in the reference (read back from the GPU)
    uint64_t handle = 0;                    // the same points resident on the GPU
    static Setup generate_srs(size_t powers, const Scalar& tau) {  // setup.rs:12-31 (G1 part)
        Setup s;
        check(bpk_srs_generate(ctx(), tau.l, powers, &s.handle), "generate_srs");
        s.powers_of_x.resize(powers);
        if (powers) check(bpk_srs_read(ctx(), s.handle, 0, powers, s.powers_of_x[0].x), "generate_srs: read");
        return s;
    }
    // optional: keep the window levels [2^(c w)] P_i next to the SRS (bpk_srs_precompute); results are unchanged
    void precompute(unsigned window_bits = 0) const { check(bpk_srs_precompute(ctx(), handle, window_bits), "precompute"); }
    G1Projective commit(const Polynomial& polynomial) const {      // setup.rs:32-37
        if (polynomial.basis != Basis::Monomial) throw Panic("assertion `left == right` failed: basis == Monomial");
        G1Projective out = G1Projective::identity();
        check(bpk_bucket_msm(ctx(), handle, polynomial.values.empty() ? nullptr : polynomial.values[0].l,
                             polynomial.values.size(), 256, 4, out.x),
              "commit");
        return out;
    }
};

}  // namespace baby_plonk
