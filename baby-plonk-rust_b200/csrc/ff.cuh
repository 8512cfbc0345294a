// Montgomery prime-field arithmetic on 32-bit limbs for sm_100a.
//
// Mirrors the VALUES of the reference's field types, not their code:
//   Fr  = bls12_381::Scalar  (lib/bls12_381/src/scalar.rs:22, 4 x u64 Montgomery, R = 2^256)
//   Fp  = bls12_381::fp::Fp  (lib/bls12_381/src/fp.rs:15,     6 x u64 Montgomery, R = 2^384)
// A u64 limb array in little-endian memory order is bit-identical to a u32 limb array of twice
// the length, so the kernels read Rust's in-memory representation unchanged.  Every public
// operation returns a fully reduced value (< modulus), so limb equality == field equality and
// results are bit-identical to scalar.rs:514-635 / fp.rs:361-660.
//
// The multiplier is written as carry-chained PTX (mad.lo.cc / madc.hi.cc); ptxas fuses each
// lo/hi pair on an aligned register pair into one IMAD.WIDE.U32(.X).  Products of even and odd
// limbs are accumulated in two separate arrays so that both carry chains stay pair-aligned.
//
// Every primitive also has a plain-C host path (carry flag emulated in a thread_local) so that
// the very same templates are unit-tested on the CPU in the build container (tests/host/ff_host_check.cpp,
// driven by tests/test_host_templates.py).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define BPK_HD __host__ __device__ __forceinline__
#define BPK_D __device__ __forceinline__
#else
#define BPK_HD inline
#define BPK_D inline
#endif

namespace bpk {

// ------------------------------------------------------------------------------------------
// carry-flag primitives
// ------------------------------------------------------------------------------------------
namespace ptx {
#if !defined(__CUDA_ARCH__)
static thread_local uint32_t g_cc = 0;  // host emulation of CC.CF
#endif

BPK_HD uint32_t add_cc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    uint64_t t = (uint64_t)a + b; g_cc = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
BPK_HD uint32_t addc_cc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    uint64_t t = (uint64_t)a + b + g_cc; g_cc = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
BPK_HD uint32_t addc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return (uint32_t)((uint64_t)a + b + g_cc);
#endif
}
BPK_HD uint32_t sub_cc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    uint64_t t = (uint64_t)a - b; g_cc = (uint32_t)(t >> 63); return (uint32_t)t;
#endif
}
BPK_HD uint32_t subc_cc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    uint64_t t = (uint64_t)a - b - g_cc; g_cc = (uint32_t)(t >> 63); return (uint32_t)t;
#endif
}
BPK_HD uint32_t subc(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return (uint32_t)((uint64_t)a - b - g_cc);
#endif
}
BPK_HD uint32_t mul_lo(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return (uint32_t)((uint64_t)a * b);
#endif
}
BPK_HD uint32_t mul_hi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
BPK_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    uint64_t t = (uint64_t)(uint32_t)((uint64_t)a * b) + c; g_cc = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
BPK_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    uint64_t t = (uint64_t)(uint32_t)((uint64_t)a * b) + c + g_cc; g_cc = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
BPK_HD uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    uint64_t t = (((uint64_t)a * b) >> 32) + c; g_cc = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
BPK_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    uint64_t t = (((uint64_t)a * b) >> 32) + c + g_cc; g_cc = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
BPK_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    return (uint32_t)((((uint64_t)a * b) >> 32) + c + g_cc);
#endif
}
}  // namespace ptx

// ------------------------------------------------------------------------------------------
// field parameters
// ------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
// -q^-1 mod 2^32 = 0xffffffff, fetched from memory by the kernels (see FrParams::m0_opaque)
static __device__ uint32_t FR_M0_GLOBAL = 0xffffffffu;
#endif

// scalar.rs:83-88 (MODULUS), :164 (INV -> low 32 bits), :167-172 (R), :175-180 (R2)
struct FrParams {
    static constexpr int N = 8;
    static constexpr uint32_t M0 = 0xffffffffu;  // -q^-1 mod 2^32
#ifndef BPK_FR_SPECIAL
#define BPK_FR_SPECIAL 1
#endif
    static constexpr bool SPECIAL_LOW64 = BPK_FR_SPECIAL != 0;  // q mod 2^64 == 2^64 - 2^32 + 1 and M0 == -1
    BPK_HD static uint32_t m0_opaque() {
#if defined(__CUDA_ARCH__)
        uint32_t v;
        asm("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(&FR_M0_GLOBAL));
        return v;
#else
        return M0;
#endif
    }
    BPK_HD static constexpr uint32_t modk(int i) { return mod(i); }  // (unused: SPECIAL_LOW64 path)
    BPK_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                   0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return m[i];
    }
    BPK_HD static constexpr uint32_t one(int i) {  // R = 2^256 mod q
        constexpr uint32_t m[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                                   0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return m[i];
    }
    BPK_HD static constexpr uint32_t r2(int i) {  // R^2 mod q
        constexpr uint32_t m[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu,
                                   0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return m[i];
    }
    static constexpr int BITS = 255;  // bit length of q
};

// fp.rs:70-77 (MODULUS), :80 (INV), :83-90 (R), :93-100 (R2)
struct FpParams {
    static constexpr int N = 12;
    static constexpr uint32_t M0 = 0xfffcfffdu;  // -p^-1 mod 2^32
    static constexpr bool SPECIAL_LOW64 = false;
    BPK_HD static constexpr uint32_t m0_opaque() { return M0; }
    BPK_HD static constexpr uint32_t modk(int i) { return mod(i); }  // immediates fuse fine for p
    BPK_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu,
                                    0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu, 0x64774b84u,
                                    0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
        return m[i];
    }
    BPK_HD static constexpr uint32_t one(int i) {  // R = 2^384 mod p
        constexpr uint32_t m[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu,
                                    0x53c758bau, 0x5f489857u, 0x70525745u, 0x77ce5853u,
                                    0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
        return m[i];
    }
    BPK_HD static constexpr uint32_t r2(int i) {  // R^2 mod p
        constexpr uint32_t m[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u,
                                    0x4c95b6d5u, 0x8de5476cu, 0x939d83c0u, 0x67eb88a9u,
                                    0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
        return m[i];
    }
    static constexpr int BITS = 381;  // bit length of p
};

// ------------------------------------------------------------------------------------------
// the field element
// ------------------------------------------------------------------------------------------
template <class P>
struct alignas(16) Fe {
    typedef P params;
    static constexpr int N = P::N;
    uint32_t l[N];

    BPK_HD static Fe zero() {
        Fe r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    BPK_HD static Fe one() {  // Montgomery one
        Fe r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::one(i);
        return r;
    }
    BPK_HD static Fe r2() {
        Fe r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::r2(i);
        return r;
    }
    BPK_HD bool is_zero() const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= l[i];
        return acc == 0;
    }
    BPK_HD bool operator==(const Fe& o) const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= (l[i] ^ o.l[i]);
        return acc == 0;
    }
    BPK_HD bool operator!=(const Fe& o) const { return !(*this == o); }
};

namespace detail {

// 32 x 32 -> 64 product with no addend: a plain IMAD.WIDE.U32 Rd, Ra, Rb, RZ (one heavy-pipe slot;
// any 64-bit or carry-in addend makes it two, profiles/r1_imad_forms.md)
BPK_HD void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("{ .reg .b64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
#else
    uint64_t t = (uint64_t)a * b;
    lo = (uint32_t)t;
    hi = (uint32_t)(t >> 32);
#endif
}

// acc[j], acc[j+1] = a[j] * bi for even j
template <int N>
BPK_HD void mul_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
#pragma unroll
    for (int j = 0; j < N; j += 2) mul_wide(acc[j], acc[j + 1], a[j], bi);
}

// (acc[j], acc[j+1]) += a[j] * bi for even j, one carry chain; the carry-out stays in CC
template <int N>
BPK_HD void cmad_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
    acc[0] = ptx::mad_lo_cc(a[0], bi, acc[0]);
    acc[1] = ptx::madc_hi_cc(a[0], bi, acc[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        acc[j] = ptx::madc_lo_cc(a[j], bi, acc[j]);
        acc[j + 1] = ptx::madc_hi_cc(a[j], bi, acc[j + 1]);
    }
}

// same, with the modulus limbs mod(j + OFF) as immediates
template <class P, int OFF>
BPK_HD void cmad_mod(uint32_t* acc, uint32_t mi) {
    constexpr int N = P::N;
    acc[0] = ptx::mad_lo_cc(P::modk(OFF), mi, acc[0]);
    acc[1] = ptx::madc_hi_cc(P::modk(OFF), mi, acc[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        acc[j] = ptx::madc_lo_cc(P::modk(j + OFF), mi, acc[j]);
        acc[j + 1] = ptx::madc_hi_cc(P::modk(j + OFF), mi, acc[j + 1]);
    }
}

// odd[] <- (odd[] >> 64) + a[j] * bi, consuming the incoming carry
template <int N>
BPK_HD void madc_n_rshift(uint32_t* odd, const uint32_t* a, uint32_t bi) {
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
        odd[j] = ptx::madc_lo_cc(a[j], bi, odd[j + 2]);
        odd[j + 1] = ptx::madc_hi_cc(a[j], bi, odd[j + 3]);
    }
    odd[N - 2] = ptx::madc_lo_cc(a[N - 2], bi, 0);
    odd[N - 1] = ptx::madc_hi(a[N - 2], bi, 0);
}

// one row of the interleaved multiply + Montgomery reduction.
// value = sum even[j] 2^(32j) + sum odd[j] 2^(32(j+1)); on exit even[0] == 0
template <class P, bool FIRST>
BPK_HD void mad_n_redc(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi) {
    constexpr int N = P::N;
    if (FIRST) {
        mul_n<N>(odd, a + 1, bi);
        mul_n<N>(even, a, bi);
    } else {
        even[0] = ptx::add_cc(even[0], odd[1]);
        madc_n_rshift<N>(odd, a + 1, bi);
        cmad_n<N>(even, a, bi);
        odd[N - 1] = ptx::addc(odd[N - 1], 0);
    }
    if (P::SPECIAL_LOW64) {
        // Fr: q = q_hi 2^64 + (2^64 - 2^32 + 1) and -q^-1 = -1 (mod 2^32), so m = -even[0] and
        //   V + m q = (V - even[0]) + 2^32 Y + 2^64 m q_hi,   Y = m 2^32 - (m - c0),  c0 = [even[0] != 0]
        // (even[0] + m = c0 2^32).  Only the six q_hi limbs need real products -- 14 instead of 16 wide
        // IMADs per row -- and the two awkward limbs 0x00000001 / 0xffffffff, whose strength reduction by
        // ptxas otherwise un-fuses both m*q carry chains, never appear as multiplier operands.
        // m = even[0] * (-1).  The factor is fetched from memory on the device: when ptxas can see that m is a
        // negation it rewrites the m * q_j products and stops fusing their lo/hi pairs into IMAD.WIDE.U32.X.
        const uint32_t m = even[0] * P::m0_opaque();
        const uint32_t c0 = even[0] != 0 ? 1u : 0u;
        const uint32_t ylo = ptx::sub_cc(0u, m - c0);
        const uint32_t yhi = ptx::subc(m, 0u);
        odd[0] = ptx::add_cc(odd[0], ylo);     // position 1
        odd[1] = ptx::addc_cc(odd[1], yhi);    // position 2
#pragma unroll
        for (int j = 2; j < N; j += 2) {       // q3, q5, q7 at positions 3, 5, 7
            odd[j] = ptx::madc_lo_cc(P::mod(j + 1), m, odd[j]);
            odd[j + 1] = ptx::madc_hi_cc(P::mod(j + 1), m, odd[j + 1]);
        }
        even[2] = ptx::mad_lo_cc(P::mod(2), m, even[2]);  // q2, q4, q6 at positions 2, 4, 6
        even[3] = ptx::madc_hi_cc(P::mod(2), m, even[3]);
#pragma unroll
        for (int j = 4; j < N; j += 2) {
            even[j] = ptx::madc_lo_cc(P::mod(j), m, even[j]);
            even[j + 1] = ptx::madc_hi_cc(P::mod(j), m, even[j + 1]);
        }
        odd[N - 1] = ptx::addc(odd[N - 1], 0);
        even[0] = 0;
        return;
    }
    uint32_t mi = even[0] * P::M0;
    cmad_mod<P, 1>(odd, mi);
    cmad_mod<P, 0>(even, mi);
    odd[N - 1] = ptx::addc(odd[N - 1], 0);
}

// r = (r >= p) ? r - p : r       for r < 2p
template <class P>
BPK_HD void final_sub(uint32_t* r) {
    constexpr int N = P::N;
    uint32_t t[N];
    t[0] = ptx::sub_cc(r[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < N; i++) t[i] = ptx::subc_cc(r[i], P::mod(i));
    uint32_t borrow = ptx::subc(0, 0);  // 0xffffffff if r < p
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = borrow ? r[i] : t[i];
}

}  // namespace detail

// Montgomery product a*b/R mod p, fully reduced; scalar.rs:562-586, fp.rs:565-609.  Two interleaved carry chains of
// IMAD.WIDE.U32.X (even / odd limb products), one Montgomery row per limb of b.  On sm_100 every form of IMAD.WIDE
// retires at 32 / clk / SM; a reduced-radix multiplier without carries and variants that moved chains to the ALU
// pipe were built, verified and measured slower (profiles/r1_imad_forms.md) -- this is the form that stayed.
// REDUCE = false ("lazy"): the final conditional subtraction is left out.  For operands below 2p the result is below
// (4 p^2 + R p) / R < 2p whenever R > 4p (Fp: R = 2^384, p < 2^381), so values may stay in [0, 2p) across a chain of
// products; sub_lazy / reduce_once below are the matching subtraction and the way back to the canonical residue.
// Fr (R = 2^256 < 4q): the FIRST operand must be canonical -- the running sum of the interleaved rows is bounded by
// (a + q) 2^32, which only fits the 9-limb accumulators for a < q -- and the second may lie in [0, 2q); the result is
// then below (2 q^2 + R q) / R < 1.91 q.
template <class P, bool REDUCE = true>
BPK_HD Fe<P> mul_cc(const Fe<P>& a, const Fe<P>& b) {
    constexpr int N = P::N;
    uint32_t even[N], odd[N];
#pragma unroll
    for (int i = 0; i < N; i += 2) {
        if (i == 0)
            detail::mad_n_redc<P, true>(even, odd, a.l, b.l[0]);
        else
            detail::mad_n_redc<P, false>(even, odd, a.l, b.l[i]);
        detail::mad_n_redc<P, false>(odd, even, a.l, b.l[i + 1]);
    }
    // merge: r[j] = even[j] + odd[j+1]
    Fe<P> r;
    r.l[0] = ptx::add_cc(even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(even[i], odd[i + 1]);
    r.l[N - 1] = ptx::addc(even[N - 1], 0);
    if (REDUCE) detail::final_sub<P>(r.l);
    return r;
}

namespace detail {
// one reduction-only row of the interleaved scheme of mad_n_redc (the a*b_i products left out): used by the
// dedicated squaring, where the full double-width square is formed first.  On exit even[0] == 0.
template <class P, bool FIRST>
BPK_HD void redc_row(uint32_t* even, uint32_t* odd) {
    constexpr int N = P::N;
    if (FIRST) {  // odd[] holds nothing yet
        const uint32_t mi = even[0] * P::M0;
#pragma unroll
        for (int j = 0; j < N; j += 2) mul_wide(odd[j], odd[j + 1], P::modk(j + 1), mi);
        cmad_mod<P, 0>(even, mi);
        odd[N - 1] = ptx::addc(odd[N - 1], 0);
        return;
    }
    even[0] = ptx::add_cc(even[0], odd[1]);
    const uint32_t mi = even[0] * P::M0;
    // odd[] <- (odd[] >> 64) + mi * (odd limbs of the modulus), consuming the carry of the addition above
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
        odd[j] = ptx::madc_lo_cc(P::modk(j + 1), mi, odd[j + 2]);
        odd[j + 1] = ptx::madc_hi_cc(P::modk(j + 1), mi, odd[j + 3]);
    }
    odd[N - 2] = ptx::madc_lo_cc(P::modk(N - 1), mi, 0);
    odd[N - 1] = ptx::madc_hi(P::modk(N - 1), mi, 0);
    cmad_mod<P, 0>(even, mi);
    odd[N - 1] = ptx::addc(odd[N - 1], 0);
}
}  // namespace detail

// Montgomery square a*a/R mod p, fully reduced.  The N(N-1)/2 off-diagonal limb products are formed once
// and doubled, so a square costs N(N+1)/2 + N^2 + N wide multiply-adds instead of 2 N^2 + N (Fp: 234
// instead of 300); the doubling and the merges run on the ALU pipe, which the multiplier leaves idle.
// Products at even limb positions accumulate in E, those at odd positions in O (O[k] sits at position
// k + 1), so that every lo/hi pair is one 64-bit aligned IMAD.WIDE.U32.X as in mul_cc.
template <class P, bool REDUCE = true>
BPK_HD Fe<P> sqr_cc(const Fe<P>& a) {
    constexpr int N = P::N;
    uint32_t E[2 * N], O[2 * N];
#pragma unroll
    for (int k = 0; k < 2 * N; k++) E[k] = O[k] = 0;
#pragma unroll
    for (int i = 0; i < N - 1; i++) {
        // even positions i + j: j = i + 2, i + 4, ...
        if (i + 2 < N) {
            int p = 2 * i + 2;
            E[p] = ptx::mad_lo_cc(a.l[i + 2], a.l[i], E[p]);
            E[p + 1] = ptx::madc_hi_cc(a.l[i + 2], a.l[i], E[p + 1]);
#pragma unroll
            for (int j = i + 4; j < N; j += 2) {
                p = i + j;
                E[p] = ptx::madc_lo_cc(a.l[j], a.l[i], E[p]);
                E[p + 1] = ptx::madc_hi_cc(a.l[j], a.l[i], E[p + 1]);
            }
            E[p + 2] = ptx::addc(E[p + 2], 0);
        }
        // odd positions i + j: j = i + 1, i + 3, ...  (position q lives in O[q - 1])
        {
            int p = 2 * i + 1;
            O[p - 1] = ptx::mad_lo_cc(a.l[i + 1], a.l[i], O[p - 1]);
            O[p] = ptx::madc_hi_cc(a.l[i + 1], a.l[i], O[p]);
#pragma unroll
            for (int j = i + 3; j < N; j += 2) {
                p = i + j;
                O[p - 1] = ptx::madc_lo_cc(a.l[j], a.l[i], O[p - 1]);
                O[p] = ptx::madc_hi_cc(a.l[j], a.l[i], O[p]);
            }
            O[p + 1] = ptx::addc(O[p + 1], 0);
        }
    }
    // S = E + (O << 32), then T = 2 S + sum a_i^2 2^(64 i)
    uint32_t T[2 * N];
    T[0] = E[0];
    T[1] = ptx::add_cc(E[1], O[0]);
#pragma unroll
    for (int k = 2; k < 2 * N - 1; k++) T[k] = ptx::addc_cc(E[k], O[k - 1]);
    T[2 * N - 1] = ptx::addc(E[2 * N - 1], O[2 * N - 2]);
#pragma unroll
    for (int k = 2 * N - 1; k >= 1; k--) T[k] = (T[k] << 1) | (T[k - 1] >> 31);
    T[0] <<= 1;
    {
        uint32_t lo, hi;
        detail::mul_wide(lo, hi, a.l[0], a.l[0]);
        T[0] = ptx::add_cc(T[0], lo);
        T[1] = ptx::addc_cc(T[1], hi);
#pragma unroll
        for (int i = 1; i < N; i++) {
            detail::mul_wide(lo, hi, a.l[i], a.l[i]);
            T[2 * i] = ptx::addc_cc(T[2 * i], lo);
            if (i < N - 1)
                T[2 * i + 1] = ptx::addc_cc(T[2 * i + 1], hi);
            else
                T[2 * i + 1] = ptx::addc(T[2 * i + 1], hi);
        }
    }
    // Montgomery reduction of the low half, N rows
    uint32_t even[N], odd[N];
#pragma unroll
    for (int j = 0; j < N; j++) even[j] = T[j];
    detail::redc_row<P, true>(even, odd);
    detail::redc_row<P, false>(odd, even);
#pragma unroll
    for (int i = 2; i < N; i += 2) {
        detail::redc_row<P, false>(even, odd);
        detail::redc_row<P, false>(odd, even);
    }
    Fe<P> r;
    r.l[0] = ptx::add_cc(even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(even[i], odd[i + 1]);
    r.l[N - 1] = ptx::addc(even[N - 1], 0);
    // + the high half
    r.l[0] = ptx::add_cc(r.l[0], T[N]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(r.l[i], T[N + i]);
    r.l[N - 1] = ptx::addc(r.l[N - 1], T[2 * N - 1]);
    if (REDUCE) detail::final_sub<P>(r.l);
    return r;
}

#if defined(__CUDACC__) && defined(BPK_FP_MUL_CALL)
// Out-of-line Fp product: a kernel whose hot loop inlines ten ~450-instruction multiplications (the MSM
// accumulate loop is ~75 KB of SASS) overflows the instruction caches; calling one shared body keeps the
// loop resident at the price of a register-passing call per product.
static __device__ __noinline__ Fe<FpParams> fp_mul_call(Fe<FpParams> a, Fe<FpParams> b) {
    return mul_cc<FpParams>(a, b);
}
static __device__ __noinline__ Fe<FpParams> fp_sqr_call(Fe<FpParams> a) { return sqr_cc<FpParams>(a); }
static __device__ __noinline__ Fe<FpParams> fp_mul_lazy_call(Fe<FpParams> a, Fe<FpParams> b) {
    return mul_cc<FpParams, false>(a, b);
}
static __device__ __noinline__ Fe<FpParams> fp_sqr_lazy_call(Fe<FpParams> a) { return sqr_cc<FpParams, false>(a); }
template <class P>
struct MulCall {
    static __device__ __forceinline__ Fe<P> run(const Fe<P>& a, const Fe<P>& b) { return mul_cc<P>(a, b); }
};
template <>
struct MulCall<FpParams> {
    static __device__ __forceinline__ Fe<FpParams> run(const Fe<FpParams>& a, const Fe<FpParams>& b) {
        return fp_mul_call(a, b);
    }
};
#endif

template <class P>
BPK_HD Fe<P> mul(const Fe<P>& a, const Fe<P>& b) {
#if defined(__CUDA_ARCH__) && defined(BPK_FP_MUL_CALL)
    return MulCall<P>::run(a, b);
#else
    return mul_cc<P>(a, b);
#endif
}

// lazy Fp product / square: operands and result in [0, 2p)
BPK_HD Fe<FpParams> mul_lazy(const Fe<FpParams>& a, const Fe<FpParams>& b) {
#if defined(__CUDA_ARCH__) && defined(BPK_FP_MUL_CALL)
    return fp_mul_lazy_call(a, b);
#else
    return mul_cc<FpParams, false>(a, b);
#endif
}
BPK_HD Fe<FpParams> sqr_lazy(const Fe<FpParams>& a) {
#if defined(__CUDA_ARCH__) && defined(BPK_FP_MUL_CALL)
    return fp_sqr_lazy_call(a);
#else
    return sqr_cc<FpParams, false>(a);
#endif
}

template <class P>
struct SqrImpl {  // Fr: the special-prime rows of mul_cc already beat a generic square
    BPK_HD static Fe<P> run(const Fe<P>& a) { return mul(a, a); }
};
#ifndef BPK_NO_DEDICATED_SQR
template <>
struct SqrImpl<FpParams> {
    BPK_HD static Fe<FpParams> run(const Fe<FpParams>& a) {
#if defined(__CUDA_ARCH__) && defined(BPK_FP_MUL_CALL)
        return fp_sqr_call(a);
#else
        return sqr_cc<FpParams>(a);
#endif
    }
};
#endif
template <class P>
BPK_HD Fe<P> sqr(const Fe<P>& a) {
    return SqrImpl<P>::run(a);
}

// scalar.rs:607-618, fp.rs:385-398
template <class P>
BPK_HD Fe<P> add(const Fe<P>& a, const Fe<P>& b) {
    constexpr int N = P::N;
    Fe<P> r;
    r.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
    r.l[N - 1] = ptx::addc(a.l[N - 1], b.l[N - 1]);  // no overflow: 2p < 2^(32N)
    detail::final_sub<P>(r.l);
    return r;
}

// scalar.rs:590-604, fp.rs:411-423
template <class P>
BPK_HD Fe<P> sub(const Fe<P>& a, const Fe<P>& b) {
    constexpr int N = P::N;
    Fe<P> r;
    r.l[0] = ptx::sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.l[i] = ptx::subc_cc(a.l[i], b.l[i]);
    uint32_t borrow = ptx::subc(0, 0);  // 0xffffffff on borrow
    r.l[0] = ptx::add_cc(r.l[0], P::mod(0) & borrow);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(r.l[i], P::mod(i) & borrow);
    r.l[N - 1] = ptx::addc(r.l[N - 1], P::mod(N - 1) & borrow);
    return r;
}

// a - b for a, b in [0, 2p): the result is brought back into [0, 2p) by adding 2p on borrow
template <class P>
BPK_HD Fe<P> sub_lazy(const Fe<P>& a, const Fe<P>& b) {
    constexpr int N = P::N;
    Fe<P> r;
    r.l[0] = ptx::sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.l[i] = ptx::subc_cc(a.l[i], b.l[i]);
    uint32_t borrow = ptx::subc(0, 0);  // 0xffffffff on borrow
    // 2p, limb by limb
    uint32_t carry = 0, two_p[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        two_p[i] = (P::mod(i) << 1) | carry;
        carry = P::mod(i) >> 31;
    }
    r.l[0] = ptx::add_cc(r.l[0], two_p[0] & borrow);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(r.l[i], two_p[i] & borrow);
    r.l[N - 1] = ptx::addc(r.l[N - 1], two_p[N - 1] & borrow);
    return r;
}

// a + b for a, b in [0, 2p), brought back into [0, 2p).  The sum may exceed the limb array (Fr: 4q > 2^256): the
// carry-out takes part in the decision, and the subtraction of 2p is then exact modulo 2^(32 N).
template <class P>
BPK_HD Fe<P> add_lazy(const Fe<P>& a, const Fe<P>& b) {
    constexpr int N = P::N;
    uint32_t carry = 0, two_p[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        two_p[i] = (P::mod(i) << 1) | carry;
        carry = P::mod(i) >> 31;
    }
    uint32_t s[N], d[N];
    s[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N; i++) s[i] = ptx::addc_cc(a.l[i], b.l[i]);
    const uint32_t c = ptx::addc(0, 0);                 // 1: the sum is 2^(32 N) + s
    d[0] = ptx::sub_cc(s[0], two_p[0]);
#pragma unroll
    for (int i = 1; i < N; i++) d[i] = ptx::subc_cc(s[i], two_p[i]);
    const uint32_t borrow = ptx::subc(0, 0);            // 0xffffffff: s < 2p
    const bool take = c != 0 || borrow == 0;
    Fe<P> r;
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = take ? d[i] : s[i];
    return r;
}

// [0, 2p) -> the canonical residue
template <class P>
BPK_HD Fe<P> reduce_once(const Fe<P>& a) {
    Fe<P> r = a;
    detail::final_sub<P>(r.l);
    return r;
}

template <class P>
BPK_HD Fe<P> neg(const Fe<P>& a) {
    return sub(Fe<P>::zero(), a);
}

template <class P>
BPK_HD Fe<P> dbl(const Fe<P>& a) {
    return add(a, a);
}

// Montgomery form -> canonical integer limbs (a * 1 / R).  scalar.rs:292-304
template <class P>
BPK_HD Fe<P> from_mont(const Fe<P>& a) {
    Fe<P> one_raw = Fe<P>::zero();
    one_raw.l[0] = 1;
    return mul(a, one_raw);
}

// canonical -> Montgomery (a * R^2 / R).  scalar.rs:282-284
template <class P>
BPK_HD Fe<P> to_mont(const Fe<P>& a) {
    return mul(a, Fe<P>::r2());
}

// ------------------------------------------------------------------------------------------
// Field inversion, 0 -> 0 (same VALUE as fp.rs:346-358 / scalar.rs:416-511, which exponentiate).
//
// Bernstein-Yang division steps ("safegcd", the delta = 1 form) on signed 30-bit limbs: 30 steps at a time are run
// on the low words of (f, g) alone and collected in a 2x2 transition matrix, which is then applied to the full
// (f, g) -- an exact division by 2^30 -- and to (d, e) modulo p.  No multiplications by field elements, no
// data-dependent branches inside a round: every lane of a warp runs the same instruction stream, which is what
// lets each MSM thread invert its own batch product (msm.cu, batched-affine additions) -- ~20 k mostly-ALU
// instructions against ~170 k wide multiply-adds for Fermat's a^(p-2), and the ALU pipe is the idle one there.
// Termination: floor((49 BITS + 57) / 17) steps suffice for BITS >= 46 (Bernstein-Yang 2019, Thm 11.2); the round
// loop also leaves as soon as g == 0, and the bound holds for ANY limbs, so malformed input cannot hang a thread.
// ------------------------------------------------------------------------------------------
namespace detail {
template <class P>
struct SafeGcd {
    static constexpr int N = P::N;
    static constexpr int NL = (P::BITS + 2 + 29) / 30;                       // values stay in (-2p, p)
    static constexpr int ROUNDS = ((49 * P::BITS + 57) / 17 + 29) / 30;      // Fp: 37 x 30 >= 1101, Fr: 25 x 30 >= 738
    static constexpr int32_t M30 = (int32_t)0x3fffffff;
    // bits [30 i, 30 i + 30) of an N-word integer
    BPK_HD static constexpr uint32_t limb30(const uint32_t* x, int i) {
        const int bit = 30 * i, w = bit >> 5, off = bit & 31;
        if (w >= N) return 0;
        uint32_t v = x[w] >> off;
        if (off > 2 && w + 1 < N) v |= x[w + 1] << (32 - off);
        return v & 0x3fffffffu;
    }
    BPK_HD static constexpr uint32_t mod30(int i) {
        const int bit = 30 * i, w = bit >> 5, off = bit & 31;
        if (w >= N) return 0;
        uint32_t v = P::mod(w) >> off;
        if (off > 2 && w + 1 < N) v |= P::mod(w + 1) << (32 - off);
        return v & 0x3fffffffu;
    }
    // p^-1 mod 2^30 by Newton iteration (p odd)
    BPK_HD static constexpr uint32_t mod_inv30() {
        uint32_t p0 = P::mod(0), x = p0;  // correct to 3 bits
        for (int i = 0; i < 5; i++) x *= 2u - p0 * x;
        return x & 0x3fffffffu;
    }
};
}  // namespace detail

template <class P>
BPK_HD Fe<P> inv(const Fe<P>& a) {
    typedef detail::SafeGcd<P> S;
    constexpr int N = P::N, NL = S::NL;
    constexpr int32_t M30 = S::M30;
    int32_t f[NL], g[NL], d[NL], e[NL];
#pragma unroll
    for (int i = 0; i < NL; i++) {
        f[i] = (int32_t)S::mod30(i);
        g[i] = (int32_t)S::limb30(a.l, i);
        d[i] = 0;
        e[i] = 0;
    }
    e[0] = 1;
    int32_t eta = -1;  // eta = -delta
#pragma unroll 1
    for (int round = 0; round < S::ROUNDS; round++) {
        int32_t nz = 0;
#pragma unroll
        for (int i = 0; i < NL; i++) nz |= g[i];
        if (nz == 0) break;
        // 30 division steps on the low words; 2^30 [f', g'] = [[u, v], [q, r]] [f, g]
        uint32_t u = 1, v = 0, q = 0, r = 1;
        uint32_t fl = (uint32_t)f[0], gl = (uint32_t)g[0];
#pragma unroll 2
        for (int i = 0; i < 30; i++) {
            uint32_t c1 = (uint32_t)(eta >> 31);  // delta > 0
            const uint32_t c2 = 0u - (gl & 1u);   // g odd
            const uint32_t x = (fl ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;  // -(f, u, v) if delta > 0
            gl += x & c2;
            q += y & c2;
            r += z & c2;
            c1 &= c2;                                   // swap case: delta > 0 and g odd
            eta = (int32_t)(((uint32_t)eta ^ c1) - (c1 + 1u));  // -eta - 1 there, eta - 1 otherwise
            fl += gl & c1;
            u += q & c1;
            v += r & c1;
            gl >>= 1;
            u <<= 1;
            v <<= 1;
        }
        // 32 x 32 -> 64 signed products (mul.wide.s32): both factors are sign-extended 32-bit values
        const int32_t iu = (int32_t)u, iv = (int32_t)v, iq = (int32_t)q, ir = (int32_t)r;
#define BPK_MW(a, b) ((int64_t)(a) * (int64_t)(b))
        {   // (d, e) <- [[u, v], [q, r]] (d, e) / 2^30 mod p, values kept in (-2p, p)
            const int32_t sd = d[NL - 1] >> 31, se = e[NL - 1] >> 31;
            int32_t md = ((int32_t)u & sd) + ((int32_t)v & se);
            int32_t me = ((int32_t)q & sd) + ((int32_t)r & se);
            int64_t cd = BPK_MW(iu, d[0]) + BPK_MW(iv, e[0]);
            int64_t ce = BPK_MW(iq, d[0]) + BPK_MW(ir, e[0]);
            md -= (int32_t)((S::mod_inv30() * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
            me -= (int32_t)((S::mod_inv30() * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
            cd += BPK_MW((int32_t)S::mod30(0), md);
            ce += BPK_MW((int32_t)S::mod30(0), me);
            cd >>= 30;
            ce >>= 30;
#pragma unroll
            for (int i = 1; i < NL; i++) {
                cd += BPK_MW(iu, d[i]) + BPK_MW(iv, e[i]) + BPK_MW((int32_t)S::mod30(i), md);
                ce += BPK_MW(iq, d[i]) + BPK_MW(ir, e[i]) + BPK_MW((int32_t)S::mod30(i), me);
                d[i - 1] = (int32_t)cd & M30;
                e[i - 1] = (int32_t)ce & M30;
                cd >>= 30;
                ce >>= 30;
            }
            d[NL - 1] = (int32_t)cd;
            e[NL - 1] = (int32_t)ce;
        }
        {   // (f, g) <- [[u, v], [q, r]] (f, g) / 2^30 (exact)
            int64_t cf = BPK_MW(iu, f[0]) + BPK_MW(iv, g[0]);
            int64_t cg = BPK_MW(iq, f[0]) + BPK_MW(ir, g[0]);
            cf >>= 30;
            cg >>= 30;
#pragma unroll
            for (int i = 1; i < NL; i++) {
                cf += BPK_MW(iu, f[i]) + BPK_MW(iv, g[i]);
                cg += BPK_MW(iq, f[i]) + BPK_MW(ir, g[i]);
                f[i - 1] = (int32_t)cf & M30;
                g[i - 1] = (int32_t)cg & M30;
                cf >>= 30;
                cg >>= 30;
            }
            f[NL - 1] = (int32_t)cf;
            g[NL - 1] = (int32_t)cg;
        }
#undef BPK_MW
    }
    // f = +-1 (gcd of p and a non-zero a); a^-1 = sign(f) d, brought from (-2p, p) to [0, p)
    {
        int32_t add = d[NL - 1] >> 31;
        const int32_t negate = f[NL - 1] >> 31;
#pragma unroll
        for (int i = 0; i < NL; i++) {
            d[i] += (int32_t)S::mod30(i) & add;
            d[i] = (d[i] ^ negate) - negate;
        }
#pragma unroll
        for (int i = 0; i < NL - 1; i++) {
            d[i + 1] += d[i] >> 30;
            d[i] &= M30;
        }
        add = d[NL - 1] >> 31;
#pragma unroll
        for (int i = 0; i < NL; i++) d[i] += (int32_t)S::mod30(i) & add;
#pragma unroll
        for (int i = 0; i < NL - 1; i++) {
            d[i + 1] += d[i] >> 30;
            d[i] &= M30;
        }
    }
    Fe<P> t;  // re-slice the 30-bit limbs into 32-bit words
#pragma unroll
    for (int j = 0; j < N; j++) {
        const int bit = 32 * j, k = bit / 30, off = bit % 30;
        uint32_t w = (uint32_t)d[k] >> off;
        if (k + 1 < NL) w |= (uint32_t)d[k + 1] << (30 - off);
        if (60 - off < 32 && k + 2 < NL) w |= (uint32_t)d[k + 2] << (60 - off);
        t.l[j] = w;
    }
    // t = (a as a plain integer)^-1 = (value R)^-1; value^-1 R = t R^2: two Montgomery multiplications by R^2
    return mul(mul(t, Fe<P>::r2()), Fe<P>::r2());
}

// a^e for a 64-bit exponent (Scalar::pow with [e,0,0,0], scalar.rs:381-392)
template <class P>
BPK_HD Fe<P> pow_u64(const Fe<P>& a, uint64_t e) {
    Fe<P> r = Fe<P>::one();
    int top = 63;
    while (top > 0 && !((e >> top) & 1ull)) top--;  // skip the leading zero bits (squarings of one)
#pragma unroll 1
    for (int b = top; b >= 0; b--) {
        r = sqr(r);
        if ((e >> b) & 1ull) r = mul(r, a);
    }
    return r;
}

typedef Fe<FrParams> fr_t;
typedef Fe<FpParams> fp_t;

}  // namespace bpk
