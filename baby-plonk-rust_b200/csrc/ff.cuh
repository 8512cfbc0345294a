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
// the very same templates are unit-tested on the CPU in the build container (tests/host_ff_test.cu).
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
    // reduced-radix form used inside mul(): 9 limbs of 29 bits (261 bits), operand a pre-shifted by 5
    static constexpr int RB = 29, NL = 9, PRESHIFT = 5;
    static constexpr uint32_t M0R = 0x1fffffffu;  // -q^-1 mod 2^29
    BPK_HD static constexpr uint32_t modr(int i) {
        constexpr uint32_t m[9] = {0x00000001u, 0x1ffffff8u, 0x1f96ffbfu, 0x1b4805ffu, 0x1d80553bu,
                                   0x0c0404d0u, 0x1520cce7u, 0x0a6533afu, 0x0073eda7u};
        return m[i];
    }
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
    // reduced-radix form used inside mul(): 14 limbs of 28 bits (392 bits), operand a pre-shifted by 8
    static constexpr int RB = 28, NL = 14, PRESHIFT = 8;
    static constexpr uint32_t M0R = 0x0ffcfffdu;  // -p^-1 mod 2^28
    BPK_HD static constexpr uint32_t modr(int i) {
        constexpr uint32_t m[14] = {0x0fffaaabu, 0x0fefffffu, 0x03ffffb9u, 0x0fffeb15u, 0x06241eabu,
                                    0x0a0f6b0fu, 0x0f6730d2u, 0x0f38512bu, 0x04774b84u, 0x04bacd76u,
                                    0x0ba7b643u, 0x0e69a4b1u, 0x01ea397fu, 0x0001a011u};
        return m[i];
    }
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

// "split" forms of madc_n_rshift / cmad_mod: the products are formed without addend on the heavy
// FMA pipe and folded in by an add-with-carry chain on the ALU pipe (IADD3.X), so that the two
// pipes share the work of a row instead of the FMA pipe doing all of it at half rate.
template <int N>
BPK_HD void madc_n_rshift_split(uint32_t* odd, const uint32_t* a, uint32_t bi) {
    uint32_t pl[N / 2], ph[N / 2];
#pragma unroll
    for (int j = 0; j < N; j += 2) mul_wide(pl[j / 2], ph[j / 2], a[j], bi);
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
        odd[j] = ptx::addc_cc(pl[j / 2], odd[j + 2]);
        odd[j + 1] = ptx::addc_cc(ph[j / 2], odd[j + 3]);
    }
    odd[N - 2] = ptx::addc_cc(pl[N / 2 - 1], 0);
    odd[N - 1] = ptx::addc(ph[N / 2 - 1], 0);
}
template <class P, int OFF>
BPK_HD void cmad_mod_split(uint32_t* acc, uint32_t mi) {
    constexpr int N = P::N;
    uint32_t pl[N / 2], ph[N / 2];
#pragma unroll
    for (int j = 0; j < N; j += 2) mul_wide(pl[j / 2], ph[j / 2], P::mod(j + OFF), mi);
    acc[0] = ptx::add_cc(acc[0], pl[0]);
    acc[1] = ptx::addc_cc(acc[1], ph[0]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        acc[j] = ptx::addc_cc(acc[j], pl[j / 2]);
        acc[j + 1] = ptx::addc_cc(acc[j + 1], ph[j / 2]);
    }
}

// one row of the interleaved multiply + Montgomery reduction.
// value = sum even[j] 2^(32j) + sum odd[j] 2^(32(j+1)); on exit even[0] == 0
// SPLIT: bit 0 -> odd a*b chain on the ALU pipe, bit 1 -> odd m*p chain, bit 2 -> even m*p chain
template <class P, bool FIRST, int SPLIT = 0>
BPK_HD void mad_n_redc(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi) {
    constexpr int N = P::N;
    if (FIRST) {
        mul_n<N>(odd, a + 1, bi);
        mul_n<N>(even, a, bi);
    } else {
        even[0] = ptx::add_cc(even[0], odd[1]);
        if (SPLIT & 1)
            madc_n_rshift_split<N>(odd, a + 1, bi);
        else
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
    if (SPLIT & 2)
        cmad_mod_split<P, 1>(odd, mi);
    else
        cmad_mod<P, 1>(odd, mi);
    if (SPLIT & 4)
        cmad_mod_split<P, 0>(even, mi);
    else
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

// Montgomery product a*b/R mod p, fully reduced -- carry-chained 32-bit-limb version (v1).
// Kept as the cross-check of mul(); every IMAD.WIDE.U32.X it issues occupies the heavy FMA pipe for
// two slots on sm_100 (measured, profiles/r1_v1_msm_accumulate_ncu.md), which is why mul() below
// does not use carry flags on the multiplier pipe.
template <class P, int SPLIT = 0>
BPK_HD Fe<P> mul_cc(const Fe<P>& a, const Fe<P>& b) {
    constexpr int N = P::N;
    uint32_t even[N], odd[N];
#pragma unroll
    for (int i = 0; i < N; i += 2) {
        if (i == 0)
            detail::mad_n_redc<P, true, SPLIT>(even, odd, a.l, b.l[0]);
        else
            detail::mad_n_redc<P, false, SPLIT>(even, odd, a.l, b.l[i]);
        detail::mad_n_redc<P, false, SPLIT>(odd, even, a.l, b.l[i + 1]);
    }
    // merge: r[j] = even[j] + odd[j+1]
    Fe<P> r;
    r.l[0] = ptx::add_cc(even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(even[i], odd[i + 1]);
    r.l[N - 1] = ptx::addc(even[N - 1], 0);
    detail::final_sub<P>(r.l);
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
template <class P>
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
    detail::final_sub<P>(r.l);
    return r;
}

namespace detail {
// (hi:lo) += a * b, 32 x 32 -> 64 multiply-add on a 64-bit accumulator held in two registers.
// The mad.lo.cc / madc.hi pair is what ptxas fuses into ONE plain IMAD.WIDE.U32 Rd, Ra, Rb, Rd; a
// `mad.wide.u32` with a 64-bit addend is instead split by ptxas into IMAD.WIDE(.., RZ) + IADD3 + IADD3.X.
BPK_HD void mad_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
#else
    uint64_t t = (((uint64_t)hi << 32) | lo) + (uint64_t)a * b;
    lo = (uint32_t)t;
    hi = (uint32_t)(t >> 32);
#endif
}
// limb k (RB bits) of the integer (x << SH), x given as N 32-bit words
template <class P, int SH>
BPK_HD uint32_t rr_limb(const uint32_t* x, int k) {
    constexpr int N = P::N, RB = P::RB;
    constexpr uint32_t MASK = (1u << RB) - 1;
    const int s = RB * k - SH;
    if (s < 0) return (x[0] << (-s)) & MASK;
    const int w = s >> 5, off = s & 31;
    if (w >= N) return 0;
    uint32_t lo = x[w];
    uint32_t hi = (w + 1 < N) ? x[w + 1] : 0u;
    uint32_t v = off ? ((lo >> off) | (hi << (32 - off))) : lo;
    return v & MASK;
}
}  // namespace detail

// Montgomery product a*b/R mod p (R = 2^(32 N)), fully reduced; scalar.rs:562-586, fp.rs:565-609.
//
// Reduced-radix formulation for the sm_100 multiplier pipe: the operands are re-sliced into NL limbs
// of RB bits (Fp: 14 x 28, Fr: 9 x 29) and every partial product a_j b_i and m_i p_j is accumulated
// into a 64-bit column with a plain IMAD.WIDE.U32 -- 2 NL products of < 2^(2 RB) never overflow 64
// bits, so the multiplier pipe sees no carry flags at all (a carry-in IMAD.WIDE.U32.X costs two pipe
// slots).  One row retires RB bits: m_i = t0 * (-p^-1) mod 2^RB, t += m_i p, t >>= RB.  NL rows divide by
// 2^(RB NL); operand a enters pre-shifted by RB NL - 32 N bits, so the result is exactly a b / 2^(32 N):
// the Montgomery domain is the reference's.  Column carries, re-slicing and the final conditional
// subtraction run on the ALU pipe, which is otherwise idle.
template <class P>
BPK_HD Fe<P> mul_rr(const Fe<P>& a, const Fe<P>& b) {
    constexpr int N = P::N, NL = P::NL, RB = P::RB;
    constexpr uint32_t MASK = (1u << RB) - 1;
    uint32_t A[NL];
#pragma unroll
    for (int k = 0; k < NL; k++) A[k] = detail::rr_limb<P, P::PRESHIFT>(a.l, k);
    uint32_t tl[NL + 1], th[NL + 1];  // 64-bit columns as (lo, hi) register pairs
#pragma unroll
    for (int k = 0; k <= NL; k++) tl[k] = th[k] = 0;
#pragma unroll
    for (int i = 0; i < NL; i++) {
        const uint32_t bi = detail::rr_limb<P, 0>(b.l, i);
#pragma unroll
        for (int j = 0; j < NL; j++) detail::mad_wide(tl[j], th[j], A[j], bi);
        const uint32_t m = (tl[0] * P::M0R) & MASK;
#pragma unroll
        for (int j = 0; j < NL; j++) detail::mad_wide(tl[j], th[j], m, P::modr(j));
        // column 0 is now a multiple of 2^RB: retire it, carry into column 1, slide the window
        const uint32_t cl = (tl[0] >> RB) | (th[0] << (32 - RB));
        const uint32_t ch = th[0] >> RB;
        tl[0] = ptx::add_cc(tl[1], cl);
        th[0] = ptx::addc(th[1], ch);
#pragma unroll
        for (int j = 1; j < NL; j++) {
            tl[j] = tl[j + 1];
            th[j] = th[j + 1];
        }
        tl[NL] = th[NL] = 0;
    }
    // carry-normalise the NL columns to RB-bit limbs
    uint32_t L[NL + 1];
    uint32_t cl = 0, ch = 0;
#pragma unroll
    for (int j = 0; j < NL; j++) {
        uint32_t vl = ptx::add_cc(tl[j], cl);
        uint32_t vh = ptx::addc(th[j], ch);
        L[j] = vl & MASK;
        cl = (vl >> RB) | (vh << (32 - RB));
        ch = vh >> RB;
    }
    L[NL] = cl;  // 0 for reduced inputs (result < 2p < 2^(RB NL))
    // re-slice to 32-bit words
    Fe<P> r;
#pragma unroll
    for (int j = 0; j < N; j++) {
        const int bit = 32 * j;
        const int k = bit / RB, off = bit % RB;
        uint32_t v = L[k] >> off;
        if (k + 1 <= NL) v |= L[k + 1] << (RB - off);
        if (2 * RB - off < 32 && k + 2 <= NL) v |= L[k + 2] << (2 * RB - off);
        r.l[j] = v;
    }
    detail::final_sub<P>(r.l);
    return r;
}

// The product the kernels use.  BPK_MUL_IMPL: 0 = carry chains only (v1), 1 = reduced radix,
// 2.. = carry chains with SPLIT = BPK_MUL_IMPL - 2 (some chains moved to the ALU pipe).
#ifndef BPK_MUL_IMPL
#define BPK_MUL_IMPL 0
#endif
#if defined(__CUDACC__) && defined(BPK_FP_MUL_CALL)
// Out-of-line Fp product: a kernel whose hot loop inlines ten ~450-instruction multiplications (the MSM
// accumulate loop is ~75 KB of SASS) overflows the instruction caches; calling one shared body keeps the
// loop resident at the price of a register-passing call per product.
template <class P, int SPLIT>
BPK_HD Fe<P> mul_cc(const Fe<P>& a, const Fe<P>& b);
static __device__ __noinline__ Fe<FpParams> fp_mul_call(Fe<FpParams> a, Fe<FpParams> b) {
    return mul_cc<FpParams, 0>(a, b);
}
static __device__ __noinline__ Fe<FpParams> fp_sqr_call(Fe<FpParams> a) { return sqr_cc<FpParams>(a); }
template <class P>
struct MulCall {
    static __device__ __forceinline__ Fe<P> run(const Fe<P>& a, const Fe<P>& b) { return mul_cc<P, 0>(a, b); }
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
#endif
#if BPK_MUL_IMPL == 1
    return mul_rr<P>(a, b);
#elif BPK_MUL_IMPL >= 2
    return mul_cc<P, BPK_MUL_IMPL - 2>(a, b);
#else
    return mul_cc<P, 0>(a, b);
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

// a^(p-2): field inversion by Fermat (fp.rs:346-358); 0 -> 0.  Kept as the cross-check of inv().
template <class P>
BPK_HD Fe<P> inv_fermat(const Fe<P>& a) {
    constexpr int N = P::N;
    Fe<P> r = Fe<P>::one();
    // exponent p - 2, limb by limb with borrow (Fr's low limb is 1)
    uint32_t ex[N];
    uint32_t borrow = 2;
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint32_t m = P::mod(i);
        ex[i] = m - borrow;
        borrow = (m < borrow) ? 1u : 0u;
    }
#pragma unroll 1
    for (int i = N - 1; i >= 0; i--) {
        uint32_t e = ex[i];
#pragma unroll 1
        for (int b = 31; b >= 0; b--) {
            r = sqr(r);
            if ((e >> b) & 1u) r = mul(r, a);
        }
    }
    return r;
}

// Field inversion, 0 -> 0 (same value as fp.rs:346-358 / scalar.rs:416-511, which exponentiate).
// Binary extended Euclid on the raw limbs: ~2 * bits iterations of shifts and add/sub chains, no
// multiplications -- about 6x shorter than the 570 dependent Montgomery multiplications of Fermat's
// method when a single thread has to normalise an MSM result (msm_finalize_kernel).
template <class P>
BPK_HD Fe<P> inv(const Fe<P>& a) {
    constexpr int N = P::N;
    if (a.is_zero()) return a;
    uint32_t u[N], v[N];
    Fe<P> x1 = Fe<P>::zero(), x2 = Fe<P>::zero();
    x1.l[0] = 1;
#pragma unroll
    for (int i = 0; i < N; i++) {
        u[i] = a.l[i];
        v[i] = P::mod(i);
    }
    // x <- x / 2 mod p
    auto halve = [](Fe<P>& x) {
        uint32_t c = 0;
        if (x.l[0] & 1u) {
            x.l[0] = ptx::add_cc(x.l[0], P::mod(0));
#pragma unroll
            for (int i = 1; i < N; i++) x.l[i] = ptx::addc_cc(x.l[i], P::mod(i));
            c = ptx::addc(0u, 0u);
        }
#pragma unroll
        for (int i = 0; i < N - 1; i++) x.l[i] = (x.l[i] >> 1) | (x.l[i + 1] << 31);
        x.l[N - 1] = (x.l[N - 1] >> 1) | (c << 31);
    };
    auto shr1 = [](uint32_t* w) {
#pragma unroll
        for (int i = 0; i < N - 1; i++) w[i] = (w[i] >> 1) | (w[i + 1] << 31);
        w[N - 1] >>= 1;
    };
    auto is_one = [](const uint32_t* w) {
        uint32_t acc = w[0] ^ 1u;
#pragma unroll
        for (int i = 1; i < N; i++) acc |= w[i];
        return acc == 0;
    };
#pragma unroll 1
    while (!is_one(u) && !is_one(v)) {
#pragma unroll 1
        while ((u[0] & 1u) == 0) {
            shr1(u);
            halve(x1);
        }
#pragma unroll 1
        while ((v[0] & 1u) == 0) {
            shr1(v);
            halve(x2);
        }
        uint32_t t[N];
        t[0] = ptx::sub_cc(u[0], v[0]);
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = ptx::subc_cc(u[i], v[i]);
        uint32_t borrow = ptx::subc(0u, 0u);
        if (borrow == 0) {  // u >= v
#pragma unroll
            for (int i = 0; i < N; i++) u[i] = t[i];
            x1 = sub(x1, x2);
        } else {
            v[0] = ptx::sub_cc(v[0], u[0]);
#pragma unroll
            for (int i = 1; i < N - 1; i++) v[i] = ptx::subc_cc(v[i], u[i]);
            v[N - 1] = ptx::subc(v[N - 1], u[N - 1]);
            x2 = sub(x2, x1);
        }
    }
    Fe<P> t = is_one(u) ? x1 : x2;  // (a as a plain integer)^-1 = (value R)^-1
    // value^-1 R = t R^2: two Montgomery multiplications by R^2
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
