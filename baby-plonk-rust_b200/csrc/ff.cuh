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
// scalar.rs:83-88 (MODULUS), :164 (INV -> low 32 bits), :167-172 (R), :175-180 (R2)
struct FrParams {
    static constexpr int N = 8;
    static constexpr uint32_t M0 = 0xffffffffu;  // -q^-1 mod 2^32
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
};

// fp.rs:70-77 (MODULUS), :80 (INV), :83-90 (R), :93-100 (R2)
struct FpParams {
    static constexpr int N = 12;
    static constexpr uint32_t M0 = 0xfffcfffdu;  // -p^-1 mod 2^32
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
};

// ------------------------------------------------------------------------------------------
// the field element
// ------------------------------------------------------------------------------------------
template <class P>
struct alignas(16) Fe {
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

// acc[j], acc[j+1] = a[j] * bi for even j
template <int N>
BPK_HD void mul_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
        acc[j] = ptx::mul_lo(a[j], bi);
        acc[j + 1] = ptx::mul_hi(a[j], bi);
    }
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
    acc[0] = ptx::mad_lo_cc(P::mod(OFF), mi, acc[0]);
    acc[1] = ptx::madc_hi_cc(P::mod(OFF), mi, acc[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        acc[j] = ptx::madc_lo_cc(P::mod(j + OFF), mi, acc[j]);
        acc[j + 1] = ptx::madc_hi_cc(P::mod(j + OFF), mi, acc[j + 1]);
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

// Montgomery product a*b/R mod p, fully reduced.  (scalar.rs:562-586, fp.rs:565-609)
template <class P>
BPK_HD Fe<P> mul(const Fe<P>& a, const Fe<P>& b) {
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
    detail::final_sub<P>(r.l);
    return r;
}

template <class P>
BPK_HD Fe<P> sqr(const Fe<P>& a) {
    return mul(a, a);
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

// a^(p-2): field inversion by Fermat (fp.rs:346-358); 0 -> 0
template <class P>
BPK_HD Fe<P> inv(const Fe<P>& a) {
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

// a^e for a 64-bit exponent (Scalar::pow with [e,0,0,0], scalar.rs:381-392)
template <class P>
BPK_HD Fe<P> pow_u64(const Fe<P>& a, uint64_t e) {
    Fe<P> r = Fe<P>::one();
#pragma unroll 1
    for (int b = 63; b >= 0; b--) {
        r = sqr(r);
        if ((e >> b) & 1ull) r = mul(r, a);
    }
    return r;
}

typedef Fe<FrParams> fr_t;
typedef Fe<FpParams> fp_t;

}  // namespace bpk
