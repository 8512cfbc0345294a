// Fr NTT / iNTT / coset NTT for sm_100a.
//
// Contract = the reference's naive DFT (src/utils.rs:63-81 ntt_381, :106-129 i_ntt_381):
//   out[i] = sum_j in[j] * w^(i j),  w = ROOT_OF_UNITY^(2^32 / n)  (inverse: w^-1 and a final n^-1),
// natural order in and out.  Field values are canonical Montgomery residues, so any correct
// algorithm is bit-identical to the O(n^2) reference.
//
// Algorithm: multi-pass Stockham autosort.  A pass with radix R = 2^logR and Ns = product of the
// previous radices maps, for every column j < n/R,
//     v[r]   = in[j + r n/R] * w_{Ns R}^{(j mod Ns) r}          (inter-pass twiddle)
//     V      = DFT_R(v)                                          (shared memory, radix-2 DIT)
//     out[(j / Ns) Ns R + (j mod Ns) + r Ns] = V[r]
// A CTA owns a tile of C consecutive columns (R x C elements staged in shared memory, split into
// two 16-byte planes so 128-bit shared accesses are conflict-free); global reads are C*32-byte
// runs, global writes are C*32-byte runs (Ns >= C) or R*32-byte runs (first pass).  Inter-pass
// twiddles come from a two-level table of the primitive 2^28-th root (2 x 8192.. entries, L2/L1
// resident); sub-FFT twiddles w_R^i are staged in shared memory per CTA.
// Algorithmic traffic: 64 n bytes per transform; this schedule moves 64 n bytes per pass.
// (oracle/prototypes/stockham_proto.py is the CPU prototype of exactly this index logic.)
#include "internal.cuh"

namespace bpk {

struct NttPassParams {
    const fr_t* in;
    fr_t* out;
    const fr_t* tw_lo;
    const fr_t* tw_hi;
    const fr_t* cs_lo;  // coset tables (or null)
    const fr_t* cs_hi;
    const fr_t* tw_direct;  // optional: w_{Ns R}^e for e < Ns R, one lookup instead of lookup-lookup-multiply
    uint32_t logn, logR, logC, logNs;
    int coset_in;   // multiply input element i by cs(i)
    int scale_out;  // 0: none, 1: multiply outputs by `scale`, 2: multiply output i by cs(i)
    int lazy_out;   // outputs may stay in [0, 2q): set for a pass whose successor is the register-blocked kernel
    fr_t scale;
};

__device__ __forceinline__ fr_t ld_fr(const fr_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    fr_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fr(fr_t* p, const fr_t& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
__device__ __forceinline__ fr_t ld_sm(const uint4* lo, const uint4* hi, uint32_t i) {
    uint4 a = lo[i], b = hi[i];
    fr_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_sm(uint4* lo, uint4* hi, uint32_t i, const fr_t& v) {
    lo[i] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    hi[i] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// element E of a two-level power table: lo[E & mask] * hi[E >> TW_LO_BITS]
__device__ __forceinline__ fr_t table_pow(const fr_t* lo, const fr_t* hi, uint32_t E) {
    fr_t a = ld_fr(lo + (E & ((1u << TW_LO_BITS) - 1)));
    uint32_t h = E >> TW_LO_BITS;
    if (h == 0) return a;  // warp-divergent only at table boundaries; saves a multiply for small E
    return mul(a, ld_fr(hi + h));
}

__global__ void __launch_bounds__(1024) ntt_pass_kernel(NttPassParams p) {
    extern __shared__ uint4 smem[];
    const uint32_t logR = p.logR, logC = p.logC, logNs = p.logNs;
    const uint32_t R = 1u << logR, C = 1u << logC, RC = R << logC;
    uint4* s_lo = smem;
    uint4* s_hi = smem + RC;
    uint4* t_lo = smem + 2 * RC;
    uint4* t_hi = t_lo + (R >> 1);
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const size_t n = (size_t)1 << p.logn;
    const fr_t* in = p.in + (size_t)blockIdx.y * n;
    fr_t* out = p.out + (size_t)blockIdx.y * n;
    const uint32_t j0 = blockIdx.x << logC;

    // sub-FFT twiddles w_R^i, i < R/2
    for (uint32_t i = tid; i < (R >> 1); i += nt) {
        fr_t w = table_pow(p.tw_lo, p.tw_hi, i << (NTT_MAX_LOG - logR));
        st_sm(t_lo, t_hi, i, w);
    }
    // load tile, apply coset shift and inter-pass twiddle, place rows bit-reversed
    const uint32_t stride_in = (uint32_t)(n >> logR);
    const uint32_t tw_shift = NTT_MAX_LOG - (logNs + logR);
    const uint32_t ns_mask = (1u << logNs) - 1;
    for (uint32_t idx = tid; idx < RC; idx += nt) {
        uint32_t c = idx & (C - 1), r = idx >> logC;
        uint32_t j = j0 + c;
        uint32_t g = j + r * stride_in;
        fr_t v = ld_fr(in + g);
        if (p.coset_in) v = mul(v, table_pow(p.cs_lo, p.cs_hi, g));
        if (logNs) {
            uint32_t k = j & ns_mask;
            if (p.tw_direct)
                v = mul(v, ld_fr(p.tw_direct + k * r));
            else
                v = mul(v, table_pow(p.tw_lo, p.tw_hi, (k * r) << tw_shift));
        }
        uint32_t rr = __brev(r) >> (32 - logR);
        if (logR == 0) rr = 0;
        st_sm(s_lo, s_hi, (rr << logC) + c, v);
    }
    __syncthreads();
    // radix-2 DIT stages, natural-order output
    const uint32_t nbf = RC >> 1;
    for (uint32_t s = 1; s <= logR; s++) {
        const uint32_t half = 1u << (s - 1);
        for (uint32_t idx = tid; idx < nbf; idx += nt) {
            uint32_t c = idx & (C - 1), b = idx >> logC;
            uint32_t lowb = b & (half - 1);
            uint32_t i = ((b >> (s - 1)) << s) | lowb;
            uint32_t i0 = (i << logC) + c, i1 = ((i + half) << logC) + c;
            fr_t u = ld_sm(s_lo, s_hi, i0);
            fr_t t = ld_sm(s_lo, s_hi, i1);
            if (s > 1) t = mul(t, ld_sm(t_lo, t_hi, lowb << (logR - s)));  // stage 1 twiddles are all 1
            st_sm(s_lo, s_hi, i0, add(u, t));
            st_sm(s_lo, s_hi, i1, sub(u, t));
        }
        __syncthreads();
    }
    // store
    for (uint32_t idx = tid; idx < RC; idx += nt) {
        uint32_t c, r;
        size_t o;
        if (logNs == 0) {  // first pass: each column writes R consecutive outputs
            r = idx & (R - 1);
            c = idx >> logR;
            o = ((size_t)(j0 + c) << logR) + r;
        } else {
            c = idx & (C - 1);
            r = idx >> logC;
            uint32_t j = j0 + c;
            o = ((size_t)(j >> logNs) << (logNs + logR)) + (j & ns_mask) + ((size_t)r << logNs);
        }
        fr_t v = ld_sm(s_lo, s_hi, (r << logC) + c);
        if (p.scale_out == 1) v = mul(v, p.scale);
        else if (p.scale_out == 2) v = mul(v, table_pow(p.cs_lo, p.cs_hi, (uint32_t)o));
        st_fr(out + o, v);
    }
}


// ---------------------------------------------------------------------------------------------
// Register-blocked variant (default for R >= 8): every thread keeps 8 rows of one column in registers and
// runs up to three radix-2 DIT stages on them between two visits of shared memory, so a 2^8-point sub-FFT
// costs 3 exchanges (and barriers) instead of 8.  The first step is fused with the global load (rows are
// fetched in bit-reversed order straight into registers) and knows its twiddles at compile time: of its 12
// butterflies only 5 need a multiplication (exponents 0 are skipped).  The last step stores straight from
// registers, except in the first pass where the coalesced store wants R consecutive outputs per column.
// The Fr product is one shared out-of-line body: 12 inlined products per step would not fit the
// instruction cache (same finding as in msm.cu).
// ---------------------------------------------------------------------------------------------
static __device__ __noinline__ fr_t fr_mul_nl(fr_t a, fr_t b) { return mul(a, b); }
// Lazy butterflies (register-blocked kernel): values in flight live in [0, 2q).  The product of a canonical twiddle
// (first operand, see mul_cc) with such a value stays below 1.91 q without its conditional subtraction; sums and
// differences are folded back below 2q; the transform's outputs are reduced once, in its last pass.
#ifndef BPK_NTT_LAZY
#define BPK_NTT_LAZY 1
#endif
static __device__ __noinline__ fr_t fr_mul_lazy_nl(fr_t w, fr_t v) { return mul_cc<FrParams, false>(w, v); }
__device__ __forceinline__ fr_t ntt_twiddle_mul(const fr_t& v, const fr_t& w) {
    return BPK_NTT_LAZY ? fr_mul_lazy_nl(w, v) : fr_mul_nl(v, w);
}
__device__ __forceinline__ fr_t ntt_add(const fr_t& a, const fr_t& b) { return BPK_NTT_LAZY ? add_lazy(a, b) : add(a, b); }
__device__ __forceinline__ fr_t ntt_sub(const fr_t& a, const fr_t& b) { return BPK_NTT_LAZY ? sub_lazy(a, b) : sub(a, b); }
// the value a pass hands on: canonical unless the next pass accepts [0, 2q)
__device__ __forceinline__ fr_t ntt_finish(const fr_t& v, const NttPassParams& p) {
    if (p.scale_out == 0) return (BPK_NTT_LAZY && !p.lazy_out) ? reduce_once(v) : v;
    return v;  // scaled outputs go through a reducing product below
}

template <int T, bool FIRST>
__device__ __forceinline__ void dit_stage8(fr_t (&x)[8], uint32_t sigma, uint32_t base, uint32_t b0, uint32_t logR,
                                           const uint4* t_lo, const uint4* t_hi) {
#pragma unroll
    for (int a = 0; a < 8; a++) {
        if (a & (1 << T)) continue;
        const int a1 = a | (1 << T);
        fr_t tt = x[a1];
        if (FIRST) {  // stages 1..3 on rows base + a with base a multiple of 8: the exponent depends on a only
            const uint32_t lowb = a & ((1 << T) - 1);
            if (lowb != 0) tt = ntt_twiddle_mul(tt, ld_sm(t_lo, t_hi, lowb << (logR - sigma)));
        } else {
            const uint32_t row = base + ((uint32_t)a << b0);
            const uint32_t lowb = row & ((1u << (sigma - 1)) - 1);
            tt = ntt_twiddle_mul(tt, ld_sm(t_lo, t_hi, lowb << (logR - sigma)));
        }
        const fr_t u = x[a];
        x[a] = ntt_add(u, tt);
        x[a1] = ntt_sub(u, tt);
    }
}

// (register caps for 5 or 6 CTAs per SM were measured: 1.05 / 1.02 ms at 2^22 against 0.875 ms -- spills cost more than
// the extra warps bring; even spelling the bound as (256, 1) changes ptxas' allocation to 140 registers and 0.925 ms)
__global__ void __launch_bounds__(256) ntt_pass_r8_kernel(NttPassParams p) {
    extern __shared__ uint4 smem[];
    const uint32_t logR = p.logR, logC = p.logC, logNs = p.logNs;
    const uint32_t R = 1u << logR, C = 1u << logC, RC = R << logC;
    uint4* s_lo = smem;
    uint4* s_hi = smem + RC;
    uint4* t_lo = smem + 2 * RC;
    uint4* t_hi = t_lo + (R >> 1);
    const uint32_t tid = threadIdx.x, nt = blockDim.x;  // nt == RC / 8
    const size_t n = (size_t)1 << p.logn;
    const fr_t* in = p.in + (size_t)blockIdx.y * n;
    fr_t* out = p.out + (size_t)blockIdx.y * n;
    const uint32_t j0 = blockIdx.x << logC;
    const uint32_t c = tid & (C - 1), g = tid >> logC;  // column, group of 8 rows
    const uint32_t j = j0 + c;
    const uint32_t ns_mask = (1u << logNs) - 1;

    for (uint32_t i = tid; i < (R >> 1); i += nt) {
        fr_t w = table_pow(p.tw_lo, p.tw_hi, i << (NTT_MAX_LOG - logR));
        st_sm(t_lo, t_hi, i, w);
    }
    __syncthreads();

    // step 1: load rows brev(8 g + a), inter-pass twiddle, stages 1..3
    fr_t x[8];
    uint32_t base = g << 3, b0 = 0;
    {
        const uint32_t stride_in = (uint32_t)(n >> logR);
        const uint32_t tw_shift = NTT_MAX_LOG - (logNs + logR);
        const uint32_t k = j & ns_mask;
#pragma unroll
        for (int a = 0; a < 8; a++) {
            const uint32_t r = __brev(base + a) >> (32 - logR);
            const uint32_t gi = j + r * stride_in;
            fr_t v = ld_fr(in + gi);
            if (p.coset_in) v = ntt_twiddle_mul(v, table_pow(p.cs_lo, p.cs_hi, gi));
            if (logNs) {
                if (p.tw_direct)
                    v = ntt_twiddle_mul(v, ld_fr(p.tw_direct + k * r));
                else
                    v = ntt_twiddle_mul(v, table_pow(p.tw_lo, p.tw_hi, (k * r) << tw_shift));
            }
            x[a] = v;
        }
        dit_stage8<0, true>(x, 1, base, 0, logR, t_lo, t_hi);
        dit_stage8<1, true>(x, 2, base, 0, logR, t_lo, t_hi);
        dit_stage8<2, true>(x, 3, base, 0, logR, t_lo, t_hi);
    }
    uint32_t done = 3;
    while (done < logR) {
#pragma unroll
        for (int a = 0; a < 8; a++) st_sm(s_lo, s_hi, ((base + ((uint32_t)a << b0)) << logC) + c, x[a]);
        __syncthreads();
        const uint32_t k = logR - done < 3 ? logR - done : 3;
        const uint32_t s = done + 1;
        b0 = s - 1 < logR - 3 ? s - 1 : logR - 3;
        base = ((g >> b0) << (b0 + 3)) | (g & ((1u << b0) - 1));
#pragma unroll
        for (int a = 0; a < 8; a++) x[a] = ld_sm(s_lo, s_hi, ((base + ((uint32_t)a << b0)) << logC) + c);
        if (k == 3) {
            dit_stage8<0, false>(x, s, base, b0, logR, t_lo, t_hi);
            dit_stage8<1, false>(x, s + 1, base, b0, logR, t_lo, t_hi);
            dit_stage8<2, false>(x, s + 2, base, b0, logR, t_lo, t_hi);
        } else if (k == 2) {
            dit_stage8<1, false>(x, s, base, b0, logR, t_lo, t_hi);
            dit_stage8<2, false>(x, s + 1, base, b0, logR, t_lo, t_hi);
        } else {
            dit_stage8<2, false>(x, s, base, b0, logR, t_lo, t_hi);
        }
        done += k;
    }
    if (logNs != 0) {  // rows base + (a << b0) of column j, straight from registers
        const size_t o0 = ((size_t)(j >> logNs) << (logNs + logR)) + (j & ns_mask);
#pragma unroll
        for (int a = 0; a < 8; a++) {
            const uint32_t r = base + ((uint32_t)a << b0);
            const size_t o = o0 + ((size_t)r << logNs);
            fr_t v = ntt_finish(x[a], p);
            if (p.scale_out == 1) v = fr_mul_nl(p.scale, v);   // (canonical factor first: v may lie in [0, 2q))
            else if (p.scale_out == 2) v = fr_mul_nl(table_pow(p.cs_lo, p.cs_hi, (uint32_t)o), v);
            st_fr(out + o, v);
        }
        return;
    }
    // first pass: each column writes R consecutive outputs -> go through shared memory for coalescing
    __syncthreads();  // (all reads of the last exchange are done before rows are overwritten)
#pragma unroll
    for (int a = 0; a < 8; a++) st_sm(s_lo, s_hi, ((base + ((uint32_t)a << b0)) << logC) + c, x[a]);
    __syncthreads();
    for (uint32_t idx = tid; idx < RC; idx += nt) {
        const uint32_t r = idx & (R - 1), cc = idx >> logR;
        const size_t o = ((size_t)(j0 + cc) << logR) + r;
        fr_t v = ntt_finish(ld_sm(s_lo, s_hi, (r << logC) + cc), p);
        if (p.scale_out == 1) v = fr_mul_nl(p.scale, v);
        else if (p.scale_out == 2) v = fr_mul_nl(table_pow(p.cs_lo, p.cs_hi, (uint32_t)o), v);
        st_fr(out + o, v);
    }
}

// ---------------------------------------------------------------------------------------------
// Experimental variant (ntt.kernel = 3): 4 rows per thread, two stages per exchange, and the two butterflies of a
// stage share ONE out-of-line call that forms both products -- four interleaved carry chains per warp instead
// of two, at a register budget that still allows 16+ warps per SM.
// ---------------------------------------------------------------------------------------------
struct fr_pair_t {
    fr_t a, b;
};
static __device__ __noinline__ fr_pair_t fr_mul2_nl(fr_t a0, fr_t b0, fr_t a1, fr_t b1) {
    fr_pair_t r;
    r.a = mul(a0, b0);
    r.b = mul(a1, b1);
    return r;
}

template <int T>
__device__ __forceinline__ void dit_stage4(fr_t (&x)[4], uint32_t sigma, uint32_t base, uint32_t b0, uint32_t logR,
                                           const uint4* t_lo, const uint4* t_hi) {
    constexpr int a = 0, c = T == 0 ? 2 : 1;            // lower rows of the two butterflies
    constexpr int a1 = a | (1 << T), c1 = c | (1 << T);
    const uint32_t mask = (1u << (sigma - 1)) - 1;
    const uint32_t ea = ((base + ((uint32_t)a << b0)) & mask) << (logR - sigma);
    const uint32_t ec = ((base + ((uint32_t)c << b0)) & mask) << (logR - sigma);
    const fr_pair_t pr = fr_mul2_nl(x[a1], ld_sm(t_lo, t_hi, ea), x[c1], ld_sm(t_lo, t_hi, ec));
    const fr_t ua = x[a], uc = x[c];
    x[a] = add(ua, pr.a);
    x[a1] = sub(ua, pr.a);
    x[c] = add(uc, pr.b);
    x[c1] = sub(uc, pr.b);
}

__global__ void __launch_bounds__(256) ntt_pass_r4d_kernel(NttPassParams p) {
    extern __shared__ uint4 smem[];
    const uint32_t logR = p.logR, logC = p.logC, logNs = p.logNs;
    const uint32_t R = 1u << logR, C = 1u << logC, RC = R << logC;
    uint4* s_lo = smem;
    uint4* s_hi = smem + RC;
    uint4* t_lo = smem + 2 * RC;
    uint4* t_hi = t_lo + (R >> 1);
    const uint32_t tid = threadIdx.x, nt = blockDim.x;  // nt == RC / 4
    const size_t n = (size_t)1 << p.logn;
    const fr_t* in = p.in + (size_t)blockIdx.y * n;
    fr_t* out = p.out + (size_t)blockIdx.y * n;
    const uint32_t j0 = blockIdx.x << logC;
    const uint32_t c = tid & (C - 1), g = tid >> logC;  // column, group of 4 rows
    const uint32_t j = j0 + c;
    const uint32_t ns_mask = (1u << logNs) - 1;

    for (uint32_t i = tid; i < (R >> 1); i += nt) {
        fr_t w = table_pow(p.tw_lo, p.tw_hi, i << (NTT_MAX_LOG - logR));
        st_sm(t_lo, t_hi, i, w);
    }
    __syncthreads();

    // step 1: rows brev(4 g + a), inter-pass twiddle (two products per call), stages 1..2
    fr_t x[4];
    uint32_t base = g << 2, b0 = 0;
    {
        const uint32_t stride_in = (uint32_t)(n >> logR);
        const uint32_t tw_shift = NTT_MAX_LOG - (logNs + logR);
        const uint32_t k = j & ns_mask;
        uint32_t rr[4];
#pragma unroll
        for (int a = 0; a < 4; a++) {
            rr[a] = __brev(base + a) >> (32 - logR);
            const uint32_t gi = j + rr[a] * stride_in;
            fr_t v = ld_fr(in + gi);
            if (p.coset_in) v = fr_mul_nl(v, table_pow(p.cs_lo, p.cs_hi, gi));
            x[a] = v;
        }
        if (logNs) {
#pragma unroll
            for (int a = 0; a < 4; a += 2) {
                fr_t w0, w1;
                if (p.tw_direct) {
                    w0 = ld_fr(p.tw_direct + k * rr[a]);
                    w1 = ld_fr(p.tw_direct + k * rr[a + 1]);
                } else {
                    w0 = table_pow(p.tw_lo, p.tw_hi, (k * rr[a]) << tw_shift);
                    w1 = table_pow(p.tw_lo, p.tw_hi, (k * rr[a + 1]) << tw_shift);
                }
                const fr_pair_t pr = fr_mul2_nl(x[a], w0, x[a + 1], w1);
                x[a] = pr.a;
                x[a + 1] = pr.b;
            }
        }
        // stage 1: twiddles 1; stage 2: pairs (0,2) twiddle 1, (1,3) twiddle w_R^(R/4)
        {
            fr_t u = x[0], t = x[1];
            x[0] = add(u, t);
            x[1] = sub(u, t);
            u = x[2];
            t = x[3];
            x[2] = add(u, t);
            x[3] = sub(u, t);
            u = x[0];
            t = x[2];
            x[0] = add(u, t);
            x[2] = sub(u, t);
            u = x[1];
            t = fr_mul_nl(x[3], ld_sm(t_lo, t_hi, 1u << (logR - 2)));
            x[1] = add(u, t);
            x[3] = sub(u, t);
        }
    }
    uint32_t done = 2;
    while (done < logR) {
#pragma unroll
        for (int a = 0; a < 4; a++) st_sm(s_lo, s_hi, ((base + ((uint32_t)a << b0)) << logC) + c, x[a]);
        __syncthreads();
        const uint32_t k = logR - done < 2 ? 1 : 2;
        const uint32_t s = done + 1;
        b0 = s - 1 < logR - 2 ? s - 1 : logR - 2;
        base = ((g >> b0) << (b0 + 2)) | (g & ((1u << b0) - 1));
#pragma unroll
        for (int a = 0; a < 4; a++) x[a] = ld_sm(s_lo, s_hi, ((base + ((uint32_t)a << b0)) << logC) + c);
        if (k == 2) {
            dit_stage4<0>(x, s, base, b0, logR, t_lo, t_hi);
            dit_stage4<1>(x, s + 1, base, b0, logR, t_lo, t_hi);
        } else {
            dit_stage4<1>(x, s, base, b0, logR, t_lo, t_hi);
        }
        done += k;
    }
    if (logNs != 0) {
        const size_t o0 = ((size_t)(j >> logNs) << (logNs + logR)) + (j & ns_mask);
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const uint32_t r = base + ((uint32_t)a << b0);
            const size_t o = o0 + ((size_t)r << logNs);
            fr_t v = x[a];
            if (p.scale_out == 1) v = fr_mul_nl(v, p.scale);
            else if (p.scale_out == 2) v = fr_mul_nl(v, table_pow(p.cs_lo, p.cs_hi, (uint32_t)o));
            st_fr(out + o, v);
        }
        return;
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; a++) st_sm(s_lo, s_hi, ((base + ((uint32_t)a << b0)) << logC) + c, x[a]);
    __syncthreads();
    for (uint32_t idx = tid; idx < RC; idx += nt) {
        const uint32_t r = idx & (R - 1), cc = idx >> logR;
        const size_t o = ((size_t)(j0 + cc) << logR) + r;
        fr_t v = ld_sm(s_lo, s_hi, (r << logC) + cc);
        if (p.scale_out == 1) v = fr_mul_nl(v, p.scale);
        else if (p.scale_out == 2) v = fr_mul_nl(v, table_pow(p.cs_lo, p.cs_hi, (uint32_t)o));
        st_fr(out + o, v);
    }
}

// out[i] = pre * base^(i << shift)
__global__ void pow_table_kernel(fr_t* out, fr_t base, fr_t pre, uint32_t count, uint32_t shift) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    fr_t v = pow_u64(base, (uint64_t)i << shift);
    st_fr(out + i, mul(v, pre));
}

// out[e] = root^(e << shift) from the two-level table: the per-pass inter-pass twiddle table
__global__ void direct_table_kernel(fr_t* out, const fr_t* lo, const fr_t* hi, uint32_t count, uint32_t shift) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    st_fr(out + e, table_pow(lo, hi, e << shift));
}

__global__ void pointwise_mul_kernel(fr_t* a, const fr_t* b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t step = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += step) st_fr(a + i, mul(ld_fr(a + i), ld_fr(b + i)));
}

__global__ void scale_kernel(fr_t* a, fr_t s, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t step = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += step) st_fr(a + i, mul(ld_fr(a + i), s));
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static fr_t fr_from_u64x4(const uint64_t v[4]) {
    fr_t r;
    for (int i = 0; i < 4; i++) {
        r.l[2 * i] = (uint32_t)v[i];
        r.l[2 * i + 1] = (uint32_t)(v[i] >> 32);
    }
    return r;
}
static fr_t fr_from_small(uint64_t x) {  // canonical small integer -> Montgomery (host templates)
    fr_t r = fr_t::zero();
    r.l[0] = (uint32_t)x;
    r.l[1] = (uint32_t)(x >> 32);
    return to_mont(r);
}

// scalar.rs:208-213 ROOT_OF_UNITY (Montgomery limbs), a primitive 2^32-th root of unity
static const uint64_t ROOT_OF_UNITY_MONT[4] = {0xb9b58d8c5f0e466aull, 0x5b1b4c801819d7ecull,
                                               0x0af53ae352a31e64ull, 0x5bf3adda19e9b27bull};

int launch_pow_table(bpk_ctx* ctx, fr_t* d_out, const fr_t& base, const fr_t& pre, uint32_t count,
                            uint32_t shift) {
    pow_table_kernel<<<(count + 127) / 128, 128, 0, ctx->stream>>>(d_out, base, pre, count, shift);
    count_launch(ctx);
    BPK_CUDA(cudaGetLastError());
    return BPK_OK;
}

int ntt_init_tables(bpk_ctx* ctx) {
    fr_t root = fr_from_u64x4(ROOT_OF_UNITY_MONT);
    for (int i = 0; i < 32 - NTT_MAX_LOG; i++) root = sqr(root);  // primitive 2^NTT_MAX_LOG-th root
    fr_t roots[2] = {root, inv(root)};
    BPK_CUDA(cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
    BPK_CUDA(cudaFuncSetAttribute(ntt_pass_r8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
    BPK_CUDA(cudaFuncSetAttribute(ntt_pass_r4d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
    for (int d = 0; d < 2; d++) {
        BPK_CUDA(cudaMalloc(&ctx->tw_lo[d], sizeof(fr_t) << TW_LO_BITS));
        BPK_CUDA(cudaMalloc(&ctx->tw_hi[d], sizeof(fr_t) << TW_HI_BITS));
        BPK_CUDA(cudaMalloc(&ctx->coset_lo[d], sizeof(fr_t) << TW_LO_BITS));
        BPK_CUDA(cudaMalloc(&ctx->coset_hi[d], sizeof(fr_t) << TW_HI_BITS));
        BPK_TRY(launch_pow_table(ctx, ctx->tw_lo[d], roots[d], fr_t::one(), 1u << TW_LO_BITS, 0));
        BPK_TRY(launch_pow_table(ctx, ctx->tw_hi[d], roots[d], fr_t::one(), 1u << TW_HI_BITS, TW_LO_BITS));
    }
    BPK_CUDA(cudaStreamSynchronize(ctx->stream));
    return BPK_OK;
}

// frees every per-size twiddle table (they are rebuilt on demand); earlier passes on the stream may still read them
int ntt_drop_direct_tables(bpk_ctx* ctx) {
    BPK_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int d = 0; d < 2; d++)
        for (auto& t : ctx->tw_direct[d])
            if (t) {
                cudaFree(t);
                t = nullptr;
            }
    ctx->tw_direct_bytes = 0;
    return BPK_OK;
}

static void plan_passes(uint32_t logn, uint32_t max_logR, std::vector<uint32_t>& out) {
    out.clear();
    if (logn == 0) return;
    uint32_t passes = (logn + max_logR - 1) / max_logR;
    uint32_t rem = logn;
    for (uint32_t i = 0; i < passes; i++) {
        uint32_t left = passes - i;
        uint32_t a = (rem + left - 1) / left;
        out.push_back(a);
        rem -= a;
    }
}

int ntt_run(bpk_ctx* ctx, const fr_t* d_in, fr_t* d_out, size_t n, size_t batch, bool inverse,
            const fr_t* shift) {
    if (n == 0 || (n & (n - 1)) != 0) return BPK_ERR_NOT_POW2;
    uint32_t logn = 0;
    while (((size_t)1 << logn) < n) logn++;
    if (logn > (uint32_t)NTT_MAX_LOG) return BPK_ERR_TOO_LARGE;
    if (batch == 0) return BPK_OK;
    if (batch > 65535) return BPK_ERR_TOO_LARGE;
    const int dir = inverse ? 1 : 0;
    // the 2- / 3-pass plans need one / two scratch copies of the whole batch: transform large batches a few rows at a
    // time so that the scratch stays under ntt.scratch_mib (a 10 x 2^26 batch would otherwise pin 43 GB)
    {
        const size_t cap = (size_t)ctx->opt_ntt_scratch_mib << 20;
        const size_t row = n * sizeof(fr_t);
        if (batch > 1 && n * batch * sizeof(fr_t) > cap) {
            const size_t rows = cap / row > 1 ? cap / row : 1;
            for (size_t b = 0; b < batch; b += rows) {
                const size_t cnt = batch - b < rows ? batch - b : rows;
                BPK_TRY(ntt_run(ctx, d_in + b * n, d_out + b * n, n, cnt, inverse, shift));
            }
            return BPK_OK;
        }
    }
    const size_t bytes = n * batch * sizeof(fr_t);

    // n^-1 (inverse) and coset tables
    fr_t scale = fr_t::one();
    if (inverse) scale = inv(fr_from_small((uint64_t)n));
    if (shift) {
        // the inverse table folds n^-1 into its low half, so it is keyed by (shift, n)
        bool need = !ctx->coset_valid[dir] || ctx->coset_shift[dir] != *shift || (inverse && ctx->coset_n != n);
        if (need) {
            StageTimer t(ctx, "ntt.coset_table");
            fr_t g = inverse ? inv(*shift) : *shift;
            BPK_TRY(launch_pow_table(ctx, ctx->coset_lo[dir], g, scale, 1u << TW_LO_BITS, 0));
            BPK_TRY(launch_pow_table(ctx, ctx->coset_hi[dir], g, fr_t::one(), 1u << TW_HI_BITS, TW_LO_BITS));
            ctx->coset_shift[dir] = *shift;
            ctx->coset_valid[dir] = true;
            if (inverse) ctx->coset_n = n;
            t.end();
        }
    }

    if (logn == 0) {  // n == 1: the transform is the identity (times n^-1 = 1; coset factor g^0 = 1)
        if (d_in != d_out) BPK_CUDA(cudaMemcpyAsync(d_out, d_in, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        return BPK_OK;
    }

    // transforms too small to fill the GPU with the 8-row kernel: 4 rows per thread on 2^9-element tiles
    // (2^16: 0.046 -> 0.033 ms, profiles/r1_ntt_kernel_ab.md)
    const bool big = (n * batch) >= ((size_t)1 << 18);
    const bool small_auto = !big && ctx->opt_ntt_kernel == 0 && logn > 9;
    uint32_t tile_log = (uint32_t)ctx->opt_ntt_tile_log2;  // log2(R * C)
    if (small_auto && tile_log > 9) tile_log = 9;
    // auto (0): one pass up to the tile size, otherwise radices <= 2^8 so that a tile keeps >= 4 adjacent columns
    // (3 x 2^20: 0.67 -> 0.61 ms against two passes of 2^10, profiles/r1_ntt_kernel_ab.md)
    uint32_t max_logR = ctx->opt_ntt_max_radix_log2 ? (uint32_t)ctx->opt_ntt_max_radix_log2 : (logn <= tile_log ? tile_log : 8);
    if (max_logR > tile_log) max_logR = tile_log;
    if (max_logR < 1) max_logR = 1;
    std::vector<uint32_t> plan;
    plan_passes(logn, max_logR, plan);
    const size_t np = plan.size();

    fr_t* tmp[2] = {nullptr, nullptr};
    if (np >= 2) BPK_TRY(ws_reserve(ctx, 0, bytes, (void**)&tmp[0]));
    if (np >= 3) BPK_TRY(ws_reserve(ctx, 1, bytes, (void**)&tmp[1]));

    // which kernel runs each pass (0: one stage per barrier, 1: register-blocked 8 rows, 2: 4 rows): a pass may hand
    // on values in [0, 2q) only to the register-blocked kernel
    std::vector<int> kind(np);
    std::vector<uint32_t> tile_c(np);
    {
        uint32_t ns = 0;
        for (size_t pi = 0; pi < np; pi++) {
            uint32_t logR = plan[pi];
            uint32_t logC = tile_log - logR;
            if (logC > logn - logR) logC = logn - logR;
            if (ns && logC > ns) logC = ns;
            tile_c[pi] = logC;
            if ((ctx->opt_ntt_kernel == 3 || small_auto) && logR >= 2 && logR + logC <= 10)
                kind[pi] = 2;
            else if (logR >= 3 && logR + logC <= 11 && (ctx->opt_ntt_kernel == 0 ? big : ctx->opt_ntt_kernel == 2))
                kind[pi] = 1;
            else
                kind[pi] = 0;
            ns += logR;
        }
    }
    StageTimer t(ctx, "ntt.pass");
    uint32_t logNs = 0;
    const fr_t* src = d_in;
    for (size_t pi = 0; pi < np; pi++) {
        uint32_t logR = plan[pi];
        uint32_t logC = tile_c[pi];
        fr_t* dst = (pi + 1 == np) ? d_out : tmp[pi & 1];
        NttPassParams p;
        p.in = src;
        p.out = dst;
        p.tw_lo = ctx->tw_lo[dir];
        p.tw_hi = ctx->tw_hi[dir];
        p.cs_lo = ctx->coset_lo[dir];
        p.cs_hi = ctx->coset_hi[dir];
        p.tw_direct = nullptr;
        if (logNs && logNs + logR <= (uint32_t)ctx->opt_ntt_direct_max_log2) {
            // w_{Ns R}^e, e < Ns R: HBM is idle in this kernel (6 % of peak) while the multiplier pipe is the
            // limiter, so 32 B more traffic per element buys one multiplication less per element
            const uint32_t lg = logNs + logR;
            fr_t*& tab = ctx->tw_direct[dir][lg];
            if (tab == nullptr) {
                // the per-size tables are a cache: keep their total under the budget (a context that transforms at many
                // sizes would otherwise pin GiBs next to the SRS), and give them back when HBM is short
                const size_t want = sizeof(fr_t) << lg;
                const size_t budget = (size_t)ctx->opt_ntt_direct_budget_mib << 20;
                if (want <= budget) {
                    if (ctx->tw_direct_bytes + want > budget) BPK_TRY(ntt_drop_direct_tables(ctx));
                    cudaError_t e = cudaMalloc(&tab, want);
                    if (e == cudaErrorMemoryAllocation) {
                        cudaGetLastError();
                        BPK_TRY(ntt_drop_direct_tables(ctx));
                        e = cudaMalloc(&tab, want);
                    }
                    if (e == cudaErrorMemoryAllocation) {  // still no room: this pass multiplies two table entries instead
                        cudaGetLastError();
                        tab = nullptr;
                    } else {
                        BPK_CUDA(e);
                        ctx->tw_direct_bytes += want;
                        direct_table_kernel<<<(unsigned)(((size_t)1 << lg) + 255) / 256, 256, 0, ctx->stream>>>(
                            tab, ctx->tw_lo[dir], ctx->tw_hi[dir], 1u << lg, NTT_MAX_LOG - lg);
                        count_launch(ctx);
                        BPK_CUDA(cudaGetLastError());
                    }
                }
            }
            p.tw_direct = tab;
        }
        p.logn = logn;
        p.logR = logR;
        p.logC = logC;
        p.logNs = logNs;
        p.coset_in = (shift && !inverse && pi == 0) ? 1 : 0;
        p.scale_out = 0;
        p.scale = scale;
        if (pi + 1 == np && inverse) p.scale_out = shift ? 2 : 1;
        p.lazy_out = (BPK_NTT_LAZY && kind[pi] == 1 && pi + 1 < np && kind[pi + 1] == 1) ? 1 : 0;
        size_t smem = ((size_t)2 << (logR + logC)) * sizeof(uint4) + ((size_t)1 << logR) * sizeof(uint4);
        if (smem > 200 * 1024) return BPK_ERR_INVALID_ARG;
        dim3 grid((unsigned)(n >> (logR + logC)), (unsigned)batch);
        unsigned threads = (unsigned)ctx->opt_ntt_threads;
        if (threads == 0) threads = (logR + logC >= 12) ? 1024 : 256;
        // register-blocked kernel unless the transform is too small to fill the GPU with 128-thread CTAs
        if (kind[pi] == 2)
            ntt_pass_r4d_kernel<<<grid, 1u << (logR + logC - 2), smem, ctx->stream>>>(p);
        else if (kind[pi] == 1)
            ntt_pass_r8_kernel<<<grid, 1u << (logR + logC - 3), smem, ctx->stream>>>(p);
        else
            ntt_pass_kernel<<<grid, threads, smem, ctx->stream>>>(p);
        count_launch(ctx);
        BPK_CUDA(cudaGetLastError());
        src = dst;
        logNs += logR;
    }
    t.end();
    return BPK_OK;
}

int pointwise_mul(bpk_ctx* ctx, fr_t* d_a, const fr_t* d_b, size_t n) {
    StageTimer t(ctx, "fr.pointwise");
    size_t blocks = (n + 255) / 256;
    size_t cap = (size_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks == 0) blocks = 1;
    pointwise_mul_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_a, d_b, n);
    count_launch(ctx);
    BPK_CUDA(cudaGetLastError());
    t.end();
    return BPK_OK;
}

}  // namespace bpk
