// Device-resident Fr vector / polynomial primitives: the O(n) work of the prover rounds that sits between
// the NTTs and the commitments.  Reference counterparts (all serial Rust on
// Vec<Scalar>): Polynomial Add / Sub / Mul<Scalar> (src/polynomial.rs:57-187), coeffs_evaluate (:34-45),
// Div by X^n - 1 and by X - zeta (:314-380), the grand-product loop of round 2 (src/prover.rs:286-317),
// monomial_z_to_z_omega (src/prover.rs:661-674).  Every value is a canonical Montgomery residue, so the
// results are bit-identical to the reference's whatever the evaluation order.
#include "internal.cuh"

namespace bpk {

__device__ __forceinline__ fr_t pld(const fr_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    fr_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void pst(fr_t* p, const fr_t& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// base^E from a two-level table (lo: base^i, i < 2^13; hi: base^(i << 13))
__device__ __forceinline__ fr_t ptable(const fr_t* lo, const fr_t* hi, uint32_t E) {
    fr_t a = pld(lo + (E & ((1u << TW_LO_BITS) - 1)));
    uint32_t h = E >> TW_LO_BITS;
    if (h == 0) return a;
    return mul(a, pld(hi + h));
}

// ---- elementwise ---------------------------------------------------------------------------------
// op: 0 a+b, 1 a-b, 2 a*b, 3 a*s, 4 a + s*b, 5 a+s (every element), 6 a - s*b
__global__ void fr_vec_op_kernel(int op, const fr_t* __restrict__ a, const fr_t* __restrict__ b, fr_t s,
                                 fr_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t step = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        fr_t x = pld(a + i), r;
        switch (op) {
            case 0: r = add(x, pld(b + i)); break;
            case 1: r = sub(x, pld(b + i)); break;
            case 2: r = mul(x, pld(b + i)); break;
            case 3: r = mul(x, s); break;
            case 4: r = add(x, mul(s, pld(b + i))); break;
            case 5: r = add(x, s); break;
            default: r = sub(x, mul(s, pld(b + i))); break;
        }
        pst(out + i, r);
    }
}

// out[i] = a[i] * c0 * g^i   (lo table already carries c0); reverse != 0 writes out[n-1-i] instead
__global__ void fr_scale_powers_kernel(const fr_t* __restrict__ a, const fr_t* __restrict__ lo,
                                       const fr_t* __restrict__ hi, fr_t* __restrict__ out, size_t n, int reverse) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t step = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        fr_t r = mul(pld(a + i), ptable(lo, hi, (uint32_t)i));
        pst(out + (reverse ? n - 1 - i : i), r);
    }
}

// q[i] = I[n-2-i] * zinv^(i+1)  (lo table carries the extra zinv)
__global__ void fr_div_linear_finish_kernel(const fr_t* __restrict__ I, const fr_t* __restrict__ lo,
                                            const fr_t* __restrict__ hi, fr_t* __restrict__ q, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t step = (size_t)gridDim.x * blockDim.x;
    for (; i + 1 < n; i += step) pst(q + i, mul(pld(I + (n - 2 - i)), ptable(lo, hi, (uint32_t)i)));
}

// q[i] = sum_{k >= 1} c[i + k n]    (c / (X^n - 1), remainder dropped)
__global__ void fr_div_vanishing_kernel(const fr_t* __restrict__ c, size_t len, size_t n, fr_t* __restrict__ q) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t step = (size_t)gridDim.x * blockDim.x;
    for (; i + n < len; i += step) {
        fr_t acc = pld(c + i + n);
        for (size_t j = i + 2 * n; j < len; j += n) acc = add(acc, pld(c + j));
        pst(q + i, acc);
    }
}

// ---- reductions ------------------------------------------------------------------------------------
// block partial sums of c[i] * x^i
__global__ void __launch_bounds__(256) fr_eval_partial_kernel(const fr_t* __restrict__ c, const fr_t* __restrict__ lo,
                                                               const fr_t* __restrict__ hi, size_t n,
                                                               fr_t* __restrict__ partial) {
    __shared__ fr_t sm[256];
    fr_t acc = fr_t::zero();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t step = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += step) acc = add(acc, mul(pld(c + i), ptable(lo, hi, (uint32_t)i)));
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s >= 1; s >>= 1) {
        if ((int)threadIdx.x < s) sm[threadIdx.x] = add(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) pst(partial + blockIdx.x, sm[0]);
}
__global__ void __launch_bounds__(256) fr_sum_kernel(const fr_t* __restrict__ in, size_t n, fr_t* __restrict__ out) {
    __shared__ fr_t sm[256];
    fr_t acc = fr_t::zero();
    for (size_t i = threadIdx.x; i < n; i += 256) acc = add(acc, pld(in + i));
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s >= 1; s >>= 1) {
        if ((int)threadIdx.x < s) sm[threadIdx.x] = add(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) pst(out, sm[0]);
}

// ---- round 2: grand product -------------------------------------------------------------------------
// r[i] = (A+b w^i+g)(B+b k1 w^i+g)(C+b k2 w^i+g) / ((A+b s1+g)(B+b s2+g)(C+b s3+g))     prover.rs:286-317
// Each thread takes RATIO_BATCH consecutive rows and shares one field inversion between them (Montgomery's
// trick: invert the product of the denominators, peel the factors off again).
constexpr int RATIO_BATCH = 4;
__global__ void __launch_bounds__(128) plonk_ratio_kernel(const fr_t* __restrict__ A, const fr_t* __restrict__ B,
                                                           const fr_t* __restrict__ C, const fr_t* __restrict__ s1,
                                                           const fr_t* __restrict__ s2, const fr_t* __restrict__ s3,
                                                           const fr_t* __restrict__ tw_lo, const fr_t* __restrict__ tw_hi,
                                                           uint32_t logn, fr_t beta, fr_t gamma, fr_t k1, fr_t k2,
                                                           fr_t* __restrict__ r, size_t n) {
    size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * RATIO_BATCH;
    if (i0 == 0) pst(r + n, fr_t::one());  // pad so that the exclusive scan yields Z[n]
    if (i0 >= n) return;
    fr_t num[RATIO_BATCH], pre[RATIO_BATCH], den[RATIO_BATCH];
    fr_t run = fr_t::one();
#pragma unroll
    for (int j = 0; j < RATIO_BATCH; j++) {
        size_t i = i0 + j;
        if (i >= n) {
            num[j] = den[j] = fr_t::one();
        } else {
            fr_t w = ptable(tw_lo, tw_hi, (uint32_t)i << (NTT_MAX_LOG - logn));
            fr_t bw = mul(beta, w);
            fr_t a = add(pld(A + i), gamma), b = add(pld(B + i), gamma), c = add(pld(C + i), gamma);
            num[j] = mul(mul(add(a, bw), add(b, mul(bw, k1))), add(c, mul(bw, k2)));
            den[j] = mul(mul(add(a, mul(beta, pld(s1 + i))), add(b, mul(beta, pld(s2 + i)))),
                         add(c, mul(beta, pld(s3 + i))));
        }
        pre[j] = run;             // product of den[0..j)
        run = mul(run, den[j]);
    }
    fr_t suffix_inv = inv(run);   // 1 / prod den
#pragma unroll
    for (int j = RATIO_BATCH - 1; j >= 0; j--) {
        size_t i = i0 + j;
        if (i < n) pst(r + i, mul(num[j], mul(suffix_inv, pre[j])));   // 1/den[j] = pre[j] / prod den[0..j]
        suffix_inv = mul(suffix_inv, den[j]);
    }
}

// ---- round 3: quotient evaluations on the coset g*<w_D> ------------------------------------------------
// One fused pass over the D-point coset evaluations of the 15 prover polynomials, in two row groups (rows D
// apart): per-proof rows `wv` = a b c z PI, per-circuit rows `cv` = ql qr qm qo qc s1 s2 s3 L1 X (these can
// be kept in HBM across proofs).
// t(x) = [ gate + alpha * perm + alpha^2 * (z - 1) L1 ] / Z_H(x)                      prover.rs:370-452
//   gate = a ql + b qr + a b qm + c qo + PI + qc
//   perm = (a + beta x + gamma)(b + beta k1 x + gamma)(c + beta k2 x + gamma) z
//        - (a + beta s1 + gamma)(b + beta s2 + gamma)(c + beta s3 + gamma) z(w x)
// z(w x) is z's own row rotated by D / n positions; 1 / Z_H takes D / n distinct values on the coset.
// The reference forms the same numerator by 16 full polynomial products (3 transforms each) and divides by
// long division; the polynomial is the same, so its coefficients are.
struct QuotientParams {
    fr_t beta, gamma, alpha, alpha2, k1, k2;
};
// Sharded form (one rank of a multi-GPU proof evaluates t on ONE sub-coset of the quotient domain): z(w x) is not a
// rotation of z's row there, so it comes as a sixth witness row (`zw_row`), and 1 / Z_H has `ratio` = max(1, 4 / ranks)
// distinct values.
__global__ void __launch_bounds__(256) plonk_quotient_kernel(const fr_t* __restrict__ wv, const fr_t* __restrict__ cv,
                                                              size_t D, uint32_t ratio, int zw_row,
                                                              const fr_t* __restrict__ zh_inv, QuotientParams P,
                                                              fr_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t step = (size_t)gridDim.x * blockDim.x;
    for (; i < D; i += step) {
        auto wit = [&](int r) { return pld(wv + (size_t)r * D + i); };
        auto cir = [&](int r) { return pld(cv + (size_t)r * D + i); };
        fr_t a = wit(0), b = wit(1), c = wit(2), z = wit(3);
        fr_t gate = mul(a, cir(0));
        gate = add(gate, mul(b, cir(1)));
        gate = add(gate, mul(mul(a, b), cir(2)));
        gate = add(gate, mul(c, cir(3)));
        gate = add(gate, add(cir(4), wit(4)));
        fr_t bx = mul(P.beta, cir(9));
        fr_t ag = add(a, P.gamma), bg = add(b, P.gamma), cg = add(c, P.gamma);
        fr_t lhs = mul(mul(add(ag, bx), add(bg, mul(bx, P.k1))), mul(add(cg, mul(bx, P.k2)), z));
        fr_t zw;
        if (zw_row) {
            zw = wit(5);
        } else {
            size_t j = i + ratio;
            if (j >= D) j -= D;
            zw = pld(wv + 3 * D + j);
        }
        fr_t rhs = mul(mul(add(ag, mul(P.beta, cir(5))), add(bg, mul(P.beta, cir(6)))),
                       mul(add(cg, mul(P.beta, cir(7))), zw));
        fr_t first = mul(sub(z, fr_t::one()), cir(8));
        fr_t num = add(gate, add(mul(P.alpha, sub(lhs, rhs)), mul(P.alpha2, first)));
        pst(out + i, mul(num, pld(zh_inv + (i & (ratio - 1)))));
    }
}

// ---- prefix scan over Fr (sum or product) -------------------------------------------------------------
// Three launches: per-tile aggregates, one block that scans the aggregates, per-tile scan with the tile's base.
// Both operations are associative and commutative on canonical residues, so the order of combination does not
// show in the result.  MUL: running product (grand product of round 2), else running sum (division by X - zeta).
constexpr int FSCAN_THREADS = 256, FSCAN_ITEMS = 4, FSCAN_TILE = FSCAN_THREADS * FSCAN_ITEMS;

template <bool MUL>
__device__ __forceinline__ fr_t fscan_op(const fr_t& a, const fr_t& b) {
    return MUL ? mul(a, b) : add(a, b);
}
template <bool MUL>
__device__ __forceinline__ fr_t fscan_identity() {
    return MUL ? fr_t::one() : fr_t::zero();
}
__device__ __forceinline__ fr_t shfl_up_fr(const fr_t& v, int d) {
    fr_t r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_up_sync(0xffffffffu, v.l[i], d);
    return r;
}
// inclusive scan of one value per thread across the block; `total` = the block aggregate (all threads)
template <bool MUL>
__device__ __forceinline__ fr_t fscan_block(fr_t v, fr_t& total, fr_t* sm /* FSCAN_THREADS / 32 + 1 */) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        fr_t t = shfl_up_fr(v, d);
        if (lane >= (uint32_t)d) v = fscan_op<MUL>(t, v);
    }
    __syncthreads();
    if (lane == 31) sm[warp] = v;
    __syncthreads();
    if (warp == 0) {
        fr_t w = lane < FSCAN_THREADS / 32 ? sm[lane] : fscan_identity<MUL>();
#pragma unroll
        for (int d = 1; d < FSCAN_THREADS / 32; d <<= 1) {
            fr_t t = shfl_up_fr(w, d);
            if (lane >= (uint32_t)d) w = fscan_op<MUL>(t, w);
        }
        if (lane < FSCAN_THREADS / 32) sm[lane] = w;  // inclusive over the warps
    }
    __syncthreads();
    total = sm[FSCAN_THREADS / 32 - 1];
    if (warp > 0) v = fscan_op<MUL>(sm[warp - 1], v);
    return v;
}

template <bool MUL>
__global__ void __launch_bounds__(FSCAN_THREADS) fr_scan_tiles_kernel(const fr_t* __restrict__ in, size_t n,
                                                                       fr_t* __restrict__ aggregates) {
    __shared__ fr_t sm[FSCAN_THREADS / 32 + 1];
    const size_t first = (size_t)blockIdx.x * FSCAN_TILE + (size_t)threadIdx.x * FSCAN_ITEMS;
    fr_t acc = fscan_identity<MUL>();
#pragma unroll
    for (int k = 0; k < FSCAN_ITEMS; k++)
        if (first + k < n) acc = fscan_op<MUL>(acc, pld(in + first + k));
    fr_t total;
    fscan_block<MUL>(acc, total, sm);
    if (threadIdx.x == 0) pst(aggregates + blockIdx.x, total);
}

// exclusive scan of the tile aggregates in place, by one block
template <bool MUL>
__global__ void __launch_bounds__(FSCAN_THREADS) fr_scan_aggregates_kernel(fr_t* __restrict__ aggregates, size_t count) {
    __shared__ fr_t sm[FSCAN_THREADS / 32 + 1];
    const size_t per = (count + FSCAN_THREADS - 1) / FSCAN_THREADS;
    const size_t lo = threadIdx.x * per, hi = lo + per < count ? lo + per : count;
    fr_t acc = fscan_identity<MUL>();
    for (size_t i = lo; i < hi; i++) acc = fscan_op<MUL>(acc, pld(aggregates + i));
    fr_t total;
    const fr_t incl = fscan_block<MUL>(acc, total, sm);
    // exclusive base of this thread's range = inclusive value of the previous thread
    __shared__ fr_t prev[FSCAN_THREADS];
    prev[threadIdx.x] = incl;
    __syncthreads();
    fr_t run = threadIdx.x ? prev[threadIdx.x - 1] : fscan_identity<MUL>();
    for (size_t i = lo; i < hi; i++) {
        const fr_t v = pld(aggregates + i);
        pst(aggregates + i, run);
        run = fscan_op<MUL>(run, v);
    }
}

// out[i] = in[0] o .. o in[i] (inclusive) or init o in[0] o .. o in[i-1] (exclusive)
template <bool MUL, bool EXCLUSIVE>
__global__ void __launch_bounds__(FSCAN_THREADS) fr_scan_apply_kernel(const fr_t* __restrict__ in, size_t n,
                                                                       const fr_t* __restrict__ bases, fr_t init,
                                                                       fr_t* __restrict__ out) {
    __shared__ fr_t sm[FSCAN_THREADS / 32 + 1];
    __shared__ fr_t prev[FSCAN_THREADS];
    const size_t first = (size_t)blockIdx.x * FSCAN_TILE + (size_t)threadIdx.x * FSCAN_ITEMS;
    fr_t x[FSCAN_ITEMS];
    fr_t acc = fscan_identity<MUL>();
#pragma unroll
    for (int k = 0; k < FSCAN_ITEMS; k++) {
        x[k] = first + k < n ? pld(in + first + k) : fscan_identity<MUL>();
        acc = fscan_op<MUL>(acc, x[k]);
    }
    fr_t total;
    const fr_t incl = fscan_block<MUL>(acc, total, sm);
    prev[threadIdx.x] = incl;
    __syncthreads();
    fr_t run = fscan_op<MUL>(init, pld(bases + blockIdx.x));
    if (threadIdx.x) run = fscan_op<MUL>(run, prev[threadIdx.x - 1]);
#pragma unroll
    for (int k = 0; k < FSCAN_ITEMS; k++) {
        if (EXCLUSIVE && first + k < n) pst(out + first + k, run);
        run = fscan_op<MUL>(run, x[k]);
        if (!EXCLUSIVE && first + k < n) pst(out + first + k, run);
    }
}

template <bool MUL, bool EXCLUSIVE>
static int fr_scan(bpk_ctx* ctx, const fr_t* in, size_t n, const fr_t& init, fr_t* out) {
    if (n == 0) return BPK_OK;
    const size_t tiles = (n + FSCAN_TILE - 1) / FSCAN_TILE;
    fr_t* aggregates;
    BPK_TRY(ws_reserve(ctx, 14, tiles * sizeof(fr_t), (void**)&aggregates));
    fr_scan_tiles_kernel<MUL><<<(unsigned)tiles, FSCAN_THREADS, 0, ctx->stream>>>(in, n, aggregates);
    fr_scan_aggregates_kernel<MUL><<<1, FSCAN_THREADS, 0, ctx->stream>>>(aggregates, tiles);
    fr_scan_apply_kernel<MUL, EXCLUSIVE><<<(unsigned)tiles, FSCAN_THREADS, 0, ctx->stream>>>(in, n, aggregates, init, out);
    count_launch(ctx, 3);
    BPK_CUDA(cudaGetLastError());
    return BPK_OK;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static unsigned grid_for(bpk_ctx* ctx, size_t n, unsigned block) {
    size_t blocks = (n + block - 1) / block;
    size_t cap = (size_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks ? blocks : 1);
}

// two-level power table of g with g^0 replaced by c0 in the low half: lo[i] = c0 g^i, hi[i] = g^(i << 13)
static int power_tables(bpk_ctx* ctx, const fr_t& g, const fr_t& c0, size_t n, fr_t** lo, fr_t** hi) {
    size_t hi_count = (n >> TW_LO_BITS) + 1;
    fr_t* base;
    BPK_TRY(ws_reserve(ctx, 12, (((size_t)1 << TW_LO_BITS) + hi_count) * sizeof(fr_t), (void**)&base));
    *lo = base;
    *hi = base + ((size_t)1 << TW_LO_BITS);
    BPK_TRY(launch_pow_table(ctx, *lo, g, c0, 1u << TW_LO_BITS, 0));
    BPK_TRY(launch_pow_table(ctx, *hi, g, fr_t::one(), (uint32_t)hi_count, TW_LO_BITS));
    return BPK_OK;
}

int fr_vec_op(bpk_ctx* ctx, int op, const fr_t* a, const fr_t* b, const fr_t& s, fr_t* out, size_t n) {
    if (n == 0) return BPK_OK;
    StageTimer t(ctx, "fr.vec_op");
    fr_vec_op_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(op, a, b, s, out, n);
    count_launch(ctx);
    BPK_CUDA(cudaGetLastError());
    t.end();
    return BPK_OK;
}

int fr_scale_powers(bpk_ctx* ctx, const fr_t* a, const fr_t& g, const fr_t& c0, fr_t* out, size_t n) {
    if (n == 0) return BPK_OK;
    if (n > ((size_t)1 << NTT_MAX_LOG)) return BPK_ERR_TOO_LARGE;
    StageTimer t(ctx, "fr.scale_powers");
    fr_t *lo, *hi;
    BPK_TRY(power_tables(ctx, g, c0, n, &lo, &hi));
    fr_scale_powers_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(a, lo, hi, out, n, 0);
    count_launch(ctx);
    BPK_CUDA(cudaGetLastError());
    t.end();
    return BPK_OK;
}

int fr_poly_eval(bpk_ctx* ctx, const fr_t* c, size_t n, const fr_t& x, fr_t* d_out) {
    if (n > ((size_t)1 << NTT_MAX_LOG)) return BPK_ERR_TOO_LARGE;
    StageTimer t(ctx, "fr.eval");
    fr_t *lo, *hi;
    BPK_TRY(power_tables(ctx, x, fr_t::one(), n ? n : 1, &lo, &hi));
    unsigned blocks = grid_for(ctx, n ? n : 1, 256);
    fr_t* partial;
    BPK_TRY(ws_reserve(ctx, 13, (size_t)blocks * sizeof(fr_t), (void**)&partial));
    fr_eval_partial_kernel<<<blocks, 256, 0, ctx->stream>>>(c, lo, hi, n, partial);
    fr_sum_kernel<<<1, 256, 0, ctx->stream>>>(partial, blocks, d_out);
    count_launch(ctx, 2);
    BPK_CUDA(cudaGetLastError());
    t.end();
    return BPK_OK;
}

// several polynomials at the same point: one pair of power tables, the partial sums of all polynomials side by side
int fr_poly_eval_many(bpk_ctx* ctx, size_t count, const fr_t* const* c, const size_t* n, const fr_t& x, fr_t* d_out) {
    size_t n_max = 1;
    for (size_t i = 0; i < count; i++) {
        if (n[i] > ((size_t)1 << NTT_MAX_LOG)) return BPK_ERR_TOO_LARGE;
        if (n[i] > n_max) n_max = n[i];
    }
    StageTimer t(ctx, "fr.eval");
    fr_t *lo, *hi;
    BPK_TRY(power_tables(ctx, x, fr_t::one(), n_max, &lo, &hi));
    const unsigned stride = grid_for(ctx, n_max, 256);
    fr_t* partial;
    BPK_TRY(ws_reserve(ctx, 13, (size_t)stride * (count ? count : 1) * sizeof(fr_t), (void**)&partial));
    for (size_t i = 0; i < count; i++) {
        const unsigned blocks = grid_for(ctx, n[i] ? n[i] : 1, 256);
        fr_eval_partial_kernel<<<blocks, 256, 0, ctx->stream>>>(c[i], lo, hi, n[i], partial + (size_t)i * stride);
        fr_sum_kernel<<<1, 256, 0, ctx->stream>>>(partial + (size_t)i * stride, blocks, d_out + i);
    }
    count_launch(ctx, 2 * count);
    BPK_CUDA(cudaGetLastError());
    t.end();
    return BPK_OK;
}

// q = c / (X - root), remainder dropped; c has n coefficients, q gets n - 1
int fr_poly_div_linear(bpk_ctx* ctx, const fr_t* c, size_t n, const fr_t& root, fr_t* q) {
    if (n < 2) return BPK_OK;
    if (n > ((size_t)1 << NTT_MAX_LOG)) return BPK_ERR_TOO_LARGE;
    if (root.is_zero()) {  // c / X: shift down
        BPK_CUDA(cudaMemcpyAsync(q, c + 1, (n - 1) * sizeof(fr_t), cudaMemcpyDeviceToDevice, ctx->stream));
        return BPK_OK;
    }
    StageTimer t(ctx, "fr.div_linear");
    fr_t *lo, *hi, *tmp;
    BPK_TRY(ws_reserve(ctx, 13, 2 * n * sizeof(fr_t), (void**)&tmp));
    fr_t* rev = tmp;        // rev[n-1-j] = c_j root^j
    fr_t* scan = tmp + n;   // inclusive prefix sums of rev
    BPK_TRY(power_tables(ctx, root, fr_t::one(), n, &lo, &hi));
    fr_scale_powers_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(c, lo, hi, rev, n, 1);
    BPK_TRY((fr_scan<false, false>(ctx, rev, n, fr_t::zero(), scan)));
    // q_i = (sum_{j>i} c_j root^j) root^-(i+1)
    fr_t rinv = inv(root);
    BPK_TRY(power_tables(ctx, rinv, rinv, n, &lo, &hi));
    fr_div_linear_finish_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(scan, lo, hi, q, n);
    count_launch(ctx, 2);
    BPK_CUDA(cudaGetLastError());
    t.end();
    return BPK_OK;
}

int fr_poly_div_vanishing(bpk_ctx* ctx, const fr_t* c, size_t len, size_t n, fr_t* q) {
    if (len <= n) return BPK_OK;
    StageTimer t(ctx, "fr.div_vanishing");
    fr_div_vanishing_kernel<<<grid_for(ctx, len - n, 256), 256, 0, ctx->stream>>>(c, len, n, q);
    count_launch(ctx);
    BPK_CUDA(cudaGetLastError());
    t.end();
    return BPK_OK;
}

int plonk_grand_product(bpk_ctx* ctx, const fr_t* A, const fr_t* B, const fr_t* C, const fr_t* s1, const fr_t* s2,
                        const fr_t* s3, size_t n, const fr_t& beta, const fr_t& gamma, const fr_t& k1, const fr_t& k2,
                        fr_t* Z) {
    if (n == 0 || (n & (n - 1))) return BPK_ERR_NOT_POW2;
    uint32_t logn = 0;
    while (((size_t)1 << logn) < n) logn++;
    if (logn > (uint32_t)NTT_MAX_LOG) return BPK_ERR_TOO_LARGE;
    StageTimer t(ctx, "plonk.grand_product");
    fr_t* r;
    BPK_TRY(ws_reserve(ctx, 13, (n + 1) * sizeof(fr_t), (void**)&r));
    size_t ratio_threads = (n + RATIO_BATCH - 1) / RATIO_BATCH;
    plonk_ratio_kernel<<<(unsigned)((ratio_threads + 127) / 128), 128, 0, ctx->stream>>>(
        A, B, C, s1, s2, s3, ctx->tw_lo[0], ctx->tw_hi[0], logn, beta, gamma, k1, k2, r, n);
    BPK_TRY((fr_scan<true, true>(ctx, r, n + 1, fr_t::one(), Z)));
    count_launch(ctx, 1);
    BPK_CUDA(cudaGetLastError());
    t.end();
    return BPK_OK;
}

int plonk_quotient_evals(bpk_ctx* ctx, const fr_t* wv, const fr_t* cv, size_t D, size_t n, const fr_t& beta,
                         const fr_t& gamma,
                         const fr_t& alpha, const fr_t& k1, const fr_t& k2, const fr_t* zh_inv_host, fr_t* out) {
    if (D == 0 || (D & (D - 1)) || n == 0 || (n & (n - 1)) || D < n || D / n > 64) return BPK_ERR_INVALID_ARG;
    uint32_t ratio = (uint32_t)(D / n);
    StageTimer t(ctx, "plonk.quotient");
    fr_t* d_zh;
    BPK_TRY(ws_reserve(ctx, 12, 64 * sizeof(fr_t), (void**)&d_zh));
    BPK_CUDA(cudaMemcpyAsync(d_zh, zh_inv_host, ratio * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    QuotientParams P{beta, gamma, alpha, mul(alpha, alpha), k1, k2};
    plonk_quotient_kernel<<<grid_for(ctx, D, 256), 256, 0, ctx->stream>>>(wv, cv, D, ratio, 0, d_zh, P, out);
    count_launch(ctx);
    BPK_CUDA(cudaGetLastError());
    // the pageable H2D copy above is staged by the runtime before it returns, so zh_inv_host may be reused
    t.end();
    return BPK_OK;
}

// the same on one sub-coset of `m` points: witness rows a b c z PI z(wX), `period` values of 1 / Z_H
int plonk_quotient_evals_shard(bpk_ctx* ctx, const fr_t* wv, const fr_t* cv, size_t m, uint32_t period, const fr_t& beta,
                               const fr_t& gamma, const fr_t& alpha, const fr_t& k1, const fr_t& k2, const fr_t* zh_inv_host,
                               fr_t* out) {
    if (m == 0 || period == 0 || period > 64 || (period & (period - 1))) return BPK_ERR_INVALID_ARG;
    StageTimer t(ctx, "plonk.quotient");
    fr_t* d_zh;
    BPK_TRY(ws_reserve(ctx, 12, 64 * sizeof(fr_t), (void**)&d_zh));
    BPK_CUDA(cudaMemcpyAsync(d_zh, zh_inv_host, period * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    QuotientParams P{beta, gamma, alpha, mul(alpha, alpha), k1, k2};
    plonk_quotient_kernel<<<grid_for(ctx, m, 256), 256, 0, ctx->stream>>>(wv, cv, m, period, 1, d_zh, P, out);
    count_launch(ctx);
    BPK_CUDA(cudaGetLastError());
    t.end();
    return BPK_OK;
}

// out[r][k] = sum_q in[r][k + q m] * s^q  (k < m, k + q m < len): the coefficients of p(X) mod (X^m - s), i.e. of the
// polynomial that agrees with p wherever X^m = s -- on a coset of the m-th roots of unity.  rows stored `stride` apart.
__global__ void __launch_bounds__(256) fr_fold_kernel(const fr_t* __restrict__ in, size_t rows, size_t stride, size_t len,
                                                       size_t m, fr_t s, fr_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t step = (size_t)gridDim.x * blockDim.x;
    for (; i < rows * m; i += step) {
        const size_t r = i / m, k = i - r * m;
        const fr_t* src = in + r * stride;
        if (k >= len) {
            pst(out + i, fr_t::zero());
            continue;
        }
        size_t top = k;
        while (top + m < len) top += m;  // Horner from the highest chunk down
        fr_t acc = pld(src + top);
        while (top >= m + k) {
            top -= m;
            acc = add(mul(acc, s), pld(src + top));
        }
        pst(out + i, acc);
    }
}

int fr_fold(bpk_ctx* ctx, const fr_t* in, size_t rows, size_t stride, size_t len, size_t m, const fr_t& s, fr_t* out) {
    if (rows == 0 || m == 0) return BPK_OK;
    StageTimer t(ctx, "fr.fold");
    fr_fold_kernel<<<grid_for(ctx, rows * m, 256), 256, 0, ctx->stream>>>(in, rows, stride, len, m, s, out);
    count_launch(ctx);
    BPK_CUDA(cudaGetLastError());
    t.end();
    return BPK_OK;
}

}  // namespace bpk
