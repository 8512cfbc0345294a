// SRS handling on the device: G1Projective <-> affine conversion and Setup::generate_srs.
//
// Reference: Setup { powers_of_x: Vec<G1Projective> } (src/setup.rs:7-10); generate_srs walks
// cur *= tau serially, 255 doublings + adds per power (setup.rs:24-28, g1.rs:754-774).  Here every
// power [tau^i]G is computed independently from an 8-bit fixed-base comb table of the generator.
#include "internal.cuh"

namespace bpk {

__device__ __forceinline__ fp_t ld_fp_s(const fp_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1], c = q[2];
    fp_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    r.l[8] = c.x; r.l[9] = c.y; r.l[10] = c.z; r.l[11] = c.w;
    return r;
}
__device__ __forceinline__ void st_fp_s(fp_t* p, const fp_t& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    q[2] = make_uint4(v.l[8], v.l[9], v.l[10], v.l[11]);
}

// G1Affine::from(&G1Projective) (g1.rs:49-63) for every SRS point; identity -> (0, 0)
__global__ void __launch_bounds__(128) srs_from_projective_kernel(const uint64_t* __restrict__ xyz, size_t n,
                                                                   affine_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fp_t* p = reinterpret_cast<const fp_t*>(xyz + 18 * i);
    fp_t X = ld_fp_s(p), Y = ld_fp_s(p + 1), Z = ld_fp_s(p + 2);
    affine_t a;
    if (Z.is_zero()) {
        a = affine_t::inf();
    } else if (Z == fp_t::one()) {
        a.x = X;
        a.y = Y;
    } else {
        a = proj_to_affine(X, Y, Z);
    }
    st_fp_s(&out[i].x, a.x);
    st_fp_s(&out[i].y, a.y);
}

// affine -> normalised G1Projective limbs (x, y, 1) / identity (0, 1, 0)  (g1.rs:468-476, 605-611)
__global__ void srs_to_projective_kernel(const affine_t* __restrict__ pts, size_t n, uint64_t* __restrict__ xyz) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fp_t x = ld_fp_s(&pts[i].x), y = ld_fp_s(&pts[i].y);
    fp_t* o = reinterpret_cast<fp_t*>(xyz + 18 * i);
    if (x.is_zero() && y.is_zero()) {
        st_fp_s(o, fp_t::zero());
        st_fp_s(o + 1, fp_t::one());
        st_fp_s(o + 2, fp_t::zero());
    } else {
        st_fp_s(o, x);
        st_fp_s(o + 1, y);
        st_fp_s(o + 2, fp_t::one());
    }
}

// g1.rs:199-214: generator, Montgomery limbs
__device__ __constant__ uint64_t G1_GEN_X[6] = {0x5cb38790fd530c16ull, 0x7817fc679976fff5ull, 0x154f95c7143ba1c1ull,
                                                0xf0ae6acdf3d0e747ull, 0xedce6ecc21dbf440ull, 0x120177419e0bfb75ull};
__device__ __constant__ uint64_t G1_GEN_Y[6] = {0xbaac93d50ce72271ull, 0x8c22631a7918fd8eull, 0xdd595f13570725ceull,
                                                0x51ac582950405194ull, 0x0e1c8c3fad0059c0ull, 0x0bbc3efc5008a26aull};

// comb table T[k][d] = [d * 2^(8k)] G, k < 32, d < 256 (T[k][0] = infinity)
__global__ void gen_table_rows_kernel(xyzz_t* table) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 32) return;
    xyzz_t B;
    for (int i = 0; i < 6; i++) {
        B.X.l[2 * i] = (uint32_t)G1_GEN_X[i];
        B.X.l[2 * i + 1] = (uint32_t)(G1_GEN_X[i] >> 32);
        B.Y.l[2 * i] = (uint32_t)G1_GEN_Y[i];
        B.Y.l[2 * i + 1] = (uint32_t)(G1_GEN_Y[i] >> 32);
    }
    B.ZZ = fp_t::one();
    B.ZZZ = fp_t::one();
    for (uint32_t i = 0; i < 8 * k; i++) xyzz_dbl(B);
    xyzz_t acc = xyzz_t::inf();
    table[k * 256] = acc;
    for (uint32_t d = 1; d < 256; d++) {
        xyzz_add(acc, B);
        table[k * 256 + d] = acc;
    }
}

__global__ void xyzz_to_affine_kernel(const xyzz_t* in, size_t n, affine_t* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xyzz_t p = in[i];
    affine_t a = xyzz_to_affine(p);
    st_fp_s(&out[i].x, a.x);
    st_fp_s(&out[i].y, a.y);
}

// out[i] = [tau^i] G
__global__ void __launch_bounds__(128) srs_generate_kernel(const affine_t* __restrict__ table, fr_t tau, size_t first,
                                                            size_t n, affine_t* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t e = from_mont(pow_u64(tau, (uint64_t)(first + i)));  // canonical tau^(first + i)
    xyzz_t acc = xyzz_t::inf();
#pragma unroll 1
    for (uint32_t k = 0; k < 32; k++) {
        uint32_t limb = 0;
#pragma unroll
        for (int j = 0; j < 8; j++)
            if ((k >> 2) == (uint32_t)j) limb = e.l[j];
        uint32_t d = (limb >> (8 * (k & 3))) & 0xffu;
        if (d) {
            affine_t q;
            q.x = ld_fp_s(&table[k * 256 + d].x);
            q.y = ld_fp_s(&table[k * 256 + d].y);
            xyzz_madd(acc, q);
        }
    }
    affine_t a = xyzz_to_affine(acc);
    st_fp_s(&out[i].x, a.x);
    st_fp_s(&out[i].y, a.y);
}

// out[i] = [2^c] in[i]: c doublings in XYZZ, back to affine.  Each thread owns PRE_BATCH consecutive
// points and shares one field inversion among them (Montgomery's trick: invert the product of the
// ZZ*ZZZ denominators, peel the individual inverses off backwards).
constexpr int PRE_BATCH = 2;  // 6 x PRE_BATCH field elements stay in registers
__global__ void __launch_bounds__(128) srs_precompute_level_kernel(const affine_t* __restrict__ in,
                                                                    affine_t* __restrict__ out, size_t n, uint32_t c) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t i0 = t * PRE_BATCH;
    if (i0 >= n) return;
    fp_t X[PRE_BATCH], Y[PRE_BATCH], D[PRE_BATCH], Zz[PRE_BATCH], Zzz[PRE_BATCH], pre[PRE_BATCH];
    bool inf[PRE_BATCH];
    fp_t run = fp_t::one();
#pragma unroll
    for (int k = 0; k < PRE_BATCH; k++) {
        inf[k] = true;
        if (i0 + k < n) {
            affine_t a;
            a.x = ld_fp_s(&in[i0 + k].x);
            a.y = ld_fp_s(&in[i0 + k].y);
            xyzz_t p = xyzz_t::from_affine(a);
#pragma unroll 1
            for (uint32_t d = 0; d < c; d++) xyzz_dbl(p);
            inf[k] = p.is_inf();
            X[k] = p.X;
            Y[k] = p.Y;
            Zz[k] = p.ZZ;
            Zzz[k] = p.ZZZ;
        }
        D[k] = inf[k] ? fp_t::one() : mul(Zz[k], Zzz[k]);
        pre[k] = run;  // product of the denominators before k
        run = mul(run, D[k]);
    }
    fp_t iv = inv(run);
#pragma unroll
    for (int k = PRE_BATCH - 1; k >= 0; k--) {
        fp_t dk_inv = mul(iv, pre[k]);  // 1 / D[k]
        iv = mul(iv, D[k]);
        if (i0 + k < n) {
            affine_t a = affine_t::inf();
            if (!inf[k]) {  // x = X / ZZ = X * ZZZ / D, y = Y / ZZZ = Y * ZZ / D
                a.x = mul(X[k], mul(dk_inv, Zzz[k]));
                a.y = mul(Y[k], mul(dk_inv, Zz[k]));
            }
            st_fp_s(&out[i0 + k].x, a.x);
            st_fp_s(&out[i0 + k].y, a.y);
        }
    }
}

// ---------------------------------------------------------------------------------------------
int srs_precompute_level(bpk_ctx* ctx, const affine_t* d_in, affine_t* d_out, size_t n, uint32_t c) {
    if (n == 0) return BPK_OK;
    StageTimer t(ctx, "srs.precompute");
    size_t threads = (n + PRE_BATCH - 1) / PRE_BATCH;
    srs_precompute_level_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(d_in, d_out, n, c);
    count_launch(ctx);
    BPK_CUDA(cudaGetLastError());
    t.end();
    return BPK_OK;
}

int srs_from_projective(bpk_ctx* ctx, const uint64_t* d_xyz, size_t n, affine_t* d_out) {
    if (n == 0) return BPK_OK;
    StageTimer t(ctx, "srs.from_projective");
    srs_from_projective_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d_xyz, n, d_out);
    count_launch(ctx);
    BPK_CUDA(cudaGetLastError());
    t.end();
    return BPK_OK;
}

int srs_to_projective(bpk_ctx* ctx, const affine_t* d_pts, size_t n, uint64_t* d_xyz) {
    if (n == 0) return BPK_OK;
    srs_to_projective_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(d_pts, n, d_xyz);
    count_launch(ctx);
    BPK_CUDA(cudaGetLastError());
    return BPK_OK;
}

int srs_generate(bpk_ctx* ctx, const fr_t& tau, size_t first, size_t n, affine_t* d_out) {
    if (ctx->gen_table == nullptr) {
        StageTimer t(ctx, "srs.gen_table");
        xyzz_t* tmp;
        BPK_TRY(ws_reserve(ctx, 7, 32 * 256 * sizeof(xyzz_t), (void**)&tmp));
        affine_t* tab;
        BPK_CUDA(cudaMalloc(&tab, 32 * 256 * sizeof(affine_t)));
        gen_table_rows_kernel<<<1, 32, 0, ctx->stream>>>(tmp);
        count_launch(ctx);
        BPK_CUDA(cudaGetLastError());
        xyzz_to_affine_kernel<<<64, 128, 0, ctx->stream>>>(tmp, 32 * 256, tab);
        count_launch(ctx);
        BPK_CUDA(cudaGetLastError());
        ctx->gen_table = tab;
        t.end();
    }
    if (n == 0) return BPK_OK;
    StageTimer t(ctx, "srs.generate");
    srs_generate_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(ctx->gen_table, tau, first, n, d_out);
    count_launch(ctx);
    BPK_CUDA(cudaGetLastError());
    t.end();
    return BPK_OK;
}

}  // namespace bpk
