// BLS12-381 G1 multi-scalar multiplication (Pippenger) for sm_100a.
//
// Contract = BucketMSM::bucket_msm (src/msm.rs:76-118): sum_i (s_i >> rshift) * P_i, judged on the
// affine image of the result (G1Affine::from, g1.rs:49-63).  The reference walks 64 4-bit windows
// serially with complete projective additions; nothing of that structure is kept:
//
//   K1 recode      scalar: Montgomery -> canonical (scalar.rs:292-304 semantics), signed c-bit digits;
//                  one (bucket key, point index | sign) pair per (scalar, window)
//   K2 sort        radix sort of the pairs by bucket key  (replaces the serial scatter, msm.rs:29-35)
//   K3 accumulate  equal-size chunks of the sorted pair list, one thread per chunk, XYZZ += affine;
//                  a bucket whose run lies inside one chunk is written directly, runs that cross chunk
//                  borders leave partial sums that K3b merges -> load balance is independent of the
//                  scalar distribution
//   K4 reduce      sum_b (b+1) B_b per window by a tree of running sums (msm.rs:42-46 is the serial form)
//   K5 finalize    Horner over windows (msm.rs:107-115), affine normalisation, G1Projective limbs out
//
// Roofline: integer multiply pipe.  Algorithmic work per accumulated pair: one mixed addition
// = 8 M + 2 S in Fp, each 300 32x32->64 products; HBM traffic per pair is 8 B of
// sorted pair + 96 B point gather -- two orders of magnitude below the arithmetic time.
#include <cub/device/device_radix_sort.cuh>

// One shared out-of-line body for the Fp product: the accumulate loop otherwise inlines ten ~450-instruction
// multiplications (~75 KB of SASS) and stalls on instruction fetch (14 % "no_instructions" samples in
// profiles/r1_final_msm_accumulate_ncu.md); measured 79.9 -> 77.3 ms at 2^24.
#define BPK_FP_MUL_CALL 1
#include "internal.cuh"

namespace bpk {

static constexpr uint32_t INVALID_KEY = 0xffffffffu;

__device__ __forceinline__ fp_t ld_fp(const fp_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1], c = q[2];
    fp_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    r.l[8] = c.x; r.l[9] = c.y; r.l[10] = c.z; r.l[11] = c.w;
    return r;
}
__device__ __forceinline__ void st_fp(fp_t* p, const fp_t& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    q[2] = make_uint4(v.l[8], v.l[9], v.l[10], v.l[11]);
}
__device__ __forceinline__ affine_t ld_affine(const affine_t* p) {
    affine_t r;
    r.x = ld_fp(&p->x);
    r.y = ld_fp(&p->y);
    return r;
}
__device__ __forceinline__ xyzz_t ld_xyzz(const xyzz_t* p) {
    xyzz_t r;
    r.X = ld_fp(&p->X);
    r.Y = ld_fp(&p->Y);
    r.ZZ = ld_fp(&p->ZZ);
    r.ZZZ = ld_fp(&p->ZZZ);
    return r;
}
__device__ __forceinline__ void st_xyzz(xyzz_t* p, const xyzz_t& v) {
    st_fp(&p->X, v.X);
    st_fp(&p->Y, v.Y);
    st_fp(&p->ZZ, v.ZZ);
    st_fp(&p->ZZZ, v.ZZZ);
}
__device__ __forceinline__ fr_t ld_fr_g(const fr_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    fr_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}

// ------------------------------------------------------------------------------------------
// K1: signed-digit recoding
// ------------------------------------------------------------------------------------------
// key_stride: buckets per window (2^(c-1)) for per-window bucket sets, 0 when every window shares one
// bucket set (precomputed SRS levels); val_stride: distance between precomputed levels in points, else 0.
__global__ void __launch_bounds__(256) msm_recode_kernel(const fr_t* __restrict__ scalars, uint32_t n, uint32_t c,
                                                          uint32_t W, uint32_t rshift, uint32_t key_stride,
                                                          uint32_t val_stride, uint32_t* __restrict__ keys,
                                                          uint32_t* __restrict__ vals) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t s = from_mont(ld_fr_g(scalars + i));  // canonical integer < q
    uint32_t l[10];
#pragma unroll
    for (int k = 0; k < 8; k++) l[k] = s.l[k];
    l[8] = 0;
    l[9] = 0;
    if (rshift) {  // value >> rshift (the c-does-not-divide-256 quirk of msm.rs:119-139)
        uint32_t ws = rshift >> 5, bs = rshift & 31;
        uint32_t t[10];
#pragma unroll
        for (int k = 0; k < 10; k++) {
            uint32_t lo = (k + ws < 8) ? l[k + ws] : 0;
            uint32_t hi = (k + ws + 1 < 8) ? l[k + ws + 1] : 0;
            t[k] = bs ? ((lo >> bs) | (hi << (32 - bs))) : lo;
        }
#pragma unroll
        for (int k = 0; k < 10; k++) l[k] = t[k];
    }
    const uint32_t half = 1u << (c - 1);
    const uint32_t mask = (1u << c) - 1;
    uint32_t carry = 0;
    for (uint32_t w = 0; w < W; w++) {
        uint32_t bit = w * c;
        uint32_t limb = bit >> 5, off = bit & 31;
        uint64_t two = 0;
        if (limb < 9) two = (uint64_t)l[limb] | ((uint64_t)l[limb + 1] << 32);
        uint32_t raw = ((uint32_t)(two >> off) & mask) + carry;
        uint32_t d, neg;
        if (raw > half) {
            d = (1u << c) - raw;
            neg = 1;
            carry = 1;
        } else {
            d = raw;
            neg = 0;
            carry = 0;
        }
        size_t o = (size_t)w * n + i;
        keys[o] = d ? (w * key_stride + d - 1) : INVALID_KEY;
        vals[o] = (w * val_stride + i) | (neg << 31);
    }
}

// ------------------------------------------------------------------------------------------
// K3: chunked bucket accumulation
// ------------------------------------------------------------------------------------------
#ifndef BPK_ACC_MINBLOCKS
#define BPK_ACC_MINBLOCKS 4  // 4 x 128 threads / SM = 128 registers per thread: ~0.5 KB of spills, but 16 warps hide the
                             // dependent-issue waits better than 12 (2^24: 74.6 -> 73.2 ms; 2 CTAs: 77.0, 5 CTAs: 77.4)
#endif
__global__ void __launch_bounds__(128, BPK_ACC_MINBLOCKS) msm_accumulate_kernel(const uint32_t* __restrict__ keys,
                                                              const uint32_t* __restrict__ vals, size_t M,
                                                              uint32_t chunk, size_t num_chunks,
                                                              const affine_t* __restrict__ points, uint32_t nb_total,
                                                              xyzz_t* __restrict__ buckets,
                                                              uint32_t* __restrict__ pkeys,
                                                              xyzz_t* __restrict__ pvals) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_chunks) return;
    const size_t a = t * chunk;
    const size_t b = (a + chunk < M) ? a + chunk : M;
    uint32_t pk0 = INVALID_KEY, pk1 = INVALID_KEY;

    uint32_t cur = INVALID_KEY;
    bool left_open = false;
    xyzz_t acc = xyzz_t::inf();

    // software pipeline: the point of entry e+1 is fetched while entry e is added, and the (key, value) pair of
    // entry e+2 is already in registers, so the gather never waits behind the two strided index loads
    uint32_t k_next = keys[a], v_next = vals[a];
    uint32_t k_next2 = INVALID_KEY, v_next2 = 0;
    if (a + 1 < b) {
        k_next2 = keys[a + 1];
        v_next2 = vals[a + 1];
    }
    affine_t p_next = affine_t::inf();
    if (k_next < nb_total) {
        p_next = ld_affine(points + (v_next & 0x7fffffffu));
        if (v_next >> 31) p_next.y = neg(p_next.y);
    }
    size_t e = a;
    for (; e < b; e++) {
        uint32_t k = k_next;
        if (k >= nb_total) break;
        affine_t p = p_next;
        k_next = k_next2;
        v_next = v_next2;
        if (e + 2 < b) {
            k_next2 = keys[e + 2];
            v_next2 = vals[e + 2];
        } else {
            k_next2 = INVALID_KEY;
        }
        if (e + 1 < b && k_next < nb_total) {
            p_next = ld_affine(points + (v_next & 0x7fffffffu));
            if (v_next >> 31) p_next.y = neg(p_next.y);
        }
        if (k != cur) {
            if (cur != INVALID_KEY) {  // close the previous run (cannot be right-open)
                if (left_open) {
                    pk0 = cur;
                    st_xyzz(pvals + 2 * t, acc);
                } else {
                    st_xyzz(buckets + cur, acc);
                }
            }
            cur = k;
            left_open = (e == a) && (a > 0) && (keys[a - 1] == k);
            acc = xyzz_t::from_affine(p);
        } else {
            xyzz_madd(acc, p);
        }
    }
    if (cur != INVALID_KEY) {
        bool right_open = (e == b) && (b < M) && (keys[b] == cur);
        if (left_open) {
            pk0 = cur;
            st_xyzz(pvals + 2 * t, acc);
        } else if (right_open) {
            pk1 = cur;
            st_xyzz(pvals + 2 * t + 1, acc);
        } else {
            st_xyzz(buckets + cur, acc);
        }
    }
    pkeys[2 * t] = pk0;
    pkeys[2 * t + 1] = pk1;
}

// K3b: a run that crosses chunk borders starts as slot 1 of some chunk t0 (the head) and continues as slot 0
// of the chunks t0+1 .. t1 whose first pair has the same key.  The thread that owns the head finds t1 by a
// binary search over the first keys of the chunks (sorted), adds short runs itself and queues long runs
// (heavy buckets: skewed scalars, narrow top windows) for a block-wide tree reduction, so that no scalar
// distribution can serialise the merge.
constexpr uint32_t MERGE_SERIAL_MAX = 8;    // runs up to this many partials: added by the head's thread
constexpr uint32_t MERGE_WARP_MAX = 256;    // up to this many: one warp per run; longer: one block per run
struct LongRun {
    uint32_t key, t0, t1, pad;
};

__global__ void __launch_bounds__(128, 3) msm_merge_partials_kernel(const uint32_t* __restrict__ pkeys,
                                                                  const xyzz_t* __restrict__ pvals,
                                                                  size_t num_chunks, xyzz_t* __restrict__ buckets,
                                                                  const uint32_t* __restrict__ keys, uint32_t chunk,
                                                                  LongRun* __restrict__ long_runs,
                                                                  uint32_t* __restrict__ long_count,
                                                                  LongRun* __restrict__ warp_runs,
                                                                  uint32_t* __restrict__ warp_count) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_chunks) return;
    uint32_t k = pkeys[2 * t + 1];
    if (k == INVALID_KEY) return;
    // chunk t+1 starts with key k (the head is right-open); find the last chunk that does
    size_t lo = t + 1, hi = num_chunks;
    while (lo + 1 < hi) {
        size_t mid = lo + (hi - lo) / 2;
        if (keys[mid * chunk] <= k)
            lo = mid;
        else
            hi = mid;
    }
    const size_t t1 = lo;
    if (t1 - t > MERGE_SERIAL_MAX) {
        LongRun r;
        r.key = k;
        r.t0 = (uint32_t)t;
        r.t1 = (uint32_t)t1;
        r.pad = 0;
        if (t1 - t > MERGE_WARP_MAX)
            long_runs[atomicAdd(long_count, 1u)] = r;
        else
            warp_runs[atomicAdd(warp_count, 1u)] = r;
        return;
    }
    xyzz_t acc = ld_xyzz(pvals + 2 * t + 1);
    for (size_t u = t + 1; u <= t1; u++) {
        xyzz_t q = ld_xyzz(pvals + 2 * u);
        xyzz_add(acc, q);
    }
    st_xyzz(buckets + k, acc);
}

__device__ __forceinline__ fp_t shfl_down_fp(const fp_t& v, int delta) {
    fp_t r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = __shfl_down_sync(0xffffffffu, v.l[i], delta);
    return r;
}

// one warp per medium run: lane-strided partial sums, then a shuffle tree
__global__ void __launch_bounds__(128) msm_merge_warp_runs_kernel(const LongRun* __restrict__ runs,
                                                                   const uint32_t* __restrict__ run_count,
                                                                   const xyzz_t* __restrict__ pvals,
                                                                   xyzz_t* __restrict__ buckets) {
    const uint32_t count = *run_count;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = warp; i < count; i += nwarps) {
        const LongRun r = runs[i];
        xyzz_t acc = xyzz_t::inf();
        if (lane == 0) acc = ld_xyzz(pvals + 2 * (size_t)r.t0 + 1);  // the head partial
        for (size_t u = (size_t)r.t0 + 1 + lane; u <= r.t1; u += 32) {
            xyzz_t q = ld_xyzz(pvals + 2 * u);
            xyzz_add(acc, q);
        }
        __syncwarp();
#pragma unroll 1
        for (int delta = 16; delta >= 1; delta >>= 1) {
            xyzz_t o;
            o.X = shfl_down_fp(acc.X, delta);
            o.Y = shfl_down_fp(acc.Y, delta);
            o.ZZ = shfl_down_fp(acc.ZZ, delta);
            o.ZZZ = shfl_down_fp(acc.ZZZ, delta);
            xyzz_add(acc, o);
            __syncwarp();
        }
        if (lane == 0) st_xyzz(buckets + r.key, acc);
    }
}

// one block per long run: strided partial sums, then a shared-memory tree
constexpr int MERGE_BLOCK = 256;
__global__ void __launch_bounds__(MERGE_BLOCK) msm_merge_long_runs_kernel(const LongRun* __restrict__ long_runs,
                                                                           const uint32_t* __restrict__ long_count,
                                                                           const xyzz_t* __restrict__ pvals,
                                                                           xyzz_t* __restrict__ buckets) {
    __shared__ xyzz_t sm[MERGE_BLOCK];
    const uint32_t count = *long_count;
    const uint32_t tid = threadIdx.x;
    for (uint32_t i = blockIdx.x; i < count; i += gridDim.x) {
        const LongRun r = long_runs[i];
        xyzz_t acc = xyzz_t::inf();
        if (tid == 0) acc = ld_xyzz(pvals + 2 * (size_t)r.t0 + 1);  // the head partial
        for (size_t u = (size_t)r.t0 + 1 + tid; u <= r.t1; u += MERGE_BLOCK) {
            xyzz_t q = ld_xyzz(pvals + 2 * u);
            xyzz_add(acc, q);
        }
        sm[tid] = acc;
        __syncthreads();
        for (uint32_t s = MERGE_BLOCK / 2; s >= 1; s >>= 1) {
            if (tid < s) {
                xyzz_t a = sm[tid];
                xyzz_t b = sm[tid + s];
                xyzz_add(a, b);
                sm[tid] = a;
            }
            __syncthreads();
        }
        if (tid == 0) st_xyzz(buckets + r.key, sm[0]);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// K4: bucket reduction tree.  Level input: per window `items` points A (and carried sums V);
// each thread folds S consecutive items:  A' = sum A_r,  V' = 2^dbls * sum_r r A_r + sum_r V_r.
// Invariant: sum_b b A0_b = S^level * sum_k k A'_k + sum_k V'_k.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) msm_reduce_level_kernel(const xyzz_t* __restrict__ A,
                                                                const xyzz_t* __restrict__ V,
                                                                xyzz_t* __restrict__ A2, xyzz_t* __restrict__ V2,
                                                                uint32_t items, uint32_t S, uint32_t W,
                                                                uint32_t dbls) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t groups = items / S;
    if (t >= groups * W) return;
    uint32_t w = t / groups, k = t % groups;
    size_t base = (size_t)w * items + (size_t)k * S;
    xyzz_t run = xyzz_t::inf(), acc = xyzz_t::inf();
    for (uint32_t r = S - 1; r >= 1; r--) {
        xyzz_t q = ld_xyzz(A + base + r);
        xyzz_add(run, q);
        xyzz_add(acc, run);
    }
    {
        xyzz_t q = ld_xyzz(A + base);
        xyzz_add(run, q);
    }
    for (uint32_t i = 0; i < dbls; i++) xyzz_dbl(acc);
    if (V) {
        for (uint32_t r = 0; r < S; r++) {
            xyzz_t q = ld_xyzz(V + base + r);
            xyzz_add(acc, q);
        }
    }
    st_xyzz(A2 + (size_t)w * groups + k, run);
    st_xyzz(V2 + (size_t)w * groups + k, acc);
}

// ------------------------------------------------------------------------------------------
// K4 (default): bucket reduction by bit planes.  With B_k the bucket of digit k + 1 (k < half = 2^L),
//     sum_k (k + 1) B_k = sum_k B_k + sum_{p < L} 2^p T_p,      T_p = sum_{k : bit p of k set} B_k.
// One pairwise tree carries the plane sums along: a level-k node covers 2^k consecutive buckets and holds
// k + 1 points [S, T_0 .. T_{k-1}] restricted to its range.  Merging the children (c0, c1) of a node:
//     S' = S(c0) + S(c1),   T_p' = T_p(c0) + T_p(c1) for p < k - 1,   T_{k-1}' = S(c1)
// (the upper child is exactly the set of buckets whose bit k - 1 is set).  One thread per (node, slot): every
// level is one launch of independent additions -- 2 unweighted additions per bucket in total, no doublings, no
// running-sum chains; the wide levels run at arithmetic throughput, each narrow level costs one addition of
// latency.  2^p is applied once, in the final Horner pass over the root's L plane sums.
// ------------------------------------------------------------------------------------------
#ifndef BPK_TREE_MINBLOCKS
#define BPK_TREE_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(128, BPK_TREE_MINBLOCKS) msm_plane_tree_level_kernel(const xyzz_t* __restrict__ in,
                                                                    xyzz_t* __restrict__ out, uint32_t k,
                                                                    size_t nodes_out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t slots = k + 1;
    if (t >= nodes_out * slots) return;
    const size_t node = t / slots;
    const uint32_t s = (uint32_t)(t - node * slots);
    const xyzz_t* c0 = in + 2 * node * k;  // children hold k points each
    const xyzz_t* c1 = c0 + k;
    if (s == k) {
        st_xyzz(out + t, ld_xyzz(c1));
        return;
    }
    xyzz_t a = ld_xyzz(c0 + s);
    xyzz_t b = ld_xyzz(c1 + s);
    xyzz_add(a, b);
    st_xyzz(out + t, a);
}

// ------------------------------------------------------------------------------------------
// K5: window Horner + output
// ------------------------------------------------------------------------------------------
__device__ void write_projective(uint64_t* out, const xyzz_t& p, bool normalise) {
    fp_t X, Y, Z;
    if (p.is_inf()) {  // identity (0, 1, 0)  g1.rs:605-611
        X = fp_t::zero();
        Y = fp_t::one();
        Z = fp_t::zero();
    } else if (normalise) {
        affine_t a = xyzz_to_affine(p);
        X = a.x;
        Y = a.y;
        Z = fp_t::one();
    } else {  // x = X/ZZ, y = Y/ZZZ  ->  (X ZZZ : Y ZZ : ZZ ZZZ)
        X = mul(p.X, p.ZZZ);
        Y = mul(p.Y, p.ZZ);
        Z = mul(p.ZZ, p.ZZZ);
    }
    uint32_t* o = reinterpret_cast<uint32_t*>(out);
    for (int i = 0; i < 12; i++) {
        o[i] = X.l[i];
        o[12 + i] = Y.l[i];
        o[24 + i] = Z.l[i];
    }
}

__global__ void msm_finalize_kernel(const xyzz_t* A, const xyzz_t* V, uint32_t W, uint32_t c, int normalise,
                                    uint64_t* out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    xyzz_t acc = xyzz_t::inf();
    for (int w = (int)W - 1; w >= 0; w--) {
        for (uint32_t i = 0; i < c; i++) xyzz_dbl(acc);
        xyzz_t tw = ld_xyzz(V + w);  // sum_b b B_b
        xyzz_t g = ld_xyzz(A + w);   // sum_b B_b
        xyzz_add(tw, g);
        xyzz_add(acc, tw);
    }
    write_projective(out, acc, normalise != 0);
}

// plane form: the L plane sums of a window are folded by Horner in 2.  With few windows (precomputed levels:
// one) the planes are split into groups of GROUP consecutive planes, one thread per group, and the group values
// are folded in 2^GROUP -- the same doublings, but a third of the additions on the serial path.
constexpr uint32_t FIN_GROUP = 4;
__global__ void __launch_bounds__(128) msm_finalize_planes_kernel(const xyzz_t* __restrict__ roots, uint32_t W,
                                                                   uint32_t L, uint32_t c, uint32_t groups,
                                                                   int normalise, uint64_t* out) {
    // roots: per window L + 1 points [S, T_0 .. T_{L-1}] (the root of the plane tree)
    const xyzz_t* plane_sums = roots + 1;
    const xyzz_t* totals = roots;
    const uint32_t stride = L + 1;
    __shared__ xyzz_t part[128];
    const uint32_t t = threadIdx.x;
    const uint32_t per = groups > 1 ? FIN_GROUP : L;  // planes per thread
    if (t < W * groups) {
        const uint32_t w = t / groups, g = t % groups;
        const int lo = (int)(g * per);
        int hi = lo + (int)per;
        if (hi > (int)L) hi = (int)L;
        xyzz_t acc = xyzz_t::inf();
        for (int p = hi - 1; p >= lo; p--) {
            xyzz_dbl(acc);
            xyzz_t v = ld_xyzz(plane_sums + (size_t)w * stride + p);
            xyzz_add(acc, v);
        }
        part[t] = acc;
    }
    __syncthreads();
    if (t < W) {  // fold the groups of window t, add the plain bucket sum
        xyzz_t acc = part[t * groups + groups - 1];
        for (int g = (int)groups - 2; g >= 0; g--) {
            for (uint32_t i = 0; i < FIN_GROUP; i++) xyzz_dbl(acc);
            xyzz_t v = part[t * groups + g];
            xyzz_add(acc, v);
        }
        xyzz_t v = ld_xyzz(totals + (size_t)t * stride);
        xyzz_add(acc, v);
        part[t * groups] = acc;  // only this thread reads or writes the slots of window t in this phase
    }
    __syncthreads();
    if (t != 0) return;
    xyzz_t acc = part[(W - 1) * groups];
    for (int k = (int)W - 2; k >= 0; k--) {
        for (uint32_t i = 0; i < c; i++) xyzz_dbl(acc);
        xyzz_t v = part[(size_t)k * groups];
        xyzz_add(acc, v);
    }
    write_projective(out, acc, normalise != 0);
}

// sum of n homogeneous projective points (X:Y:Z), normalised output
__global__ void g1_sum_kernel(const uint64_t* pts, uint32_t n, uint64_t* out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    xyzz_t acc = xyzz_t::inf();
    for (uint32_t i = 0; i < n; i++) {
        const fp_t* p = reinterpret_cast<const fp_t*>(pts + 18 * (size_t)i);
        fp_t X = ld_fp(p), Y = ld_fp(p + 1), Z = ld_fp(p + 2);
        if (Z.is_zero()) continue;
        xyzz_t q;  // x = X/Z = XZ/Z^2, y = Y/Z = YZ^2/Z^3
        q.ZZ = sqr(Z);
        q.ZZZ = mul(q.ZZ, Z);
        q.X = mul(X, Z);
        q.Y = mul(Y, q.ZZ);
        xyzz_add(acc, q);
    }
    write_projective(out, acc, true);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static uint32_t pick_window(bpk_ctx* ctx, size_t n) {
    if (ctx->opt_msm_window >= 2 && ctx->opt_msm_window <= 22) return (uint32_t)ctx->opt_msm_window;
    double best = 1e300;
    uint32_t best_c = 4;
    for (uint32_t c = 3; c <= 16; c++) {
        uint32_t W = (256 + c - 1) / c;
        // pair additions + bucket-tree additions (full adds ~1.5x a mixed add, 3 per bucket) + a latency
        // term for the serial depth of the tree / Horner tail expressed in pair-addition equivalents
        double cost = (double)W * ((double)n + 4.5 * (double)(1u << (c - 1))) + 2000.0 * c;
        if (cost < best) {
            best = cost;
            best_c = c;
        }
    }
    return best_c;
}

// ---- plan: window geometry shared by the phases of one MSM ----
struct MsmPlan {
    bool pre;        // all windows share one bucket set (precomputed SRS levels)
    uint32_t c, W, half, WB, nb_total;
};

static int msm_make_plan(bpk_ctx* ctx, const MsmPoints& pts, size_t n, MsmPlan* plan) {
    if (n >= ((size_t)1 << 31)) return BPK_ERR_TOO_LARGE;
    plan->pre = pts.pre_c != 0;
    plan->c = plan->pre ? pts.pre_c : pick_window(ctx, n);
    plan->W = plan->pre ? pts.pre_W : (256 + plan->c - 1) / plan->c;
    plan->half = 1u << (plan->c - 1);
    plan->WB = plan->pre ? 1 : plan->W;
    plan->nb_total = plan->WB * plan->half;
    if ((size_t)plan->W * n >= ((size_t)1 << 31)) return BPK_ERR_TOO_LARGE;
    if (plan->pre && (size_t)plan->W * pts.level_stride >= ((size_t)1 << 31)) return BPK_ERR_TOO_LARGE;
    return BPK_OK;
}

static int msm_empty_result(bpk_ctx* ctx, uint64_t* d_out) {  // empty sum: identity (0, R, 0)
    xyzz_t* zero;
    BPK_TRY(ws_reserve(ctx, 5, 2 * sizeof(xyzz_t), (void**)&zero));
    BPK_CUDA(cudaMemsetAsync(zero, 0, 2 * sizeof(xyzz_t), ctx->stream));
    msm_finalize_kernel<<<1, 1, 0, ctx->stream>>>(zero, zero + 1, 1, 1, 1, d_out);
    count_launch(ctx);
    BPK_CUDA(cudaGetLastError());
    return BPK_OK;
}

static uint32_t msm_chunk(bpk_ctx* ctx, size_t M) {  // sorted pairs per accumulate thread
    uint32_t chunk = (uint32_t)ctx->opt_msm_chunk;
    if (chunk == 0) {
        // about 2048 threads per SM, at most 256 pairs each; then shrink the chunk so that the grid is a whole
        // number of waves of resident threads (threads take equal time, a nearly empty last wave costs a full one:
        // 2^20: 44 -> 36..45 pairs 7.41 -> 7.25 ms, 2^24: 256 -> 242 pairs 84.2 -> 83.6 ms)
        size_t target = M / ((size_t)ctx->sm_count * 2048);
        target = target < 16 ? 16 : (target > 256 ? 256 : target);
        const size_t resident = (size_t)ctx->sm_count * BPK_ACC_MINBLOCKS * 128;
        const size_t waves = (M + target * resident - 1) / (target * resident);
        size_t c = (M + waves * resident - 1) / (waves * resident);
        chunk = (uint32_t)(c < 16 ? 16 : c);
    }
    return chunk;
}

// grow the per-MSM workspaces for `n` pairs up front (growing later would synchronise the stream)
static int msm_reserve(bpk_ctx* ctx, const MsmPlan& pl, size_t n) {
    const size_t M = (size_t)pl.W * n;
    void* p;
    BPK_TRY(ws_reserve(ctx, 2, 4 * M * sizeof(uint32_t), &p));
    int end_bit = 1;
    while (((uint64_t)1 << end_bit) <= pl.nb_total) end_bit++;
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (uint32_t*)p, (uint32_t*)p, (uint32_t*)p, (uint32_t*)p, (int)M, 0,
                                    end_bit, ctx->stream);
    BPK_TRY(ws_reserve(ctx, 3, sort_bytes, &p));
    const uint32_t chunk = msm_chunk(ctx, M);
    const size_t num_chunks = (M + chunk - 1) / chunk;
    BPK_TRY(ws_reserve(ctx, 5, 2 * num_chunks * (sizeof(xyzz_t) + sizeof(uint32_t)), &p));
    BPK_TRY(ws_reserve(ctx, 11, 2 * (num_chunks / MERGE_SERIAL_MAX + 2) * sizeof(LongRun) + 16, &p));
    return BPK_OK;
}

// phase 1: recode, sort, accumulate, merge: `buckets` (nb_total entries) receives the bucket sums of the n pairs
static int msm_fill_buckets(bpk_ctx* ctx, const MsmPlan& pl, const MsmPoints& pts, const fr_t* d_scalars, size_t n,
                            unsigned rshift, xyzz_t* buckets) {
    const affine_t* d_points = pts.base;
    const bool pre = pl.pre;
    const uint32_t c = pl.c, W = pl.W, half = pl.half, nb_total = pl.nb_total;
    const size_t M = (size_t)W * n;

    // workspace carve-up
    uint32_t *keys_in, *vals_in, *keys_out, *vals_out;
    {
        void* base;
        BPK_TRY(ws_reserve(ctx, 2, 4 * M * sizeof(uint32_t), &base));
        keys_in = (uint32_t*)base;
        vals_in = keys_in + M;
        keys_out = vals_in + M;
        vals_out = keys_out + M;
    }
    int end_bit = 1;
    while (((uint64_t)1 << end_bit) <= nb_total) end_bit++;  // INVALID keys only need to sort last
    // INVALID_KEY = 0xffffffff has all low bits set, so within end_bit bits it is >= every valid key
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys_in, keys_out, vals_in, vals_out, (int)M, 0, end_bit,
                                    ctx->stream);
    void* sort_tmp;
    BPK_TRY(ws_reserve(ctx, 3, sort_bytes, &sort_tmp));

    const uint32_t chunk = msm_chunk(ctx, M);
    const size_t num_chunks = (M + chunk - 1) / chunk;
    ctx->last_c = c;
    ctx->last_W = W;
    ctx->last_chunk = chunk;
    ctx->last_buckets = nb_total;

    uint32_t* pkeys;
    xyzz_t* pvals;
    {
        void* base;
        size_t pv_bytes = 2 * num_chunks * sizeof(xyzz_t);
        BPK_TRY(ws_reserve(ctx, 5, pv_bytes + 2 * num_chunks * sizeof(uint32_t), &base));
        pvals = (xyzz_t*)base;
        pkeys = (uint32_t*)((char*)base + pv_bytes);
    }

    {
        StageTimer t(ctx, "msm.recode");
        msm_recode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
            d_scalars, (uint32_t)n, c, W, rshift, pre ? 0u : half, pre ? (uint32_t)pts.level_stride : 0u, keys_in, vals_in);
        count_launch(ctx);
        BPK_CUDA(cudaGetLastError());
        t.end();
    }
    {
        StageTimer t(ctx, "msm.sort");
        cudaError_t e = cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, keys_in, keys_out, vals_in, vals_out,
                                                        (int)M, 0, end_bit, ctx->stream);
        BPK_CUDA(e);
        count_launch(ctx, 1 + (uint64_t)((end_bit + 7) / 8));
        t.end();
    }
    {
        StageTimer t(ctx, "msm.accumulate");
        BPK_CUDA(cudaMemsetAsync(buckets, 0, (size_t)nb_total * sizeof(xyzz_t), ctx->stream));
        msm_accumulate_kernel<<<(unsigned)((num_chunks + 127) / 128), 128, 0, ctx->stream>>>(
            keys_out, vals_out, M, chunk, num_chunks, d_points, nb_total, buckets, pkeys, pvals);
        count_launch(ctx);
        BPK_CUDA(cudaGetLastError());
        t.end();
    }
    {
        StageTimer t(ctx, "msm.merge");
        LongRun *long_runs, *warp_runs;
        uint32_t *long_count, *warp_count;
        {
            void* base;
            const size_t cap = num_chunks / MERGE_SERIAL_MAX + 2;  // a queued run covers > MERGE_SERIAL_MAX chunks
            BPK_TRY(ws_reserve(ctx, 11, 2 * cap * sizeof(LongRun) + 16, &base));
            long_count = (uint32_t*)base;
            warp_count = long_count + 1;
            long_runs = (LongRun*)((char*)base + 16);
            warp_runs = long_runs + cap;
        }
        BPK_CUDA(cudaMemsetAsync(long_count, 0, 2 * sizeof(uint32_t), ctx->stream));
        msm_merge_partials_kernel<<<(unsigned)((num_chunks + 127) / 128), 128, 0, ctx->stream>>>(
            pkeys, pvals, num_chunks, buckets, keys_out, chunk, long_runs, long_count, warp_runs, warp_count);
        msm_merge_warp_runs_kernel<<<(unsigned)ctx->sm_count * 4, 128, 0, ctx->stream>>>(warp_runs, warp_count, pvals,
                                                                                         buckets);
        msm_merge_long_runs_kernel<<<(unsigned)ctx->sm_count * 2, MERGE_BLOCK, 0, ctx->stream>>>(long_runs, long_count,
                                                                                              pvals, buckets);
        count_launch(ctx, 3);
        BPK_CUDA(cudaGetLastError());
        t.end();
    }
    return BPK_OK;
}

// buckets_a[i] += buckets_b[i]: joins the bucket sets of two slices of one MSM
__global__ void __launch_bounds__(128, 3) msm_add_buckets_kernel(xyzz_t* __restrict__ a, const xyzz_t* __restrict__ b,
                                                               size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    xyzz_t y = ld_xyzz(b + i);
    if (y.is_inf()) return;
    xyzz_t x = ld_xyzz(a + i);
    xyzz_add(x, y);
    st_xyzz(a + i, x);
}

// phase 2: bucket reduction + window Horner + output
static int msm_reduce_buckets(bpk_ctx* ctx, const MsmPlan& pl, xyzz_t* buckets, bool normalise, uint64_t* d_out) {
    const uint32_t c = pl.c, W = pl.W, half = pl.half, WB = pl.WB;
    if (ctx->opt_msm_reduce == 0) {
        // bit-plane reduction (see K4): one tree whose nodes carry the plane sums, then Horner over the planes
        const uint32_t L = c - 1;  // half == 1 << L
        // ping-pong level buffers: level k holds WB * (half >> k) * (k + 1) points, largest at k = 1 and k = 2
        const size_t buf_a = (size_t)WB * half;                    // odd levels  (k = 1: WB * half points)
        const size_t buf_b = (size_t)WB * (half / 4 + 1) * 3;      // even levels (k = 2: 3/4 WB * half points)
        xyzz_t* lvl;
        BPK_TRY(ws_reserve(ctx, 6, (buf_a + buf_b + 1) * sizeof(xyzz_t), (void**)&lvl));
        const xyzz_t* roots = buckets;  // L == 0: the single bucket of each window is its own root
        {
            StageTimer t(ctx, "msm.reduce");
            const xyzz_t* in = buckets;
            for (uint32_t k = 1; k <= L; k++) {
                xyzz_t* out = (k & 1) ? lvl : lvl + buf_a;
                const size_t nodes = (size_t)WB * (half >> k);
                const size_t threads = nodes * (k + 1);
                msm_plane_tree_level_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(in, out, k, nodes);
                count_launch(ctx);
                in = out;
            }
            roots = in;
            BPK_CUDA(cudaGetLastError());
            t.end();
        }
        {
            StageTimer t(ctx, "msm.finalize");
            uint32_t groups = (L + FIN_GROUP - 1) / FIN_GROUP;
            if (groups == 0 || WB * groups > 128) groups = 1;
            msm_finalize_planes_kernel<<<1, 128, 0, ctx->stream>>>(roots, WB, L, c, groups, normalise ? 1 : 0, d_out);
            count_launch(ctx);
            BPK_CUDA(cudaGetLastError());
            t.end();
        }
        return BPK_OK;
    }
    // reduction tree: fold `fanin` items per thread per level (serial depth 3 x fanin additions per level)
    const uint32_t W_h = W;  // (window count of the Horner pass is WB below)
    (void)W_h;
    uint32_t fanin = (uint32_t)ctx->opt_msm_fanin;
    if (fanin < 2 || fanin > 32 || (fanin & (fanin - 1))) fanin = 8;
    const xyzz_t* A = buckets;
    const xyzz_t* V = nullptr;
    {
        StageTimer t(ctx, "msm.reduce");
        xyzz_t* lvl;
        // level outputs: items/32 (+ /1024 + ...) per window, A and V each; 2 * nb_total/16 is ample
        size_t lvl_elems = (size_t)WB * (half + 64);
        BPK_TRY(ws_reserve(ctx, 6, 2 * lvl_elems * sizeof(xyzz_t), (void**)&lvl));
        uint32_t items = half;
        uint32_t dbls = 0;
        size_t off = 0;
        while (items > 1) {
            uint32_t S = items >= fanin ? fanin : items;
            uint32_t groups = items / S;
            xyzz_t* A2 = lvl + off;
            xyzz_t* V2 = lvl + off + (size_t)WB * groups;
            off += 2 * (size_t)WB * groups;
            uint32_t threads = WB * groups;
            msm_reduce_level_kernel<<<(threads + 127) / 128, 128, 0, ctx->stream>>>(A, V, A2, V2, items, S, WB, dbls);
            count_launch(ctx);
            BPK_CUDA(cudaGetLastError());
            A = A2;
            V = V2;
            uint32_t logS = 0;
            while ((1u << logS) < S) logS++;
            dbls += logS;
            items = groups;
        }
        if (V == nullptr) {  // c == 1: a single bucket per window, weight 1, no tree level ran
            xyzz_t* z = lvl + off;
            BPK_CUDA(cudaMemsetAsync(z, 0, (size_t)WB * sizeof(xyzz_t), ctx->stream));
            V = z;
        }
        t.end();
    }
    {
        StageTimer t(ctx, "msm.finalize");
        msm_finalize_kernel<<<1, 1, 0, ctx->stream>>>(A, V, WB, c, normalise ? 1 : 0, d_out);
        count_launch(ctx);
        BPK_CUDA(cudaGetLastError());
        t.end();
    }
    return BPK_OK;
}

int msm_run(bpk_ctx* ctx, const MsmPoints& pts, const fr_t* d_scalars, size_t n, unsigned rshift,
            bool normalise, uint64_t* d_out) {
    if (n >= ((size_t)1 << 31)) return BPK_ERR_TOO_LARGE;
    if (n == 0) return msm_empty_result(ctx, d_out);
    MsmPlan pl;
    BPK_TRY(msm_make_plan(ctx, pts, n, &pl));
    xyzz_t* buckets;
    BPK_TRY(ws_reserve(ctx, 4, (size_t)pl.nb_total * sizeof(xyzz_t), (void**)&buckets));
    BPK_TRY(msm_fill_buckets(ctx, pl, pts, d_scalars, n, rshift, buckets));
    return msm_reduce_buckets(ctx, pl, buckets, normalise, d_out);
}

// Scalars in HOST memory.  Large MSMs are cut into a small head slice and the rest: the head is uploaded and its
// bucket sums are accumulated while the copy engine brings the rest over PCIe on a second stream; the two bucket
// sets are added and reduced once.  With pinned host memory only the head's upload (1/8 of the bytes) stays
// exposed.  d_stage: device buffer for n scalars.
int msm_run_from_host(bpk_ctx* ctx, const MsmPoints& pts, const uint64_t* h_scalars, fr_t* d_stage, size_t n,
                      unsigned rshift, bool normalise, uint64_t* d_out) {
    if (n >= ((size_t)1 << 31)) return BPK_ERR_TOO_LARGE;
    const size_t head = n / 8;
    if (n < ((size_t)1 << 22) || ctx->opt_msm_host_slices == 0) {
        if (n) BPK_CUDA(cudaMemcpyAsync(d_stage, h_scalars, n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
        return msm_run(ctx, pts, d_stage, n, rshift, normalise, d_out);
    }
    MsmPlan pl;
    BPK_TRY(msm_make_plan(ctx, pts, n, &pl));
    xyzz_t *buckets, *buckets_b;
    BPK_TRY(ws_reserve(ctx, 4, (size_t)pl.nb_total * sizeof(xyzz_t), (void**)&buckets));
    BPK_TRY(ws_reserve(ctx, 15, (size_t)pl.nb_total * sizeof(xyzz_t), (void**)&buckets_b));
    BPK_TRY(msm_reserve(ctx, pl, n - head));
    if (!ctx->copy_stream) BPK_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->copy_done) BPK_CUDA(cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming));
    if (!ctx->lane_fork) BPK_CUDA(cudaEventCreateWithFlags(&ctx->lane_fork, cudaEventDisableTiming));
    // the staging buffer may still be read by earlier work on the main stream
    BPK_CUDA(cudaEventRecord(ctx->lane_fork, ctx->stream));
    BPK_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->lane_fork, 0));
    BPK_CUDA(cudaMemcpyAsync(d_stage, h_scalars, head * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    BPK_TRY(msm_fill_buckets(ctx, pl, pts, d_stage, head, rshift, buckets));
    BPK_CUDA(cudaMemcpyAsync(d_stage + head, h_scalars + 4 * head, (n - head) * sizeof(fr_t), cudaMemcpyHostToDevice,
                             ctx->copy_stream));
    BPK_CUDA(cudaEventRecord(ctx->copy_done, ctx->copy_stream));
    BPK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0));
    MsmPoints rest = pts;
    rest.base = pts.base + head;
    BPK_TRY(msm_fill_buckets(ctx, pl, rest, d_stage + head, n - head, rshift, buckets_b));
    {
        StageTimer t(ctx, "msm.join");
        msm_add_buckets_kernel<<<(unsigned)(((size_t)pl.nb_total + 127) / 128), 128, 0, ctx->stream>>>(buckets, buckets_b,
                                                                                                pl.nb_total);
        count_launch(ctx);
        BPK_CUDA(cudaGetLastError());
        t.end();
    }
    return msm_reduce_buckets(ctx, pl, buckets, normalise, d_out);
}

int g1_sum_run(bpk_ctx* ctx, const uint64_t* d_points_xyz, size_t n, uint64_t* d_out_xyz) {
    StageTimer t(ctx, "g1.sum");
    g1_sum_kernel<<<1, 1, 0, ctx->stream>>>(d_points_xyz, (uint32_t)n, d_out_xyz);
    count_launch(ctx);
    BPK_CUDA(cudaGetLastError());
    t.end();
    return BPK_OK;
}

}  // namespace bpk
