// BLS12-381 G1 multi-scalar multiplication (Pippenger) for sm_100a.
//
// Contract = BucketMSM::bucket_msm (src/msm.rs:76-118): sum_i (s_i >> rshift) * P_i, judged on the
// affine image of the result (G1Affine::from, g1.rs:49-63).  The reference walks 64 4-bit windows
// serially with complete projective additions; nothing of that structure is kept:
//
//   K1 count       scalar: Montgomery -> canonical (scalar.rs:292-304 semantics), signed c-bit digits; histogram
//                  of the bucket keys (warp-aggregated reductions in L2)
//   K2 layout      one scan over the buckets lays out EVERY level of the pairwise tree below: bucket b holds
//                  ceil(c_b / 2^l) entries at level l, padded to an even count
//   K3 scatter     the digits again, each (bucket, point | sign) pair placed at its bucket's cursor: a counting
//                  sort fused with the recoding (replaces the serial scatter, msm.rs:29-35)
//   K4 affine tree level l adds the entries 2j, 2j+1 of every bucket in AFFINE coordinates -- all additions of a
//                  level are independent, so each thread shares ONE field inversion over a batch of them
//                  (Montgomery's trick; per-thread division-step inversion, ff.cuh): 5 M + 1 S per addition instead
//                  of the 8 M + 2 S of XYZZ += affine.  A bucket that is down to one entry is written out.
//   K5 tail        whatever the affine levels leave (heavy buckets of skewed scalars, or everything for tiny inputs)
//                  goes through equal-size chunks of the sorted list in XYZZ coordinates, one thread per chunk;
//                  runs that cross chunk borders leave partial sums that three merge kernels fold -> load balance
//                  is independent of the scalar distribution
//   K6 reduce      sum_b (b+1) B_b by bit planes over one pairwise tree (msm.rs:42-46 is the serial form)
//   K7 finalize    Horner over planes / windows (msm.rs:107-115), affine normalisation, G1Projective limbs out
//
// Roofline: integer multiply pipe.  Algorithmic work per accumulated pair: one mixed addition
// = 8 M + 2 S in Fp, each 300 32x32->64 products; executed: 5 M + 1 S.  HBM: 8 B of sorted pair + 96 B point gather
// per pair and pass at level 0, 192 B in + 96 B out per addition above, 96 B of prefix-product scratch.

// One shared out-of-line body for the Fp product: the addition loops otherwise inline ~450-instruction
// multiplications many times over and stall on instruction fetch (14 % "no_instructions" samples in
// profiles/r1_final_msm_accumulate_ncu.md); measured 79.9 -> 77.3 ms at 2^24 in round 1.
#ifndef BPK_MSM_INLINE_MUL   // A/B: inline every product (build.py --variant inl BPK_MSM_INLINE_MUL)
#define BPK_FP_MUL_CALL 1
#endif
#include "internal.cuh"

namespace bpk {

static constexpr uint32_t INVALID_KEY = 0xffffffffu;
static constexpr int MSM_MAX_LEVELS = 29;  // affine levels the layout tables can describe (2^29 entries per bucket)

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
__device__ __forceinline__ void st_affine(affine_t* p, const affine_t& v) {
    st_fp(&p->x, v.x);
    st_fp(&p->y, v.y);
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
// K1 / K3: signed-digit recoding, shared by the counting and the scattering pass
// ------------------------------------------------------------------------------------------
// key_stride: buckets per window (2^(c-1)) for per-window bucket sets, 0 when every window shares one
// bucket set (precomputed SRS levels); val_stride: distance between precomputed levels in points, else 0.
struct RecodeArgs {
    const fr_t* scalars;   // Montgomery form (the caller's)
    uint32_t* digits;      // optional [W][n] cache of the recoded windows (bucket | sign << 31), written by the counting pass
    uint32_t n, c, W, rshift, key_stride, val_stride;
};

// canonical integer s -> (s >> rshift) as 10 limbs (two of padding for the window reads)
__device__ __forceinline__ void recode_shift(const RecodeArgs& a, const fr_t& s, uint32_t l[10]) {
#pragma unroll
    for (int k = 0; k < 8; k++) l[k] = s.l[k];
    l[8] = 0;
    l[9] = 0;
    if (a.rshift) {  // value >> rshift (the c-does-not-divide-256 quirk of msm.rs:119-139)
        uint32_t ws = a.rshift >> 5, bs = a.rshift & 31;
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
}

// scalar i: Montgomery -> canonical (scalar.rs:292-304 semantics) -> shifted limbs
__device__ __forceinline__ void recode_limbs(const RecodeArgs& a, uint32_t i, uint32_t l[10]) {
    const fr_t s = from_mont(ld_fr_g(a.scalars + i));  // canonical integer < q
    recode_shift(a, s, l);
}

// window w of the signed recoding: key = bucket (INVALID_KEY for a zero digit), val = point index | sign << 31
__device__ __forceinline__ void recode_window(const RecodeArgs& a, const uint32_t l[10], uint32_t i, uint32_t w,
                                              uint32_t& carry, uint32_t& key, uint32_t& val) {
    const uint32_t half = 1u << (a.c - 1), mask = (1u << a.c) - 1;
    const uint32_t bit = w * a.c, limb = bit >> 5, off = bit & 31;
    uint64_t two = 0;
    if (limb < 9) two = (uint64_t)l[limb] | ((uint64_t)l[limb + 1] << 32);
    const uint32_t raw = ((uint32_t)(two >> off) & mask) + carry;
    uint32_t d, neg;
    if (raw > half) {
        d = (1u << a.c) - raw;
        neg = 1;
        carry = 1;
    } else {
        d = raw;
        neg = 0;
        carry = 0;
    }
    key = d ? (w * a.key_stride + d - 1) : INVALID_KEY;
    val = (w * a.val_stride + i) | (neg << 31);
}

// Histogram of the bucket keys.  AGG: lanes of a warp that hit the same bucket are combined before the L2 reduction
// (match.any, ~300 cycles per warp instruction) -- worth it when there are few buckets; with many buckets and uniform
// digits it only costs, so the large path takes plain reductions plus the one vote that catches a warp whose lanes
// all hit the same bucket (equal scalars).
template <bool AGG>
__global__ void __launch_bounds__(256) msm_count_kernel(RecodeArgs a, uint32_t* __restrict__ cnt) {
    // a.digits != null: keep the recoded windows for the phases of the large scatter
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const bool active = i < a.n;
    uint32_t l[10];
    if (active) recode_limbs(a, i, l);
    uint32_t carry = 0;
    for (uint32_t w = 0; w < a.W; w++) {
        uint32_t key = INVALID_KEY, val = 0;
        if (active) recode_window(a, l, i, w, carry, key, val);
        if (a.digits && active) a.digits[(size_t)w * a.n + i] = (key & 0x7fffffffu) | (val & 0x80000000u);  // zero digit: 0x7fffffff
        if (AGG) {
            const uint32_t peers = __match_any_sync(0xffffffffu, key);
            if (key != INVALID_KEY && (peers & ((1u << lane) - 1)) == 0) atomicAdd(cnt + key, (uint32_t)__popc(peers));
        } else {
            const uint32_t k0 = __shfl_sync(0xffffffffu, key, 0);
            if (__all_sync(0xffffffffu, key == k0)) {
                if (lane == 0 && key != INVALID_KEY) atomicAdd(cnt + key, 32u);
            } else if (key != INVALID_KEY) {
                atomicAdd(cnt + key, 1u);
            }
        }
    }
}

// position of one entry: the bucket's (or partition's) cursor advances by one per entry.  `agg` (warp-uniform): combine
// the lanes of a warp that hit the same cursor first -- one atomic per distinct cursor instead of one per lane.
__device__ __forceinline__ uint32_t cursor_take(uint32_t* cursors, uint32_t idx, bool valid, bool agg, uint32_t lane) {
    if (agg) {
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? idx : INVALID_KEY);
        const uint32_t leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (valid && lane == leader) base = atomicAdd(cursors + idx, (uint32_t)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        return base + __popc(peers & ((1u << lane) - 1));
    }
    return valid ? atomicAdd(cursors + idx, 1u) : 0u;
}

// Counting-sort scatter, one pass: cursor[b] starts at the first level-0 slot of bucket b (K2).  The order of the entries
// inside a bucket depends on the order in which warps arrive; the bucket SUM (a group element, output in affine form) does
// not.  Used while the level-0 list is small enough to stay in L2; beyond that the random 8-byte stores become partial
// sector writes to HBM (measured 6.7 ms for 2^24 x 12 entries) and the two-pass form below takes over.
__global__ void __launch_bounds__(256) msm_scatter_kernel(RecodeArgs a, uint32_t* __restrict__ cursor,
                                                           const uint32_t* __restrict__ skew, uint32_t force_agg,
                                                           uint2* __restrict__ kv0) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const bool active = i < a.n;
    const bool agg = force_agg || *skew != 0;
    uint32_t l[10];
    if (active) recode_limbs(a, i, l);
    uint32_t carry = 0;
    for (uint32_t w = 0; w < a.W; w++) {
        uint32_t key = INVALID_KEY, val = 0;
        if (active) recode_window(a, l, i, w, carry, key, val);
        const uint32_t pos = cursor_take(cursor, key, key != INVALID_KEY, agg, lane);
        if (key != INVALID_KEY) kv0[pos] = make_uint2(key, val);
    }
}

// Scatter for large inputs, in PHASES over the bucket range.  A phase places only the entries whose bucket lies in
// [key_lo, key_hi): its random 8-byte stores are confined to that range's slice of the level-0 list, so that L2 merges
// more of them into whole sectors before they reach HBM (measured at 2^24 x 12 entries: 1 phase 6.9 ms, 5 phases 4.1 ms).
// The recoded windows come from the cache the counting pass wrote ([W][n] words, read with the streaming hint), so a phase
// costs one 4-byte load and a compare per entry.  `skew` (set by the layout scan when one bucket holds a large share
// of the entries) switches to warp-aggregated cursor updates.
#ifndef BPK_RANGE_ITEMS
#define BPK_RANGE_ITEMS 8
#endif
constexpr uint32_t RANGE_ITEMS = BPK_RANGE_ITEMS;  // cached digits per thread and phase (16-byte loads)
__global__ void __launch_bounds__(256) msm_scatter_range_kernel(RecodeArgs a, uint32_t* __restrict__ cursor,
                                                                 const uint32_t* __restrict__ skew, uint32_t key_lo,
                                                                 uint32_t key_hi, uint2* __restrict__ kv0) {
    const size_t M = (size_t)a.W * a.n;
    const size_t e0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * RANGE_ITEMS;
    const uint32_t lane = threadIdx.x & 31;
    const bool agg = *skew != 0;
    uint32_t d[RANGE_ITEMS];
    if (e0 + RANGE_ITEMS <= M) {
        const uint4* q = reinterpret_cast<const uint4*>(a.digits + e0);
#pragma unroll
        for (uint32_t k = 0; k < RANGE_ITEMS / 4; k++) {
            const uint4 u = __ldcs(q + k);
            d[4 * k] = u.x; d[4 * k + 1] = u.y; d[4 * k + 2] = u.z; d[4 * k + 3] = u.w;
        }
    } else {
#pragma unroll
        for (uint32_t k = 0; k < RANGE_ITEMS; k++) d[k] = e0 + k < M ? a.digits[e0 + k] : 0x7fffffffu;
    }
    // all cursor updates of the thread first, then its stores: the updates are what the phase waits for (ncu:
    // `long_scoreboard` 78 per issue), one in flight per thread leaves the L2 idle
    uint32_t pos[RANGE_ITEMS];
#pragma unroll
    for (uint32_t k = 0; k < RANGE_ITEMS; k++) {
        const uint32_t key = d[k] & 0x7fffffffu;
        const bool mine = key >= key_lo && key < key_hi;   // a zero digit (0x7fffffff) is above every range
        if (agg) pos[k] = cursor_take(cursor, key, mine, true, lane);
        else pos[k] = mine ? atomicAdd(cursor + key, 1u) : 0u;
    }
    uint32_t w = (uint32_t)(e0 / a.n), i = (uint32_t)(e0 - (size_t)w * a.n);
#pragma unroll
    for (uint32_t k = 0; k < RANGE_ITEMS; k++) {
        const uint32_t key = d[k] & 0x7fffffffu;
        const bool mine = key >= key_lo && key < key_hi;
        if (mine) kv0[pos[k]] = make_uint2(key, (w * a.val_stride + i) | (d[k] & 0x80000000u));
        if (++i == a.n) {
            i = 0;
            w++;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K2: layout of all tree levels from the histogram.  Bucket b with c0 entries holds c_l = ceil(c0 / 2^l) entries at
// level l; a bucket that is down to one entry (l >= 1) has left the tree (its sum sits in the bucket array).  Every
// level but the last pads a bucket to an even number of slots, so that the pair (2j, 2j + 1) never straddles two
// buckets; level L (the input of the XYZZ tail) is dense.  off[l][b] = first slot of bucket b at level l.
// ------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t level_entries(uint32_t c0, int l) { return (c0 + ((1u << l) - 1u)) >> l; }
__device__ __forceinline__ uint32_t level_slots(uint32_t c0, int l, int L) {
    if (c0 == 0) return 0;
    const uint32_t c = level_entries(c0, l);
    if (l > 0 && c < 2) return 0;
    return l == L ? c : (c + 1u) & ~1u;
}

// block-wide exclusive scan of one value per thread (SCAN_THREADS threads); returns the exclusive prefix, total to all
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t& total, uint32_t* smem /* 9 words */) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += t;
    }
    __syncthreads();  // smem may still be read from a previous call
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < SCAN_THREADS / 32 ? smem[lane] : 0, wi = w;
#pragma unroll
        for (int d = 1; d < SCAN_THREADS / 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= (uint32_t)d) wi += t;
        }
        if (lane < SCAN_THREADS / 32) smem[lane] = wi - w;  // exclusive warp bases
        if (lane == SCAN_THREADS / 32 - 1) smem[8] = wi;
    }
    __syncthreads();
    total = smem[8];
    return incl - v + smem[warp];
}

// stats: [0] entries at level 0, [1] entries at level L (finished buckets count 1), [2] non-empty buckets,
// [3] (low word) skew flag: some bucket holds more than 2^15 entries
__global__ void __launch_bounds__(SCAN_THREADS) msm_level_sums_kernel(const uint32_t* __restrict__ cnt, uint32_t nb, int L,
                                                                       uint32_t nblk, uint32_t* __restrict__ blocksums,
                                                                       unsigned long long* __restrict__ stats) {
    __shared__ uint32_t smem[9];
    uint32_t c0[SCAN_ITEMS];
    const uint32_t first = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) c0[k] = first + k < nb ? cnt[first + k] : 0;
    for (int l = 0; l <= L; l++) {
        uint32_t s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) s += level_slots(c0[k], l, L);
        uint32_t total;
        block_exclusive_scan(s, total, smem);
        if (threadIdx.x == 0) blocksums[(size_t)l * nblk + blockIdx.x] = total;
    }
    uint32_t e0 = 0, eL = 0, ne = 0, heavy = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        e0 += c0[k];
        eL += c0[k] ? level_entries(c0[k], L) : 0;
        ne += c0[k] != 0;
        heavy |= c0[k] > (1u << 15);
    }
    if (heavy) *reinterpret_cast<volatile uint32_t*>(stats + 3) = 1u;
    uint32_t t0, tL, tn;
    block_exclusive_scan(e0, t0, smem);
    block_exclusive_scan(eL, tL, smem);
    block_exclusive_scan(ne, tn, smem);
    if (threadIdx.x == 0) {
        atomicAdd(stats + 0, (unsigned long long)t0);
        atomicAdd(stats + 1, (unsigned long long)tL);
        atomicAdd(stats + 2, (unsigned long long)tn);
    }
}

// one block per level: exclusive scan of the per-block sums in place, level total out
__global__ void __launch_bounds__(1024) msm_level_scan_kernel(uint32_t* __restrict__ blocksums, uint32_t nblk,
                                                               uint32_t* __restrict__ totals) {
    __shared__ uint32_t part[1024];
    uint32_t* row = blocksums + (size_t)blockIdx.x * nblk;
    const uint32_t per = (nblk + 1023) / 1024;
    const uint32_t lo = threadIdx.x * per, hi = lo + per < nblk ? lo + per : nblk;
    uint32_t s = 0;
    for (uint32_t i = lo; i < hi; i++) s += row[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {  // Hillis-Steele on 1024 partial sums
        uint32_t t = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - s;
    for (uint32_t i = lo; i < hi; i++) {
        uint32_t v = row[i];
        row[i] = run;
        run += v;
    }
    if (threadIdx.x == 1023) totals[blockIdx.x] = part[1023];
}

__global__ void __launch_bounds__(SCAN_THREADS) msm_level_offsets_kernel(const uint32_t* __restrict__ cnt, uint32_t nb, int L,
                                                                          uint32_t nblk, const uint32_t* __restrict__ blockbase,
                                                                          uint32_t* __restrict__ off /* [(L+1)][nb] */,
                                                                          uint32_t* __restrict__ cursor,
                                                                          uint2* __restrict__ kv0) {
    __shared__ uint32_t smem[9];
    uint32_t c0[SCAN_ITEMS];
    const uint32_t first = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) c0[k] = first + k < nb ? cnt[first + k] : 0;
    for (int l = 0; l <= L; l++) {
        uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            v[k] = level_slots(c0[k], l, L);
            s += v[k];
        }
        uint32_t total;
        uint32_t run = block_exclusive_scan(s, total, smem) + blockbase[(size_t)l * nblk + blockIdx.x];
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (first + k < nb) {
                off[(size_t)l * nb + first + k] = run;
                if (l == 0) {
                    cursor[first + k] = run;
                    if (L > 0 && (c0[k] & 1u)) kv0[run + c0[k]] = make_uint2(INVALID_KEY, 0);  // the pad slot
                }
            }
            run += v[k];
        }
    }
}

// ------------------------------------------------------------------------------------------
// K4: one level of the pairwise tree in affine coordinates
// ------------------------------------------------------------------------------------------
struct AffLevelArgs {
    const uint2* kv0;          // level 0: (bucket, point | sign) per slot
    const uint32_t* keys_in;   // level >= 1: bucket per slot
    const affine_t* pts_in;    // level 0: the SRS (all precomputed levels); level >= 1: this level's points
    uint32_t* keys_out;
    affine_t* pts_out;         // level + 1
    const uint32_t* cnt0;
    const uint32_t* off_in;    // off[level]
    const uint32_t* off_out;   // off[level + 1]
    const uint32_t* totals;    // slots per level
    xyzz_t* buckets;           // finished buckets, as XYZZ (x, y, 1, 1)
    uint4* scratch;            // prefix products: [bmax][3][threads of the launch]
    uint32_t* rare_bits;       // one flag per pair: degenerate, left to msm_affine_rare_kernel
    uint32_t* claim;           // pairs of this level handed out so far
    uint32_t level, bmax, out_is_tail;
};

// 16 warps per SM either way (128 registers per thread): four CTAs of four warps at the large levels, ONE CTA of sixteen
// at the small ones (where a warp has a single batch): the same code, measured 4 % faster there (2^21 pairs: accumulate
// 8.00 -> 7.69 ms) and 2.7 % slower at the levels with guided batches (profiles/r2_affine_v2.md).
constexpr int AFF_THREADS = 128, AFF_THREADS_SMALL = 512;
constexpr int AFF_WARPS_PER_SM = 16;
#ifndef BPK_AFF_MIN_BATCH
#define BPK_AFF_MIN_BATCH 32
#endif
#ifndef BPK_AFF_GUIDED_ROUNDS
#define BPK_AFF_GUIDED_ROUNDS 2u   // guided sizes from this many full batches per warp
#endif
#ifndef BPK_AFF_GUIDED_FACTOR
#define BPK_AFF_GUIDED_FACTOR 2u   // a claim = this many even shares of what is left
#endif
constexpr uint32_t AFF_MIN_BATCH = BPK_AFF_MIN_BATCH;   // the shortest batch a warp claims (steps of 32 pairs)
// How the level kernel touches memory (profiles/r2_affine_ab.md, profiles/r2_affine_v2.md):
//  * operands: what bounds a kernel in which every lane loads its own 96-byte points is the SM's single L1TEX queue (a
//    warp-wide 16-byte load whose lanes touch 32 different lines costs 32 wavefronts), so the warp gathers
//    COOPERATIVELY: 3 neighbouring lanes copy the 16-byte pieces of one coordinate with cp.async, eight points per
//    instruction, into the warp's stage in shared memory;
//  * nothing is loaded from global or local memory into registers inside the loops: a register load must have landed
//    before the next out-of-line product is called (the callee may use the register), which exposes its whole latency.
//    The pair lists, the bucket layout words and the prefix products all arrive through cp.async one step ahead.  The
//    degenerate pairs (identity operands, doubling, opposite points) are only flagged here and added by
//    msm_affine_rare_kernel after the level: a call to a function that handles them -- even on a branch that is never
//    taken -- puts every value that lives across it into local memory;
//  * the operands stay in the stage and are read where they are used (x1, x2 twice) instead of living in registers
//    across the six products of an addition, and the values that would otherwise be carried across most of them have a
//    place in shared memory -- the running inverse and the slope a home of their own, y1 stays where it was copied and
//    x1 - x3 is parked where Q.y was: an out-of-line product leaves the caller ~55 registers, and a value spilled to
//    local memory comes back with L2 latency (the hot spill slots of 16 warps do not fit L1; measured: 25 % of all
//    stall samples).  The x copies for the next pair are issued after the last read of x1, x2 (before the last
//    product), the y copies at the end of the step.
// Per warp: 32 pairs x 208 B (P.x P.y Q.x Q.y + 16 B so that the lanes' 16-byte reads fall into different banks), 32
// prefix products, 32 pair-list words x 16 B, 32 x (count, offsets) x 16 B, 32 running inverses, 32 slopes (48 B each,
// as [3][32] 16-byte pieces).
#ifndef BPK_AFF_COPY_UNROLL
#define BPK_AFF_COPY_UNROLL 8   // the 8 copy instructions per coordinate, fully unrolled (accumulate at 2^24: unroll 2 53.4, 4 52.6, 8 51.6 ms)
#endif
constexpr int AFF_COPY_UNROLL = BPK_AFF_COPY_UNROLL;
#ifndef BPK_AFF_LAZY
#define BPK_AFF_LAZY 1    // products of the tree stay in [0, 2p): six conditional subtractions fewer per addition
#endif
constexpr uint32_t AFF_PAIR_STRIDE = 208;
constexpr uint32_t AFF_PRE_OFF = 32 * AFF_PAIR_STRIDE, AFF_META_OFF = AFF_PRE_OFF + 32 * 48, AFF_DEST_OFF = AFF_META_OFF + 32 * 16,
                   AFF_ACC_OFF = AFF_DEST_OFF + 32 * 16, AFF_LAM_OFF = AFF_ACC_OFF + 32 * 48;
constexpr int AFF_WARP_SMEM = AFF_LAM_OFF + 32 * 48;                           // 12 KB per warp
constexpr size_t AFF_SMEM_BYTES = (size_t)(AFF_THREADS / 32) * AFF_WARP_SMEM;              // 48 KB per CTA
constexpr size_t AFF_SMEM_BYTES_SMALL = (size_t)(AFF_THREADS_SMALL / 32) * AFF_WARP_SMEM;  // 192 KB per CTA

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all_but_last() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
// shared memory by 32-bit address: one register per pointer, and statements that stay where they are written
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <uint32_t STRIDE = 16>
__device__ __forceinline__ fp_t lds_fp(uint32_t addr) {  // three 16-byte pieces, STRIDE bytes apart
    const uint4 a = lds128(addr), b = lds128(addr + STRIDE), c = lds128(addr + 2 * STRIDE);
    fp_t r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    r.l[8] = c.x; r.l[9] = c.y; r.l[10] = c.z; r.l[11] = c.w;
    return r;
}
template <uint32_t STRIDE = 16>
__device__ __forceinline__ void sts_fp(uint32_t addr, const fp_t& v) {
    sts128(addr, make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]));
    sts128(addr + STRIDE, make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]));
    sts128(addr + 2 * STRIDE, make_uint4(v.l[8], v.l[9], v.l[10], v.l[11]));
}
__device__ __forceinline__ fp_t aff_mul(const fp_t& x, const fp_t& y) { return BPK_AFF_LAZY ? mul_lazy(x, y) : mul(x, y); }
__device__ __forceinline__ fp_t aff_sqr(const fp_t& x) { return BPK_AFF_LAZY ? sqr_lazy(x) : sqr(x); }
__device__ __forceinline__ fp_t aff_sub(const fp_t& x, const fp_t& y) { return BPK_AFF_LAZY ? sub_lazy(x, y) : sub(x, y); }
__device__ __forceinline__ fp_t aff_canon(const fp_t& x) { return BPK_AFF_LAZY ? reduce_once(x) : x; }

constexpr uint32_t AFF_F_PAD = 1, AFF_F_NEG_P = 2, AFF_F_NEG_Q = 4, AFF_F_FINISHED = 8, AFF_F_MARK_PAD = 16, AFF_F_X_ZERO = 32;

template <bool LEVEL0, int THREADS>
__global__ void __launch_bounds__(THREADS, AFF_WARPS_PER_SM * 32 / THREADS) msm_affine_level_kernel(const __grid_constant__ AffLevelArgs a) {
    extern __shared__ uint4 aff_smem[];
    const uint32_t S = a.totals[a.level] >> 1;  // pairs of this level
    if (S == 0) return;
    const uint32_t T = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    // A warp works on 32 neighbouring pairs at every step and shares one inversion per lane over a batch of B steps.  The
    // batches are CLAIMED from a counter (at the large levels with guided sizes between AFF_MIN_BATCH and bmax steps):
    // with a fixed share per warp the schedulers' preference for their oldest warp let some warps finish a fifth of the
    // kernel before others (ncu: 3.17 of 4 warps resident per scheduler on average), and the pipe idles with them.
    // Pair numbers fit 32 bits: S <= 2^31, and the counter overshoots S by less than 32 bmax per warp.
    const uint32_t nwarps = T >> 5;
    constexpr uint32_t PS = 32;                           // pairs per step of a warp
    // Guided sizes only where a warp's even share is several full batches (the two lowest levels of a large MSM): every
    // batch costs an inversion (~7 additions), so short batches at the end of a small level cost more than the idle
    // warps did (measured at 2^24: level 0 26.3 -> 24.1 ms, level 1 12.5 -> 11.8, level 2 with guided sizes 6.44 -> 6.54).
    // Elsewhere: equal batches, as few per warp as bmax allows.
    const uint32_t even = (S - 1) / (PS * nwarps) + 1;    // steps per warp of an even split
    const bool guided = even >= BPK_AFF_GUIDED_ROUNDS * a.bmax;
    const uint32_t rounds = (even - 1) / a.bmax + 1;
    const uint32_t Beq = (even - 1) / rounds + 1;
    uint4* const sc = a.scratch + tid;
    // this warp's stage (layout above): the pair of lane l at 208 l, this lane's 16-byte slots of the other arrays at
    // my16 + the array's offset
    const uint32_t wstage = (uint32_t)__cvta_generic_to_shared(aff_smem) + (threadIdx.x >> 5) * AFF_WARP_SMEM;
    const uint32_t my_pair = wstage + lane * AFF_PAIR_STRIDE;
    const uint32_t my16 = wstage + lane * 16;
    const uint32_t my_pre = my16 + AFF_PRE_OFF, my_meta = my16 + AFF_META_OFF, my_dest = my16 + AFF_DEST_OFF;
    const uint32_t my_acc = my16 + AFF_ACC_OFF, my_lam = my16 + AFF_LAM_OFF;

    // pair-list word of pair p: level 0 (key, val) of both slots; above: the two keys.  A lane without a pair at this step
    // gets a harmless one (point 0, padded) so that it can take part in the warp's copies.
    auto load_meta = [&](uint32_t p) -> uint4 {
        if (p >= S) return make_uint4(0, 0, INVALID_KEY, 0);
        if (LEVEL0) return reinterpret_cast<const uint4*>(a.kv0)[p];
        const uint2 kk = reinterpret_cast<const uint2*>(a.keys_in)[p];
        return make_uint4(kk.x, 0, kk.y, 0);
    };
    auto stage_meta = [&](uint32_t p, uint32_t slot_addr) {   // the same word, into a slot of this lane in the stage
        if (p >= S) sts128(slot_addr, LEVEL0 ? make_uint4(0, 0, INVALID_KEY, 0) : make_uint4(0, INVALID_KEY, 0, 0));
        else if (LEVEL0) cp_async16(slot_addr, reinterpret_cast<const uint4*>(a.kv0) + p);
        else cp_async8(slot_addr, reinterpret_cast<const uint2*>(a.keys_in) + p);
    };
    auto staged_meta = [&](uint32_t slot_addr) -> uint4 {
        const uint4 s = lds128(slot_addr);
        return LEVEL0 ? s : make_uint4(s.x, 0, s.y, 0);
    };
    auto stage_dest = [&](uint32_t key) {                 // where the sum of a pair of bucket `key` goes
        cp_async4(my_dest, a.cnt0 + key);
        cp_async4(my_dest + 4, a.off_in + key);
        cp_async4(my_dest + 8, a.off_out + key);
    };
    // Cooperative copy of one coordinate (SRC 0: x, 1: y) of the operands of the warp's 32 pairs (first pair p0) into the
    // stage: 3 neighbouring lanes copy the 16-byte pieces of one coordinate, 8 points per instruction (even, so that the
    // destination is linear in i).  Point q of the warp (q = 2 lane' + which) lives at byte 208 lane' + 96 which, the
    // coordinate in its first (HALF 0) or second 48 bytes.  vP, vQ: this lane's level-0 point words (a padded pair
    // copies P twice).
    auto stage_points = [&](uint32_t p0, uint32_t vP, uint32_t vQ, const int SRC, const int HALF) {
        const uint32_t sub = lane / 3, piece = lane - sub * 3;
        const bool on = sub < 8;                           // 24 lanes copy
        const uint32_t dst0 = wstage + (sub >> 1) * AFF_PAIR_STRIDE + (sub & 1) * 96 + HALF * 48 + piece * 16;
        const uint32_t iP = vP & 0x7fffffffu, iQ = vQ & 0x7fffffffu;
#pragma unroll AFF_COPY_UNROLL
        for (int i = 0; i < 8; i++) {
            const affine_t* src;
            if (LEVEL0) {
                const uint32_t from = (i * 4 + (sub >> 1)) & 31u;
                const uint32_t wP = __shfl_sync(0xffffffffu, iP, from), wQ = __shfl_sync(0xffffffffu, iQ, from);
                src = a.pts_in + ((sub & 1) ? wQ : wP);
            } else {
                src = a.pts_in + 2 * (size_t)p0 + i * 8 + sub;
            }
            if (on) cp_async16(dst0 + i * 4 * AFF_PAIR_STRIDE, reinterpret_cast<const uint4*>(src) + SRC * 3 + piece);
        }
    };
    // the running product after pair jj of this thread's batch: scratch (written by the forward pass) -> stage
    auto stage_prefix = [&](uint32_t jj) {
        const uint4* src = sc + (size_t)jj * 3 * T;
        cp_async16(my_pre, src);
        cp_async16(my_pre + 32 * 16, src + T);
        cp_async16(my_pre + 64 * 16, src + 2 * (size_t)T);
    };
    // level 0: the sign of an entry is bit 31 of its point word
    auto signed_y = [&](uint32_t addr, uint32_t minus) -> fp_t {
        const fp_t y = lds_fp(addr);
        return LEVEL0 && minus ? neg(y) : y;
    };
    auto flags_of = [&](const uint4& m) -> uint32_t {
        const bool pad = m.z == INVALID_KEY;
        return (pad ? AFF_F_PAD : 0u) | (LEVEL0 && (m.y >> 31) ? AFF_F_NEG_P : 0u) |
               (LEVEL0 && !pad && (m.w >> 31) ? AFF_F_NEG_Q : 0u);
    };

    // (Starting the CTAs of an SM with first batches of different length, so that their forward / inversion / backward
    // phases interleave, was measured: 59.5 against 58.6 ms -- the warps are not phase-locked.)
    for (;;) {
        uint32_t base0 = 0, B = 0;                        // the warp's first pair of this batch, its steps
        if (lane == 0) {
            B = Beq;
            if (guided) {                                 // twice an even share of what is left
                const uint32_t seen = *reinterpret_cast<volatile uint32_t*>(a.claim);
                const uint32_t left = seen < S ? S - seen : 0u;
                B = BPK_AFF_GUIDED_FACTOR * (left / (PS * nwarps));
                if (B < AFF_MIN_BATCH) B = AFF_MIN_BATCH;
            }
            if (B > a.bmax) B = a.bmax;                   // (the prefix scratch holds bmax steps per thread)
            base0 = atomicAdd(a.claim, PS * B);
        }
        base0 = __shfl_sync(0xffffffffu, base0, 0);
        B = __shfl_sync(0xffffffffu, B, 0);
        if (base0 >= S) break;                            // warp-uniform
        const uint32_t base = base0 + lane;
        uint32_t nj = 0;
        if (base < S) {
            const uint32_t left = (S - base + PS - 1) / PS;
            nj = left < B ? left : B;
        }
        const uint32_t njw = __shfl_sync(0xffffffffu, nj, 0);   // lane 0 has the most

        // ---- forward: denominators and their running product, two steps per turn: only the x coordinates are needed,
        // those of the even step sit in the x half of the stage and those of the odd step in the y half, so that the
        // copies for the next two steps have two products to land (with one step per turn the wait at the top of the
        // loop was 7 % of all stall samples).  A degenerate pair (identity operand, equal x) is flagged and left out.
        fp_t prod = fp_t::one();
        {
            const uint32_t my_meta2 = my_dest;             // the layout-word slot holds the odd step's list word here
            uint32_t pad0, pad1;
            {
                const uint4 m0 = load_meta(base), m1 = load_meta(base + PS);
                pad0 = m0.z == INVALID_KEY;
                pad1 = m1.z == INVALID_KEY;
                stage_meta(base + 2 * PS, my_meta);
                stage_meta(base + 3 * PS, my_meta2);
                stage_points(base0, m0.y, pad0 ? m0.y : m0.w, 0, 0);
                if (njw > 1) stage_points(base0 + PS, m1.y, pad1 ? m1.y : m1.w, 0, 1);   // (never past the level's lists)
                cp_async_commit();
            }
            uint32_t p = base;                             // pair of step j
            uint4* s = sc;                                 // its slot of the prefix scratch
            auto step = [&](uint32_t j, uint32_t pj, uint4* sj, bool pad, const fp_t& x1, const fp_t& x2) {
                fp_t den = sub(x2, x1);
                if (pad) {
                    den = fp_t::one();
                } else if (den.is_zero() || x1.is_zero() || x2.is_zero()) {   // degenerate: not in this product
                    den = fp_t::one();
                    atomicOr(a.rare_bits + (pj >> 5), 1u << (pj & 31));
                }
                prod = j == 0 ? den : aff_mul(prod, den);
                sj[0] = make_uint4(prod.l[0], prod.l[1], prod.l[2], prod.l[3]);
                sj[T] = make_uint4(prod.l[4], prod.l[5], prod.l[6], prod.l[7]);
                sj[2 * (size_t)T] = make_uint4(prod.l[8], prod.l[9], prod.l[10], prod.l[11]);
            };
            for (uint32_t j = 0; j < njw; j += 2, p += 2 * PS, s += 6 * (size_t)T) {
                cp_async_wait_all();
                __syncwarp();
                const fp_t x1a = lds_fp(my_pair), x2a = pad0 ? x1a : lds_fp(my_pair + 96);
                const fp_t x1b = lds_fp(my_pair + 48), x2b = pad1 ? x1b : lds_fp(my_pair + 144);
                const uint4 mA = staged_meta(my_meta), mB = staged_meta(my_meta2);   // steps j + 2, j + 3
                __syncwarp();
                const bool padA = mA.z == INVALID_KEY, padB = mB.z == INVALID_KEY;
                if (j + 2 < njw) {
                    stage_meta(p + 4 * PS, my_meta);
                    stage_meta(p + 5 * PS, my_meta2);
                    stage_points(p - lane + 2 * PS, mA.y, padA ? mA.y : mA.w, 0, 0);
                    if (j + 3 < njw) stage_points(p - lane + 3 * PS, mB.y, padB ? mB.y : mB.w, 0, 1);
                    cp_async_commit();
                }
                if (j < nj) step(j, p, s, pad0, x1a, x2a);
                if (j + 1 < nj) step(j + 1, p + PS, s + 3 * (size_t)T, pad1, x1b, x2b);
                pad0 = padA;
                pad1 = padB;
            }
        }

        sts_fp<512>(my_acc, inv(prod));  // 1 / (den_0 ... den_{nj-1}): the running inverse, at home in the stage

        // ---- backward: peel the inverses off, finish the additions, place the results.  Two copy groups per step: the x
        // coordinates, list word, layout words and prefix of the next pair go out once x1, x2 have been read for the last
        // time (before the last product); its y coordinates after the step's last read of y1 and t (which is parked where
        // Q.y was), and are waited for only after the next step's first two products.
        {
            uint32_t keyN, flagsN, vPN, vQN;
            uint32_t p = base + (njw - 1) * PS;            // pair of step j
            {
                const uint4 m = load_meta(p);
                keyN = m.x;
                flagsN = flags_of(m);
                vPN = m.y;
                vQN = (flagsN & AFF_F_PAD) ? m.y : m.w;
                if (njw > 1) stage_meta(p - PS, my_meta);
                stage_dest(keyN);
                stage_points(p - lane, vPN, vQN, 0, 0);
                if (njw > 1 && nj == njw) stage_prefix(njw - 2);   // needed at the first step if this lane takes part in it
                cp_async_commit();
                stage_points(p - lane, vPN, vQN, 1, 1);
                cp_async_commit();
            }
            for (uint32_t j = njw; j-- > 0; p -= PS) {
                const uint32_t key = keyN;
                uint32_t flags = flagsN;
                const bool active = j < nj;
                cp_async_wait_all_but_last();
                __syncwarp();
                // where the result goes (the layout words of `key` were staged one step ahead)
                const uint4 d = lds128(my_dest);
                const uint32_t cn = (d.x + ((2u << a.level) - 1u)) >> (a.level + 1);  // entries of the bucket at level + 1
                const uint32_t rel = p - (d.y >> 1);
                const uint32_t slot = d.z + rel;
                if (cn <= 1) flags |= AFF_F_FINISHED;
                if (!a.out_is_tail && (cn & 1u) && rel == cn - 1) flags |= AFF_F_MARK_PAD;
                const bool pad = flags & AFF_F_PAD;
                // the sum goes to the next level's list, or, when it is the bucket's last, to the bucket (x, y first in both)
                auto dst_fp = [&](int which) -> fp_t* {
                    return (flags & AFF_F_FINISHED) ? &a.buckets[key].X + which : &a.pts_out[slot].x + which;
                };

                fp_t dinv;
                bool plain = false;
                if (active) {
                    const fp_t x1 = lds_fp(my_pair);
                    const fp_t x2 = pad ? x1 : lds_fp(my_pair + 96);
                    fp_t den = sub(x2, x1);
                    if (pad || !(den.is_zero() || x1.is_zero() || x2.is_zero())) {   // else: flagged by the forward pass
                        plain = true;
                        if (pad) den = fp_t::one();
                        if (j > 0) {
                            dinv = aff_mul(lds_fp<512>(my_acc), lds_fp<512>(my_pre));
                            sts_fp<512>(my_acc, aff_mul(lds_fp<512>(my_acc), den));
                        } else {
                            dinv = lds_fp<512>(my_acc);
                        }
                    }
                }
                cp_async_wait_all();                       // the y coordinates
                __syncwarp();
                if (plain) {
                    {
                        const fp_t y1 = signed_y(my_pair + 48, flags & AFF_F_NEG_P);
                        const fp_t y2 = pad ? y1 : signed_y(my_pair + 144, flags & AFF_F_NEG_Q);
                        if (LEVEL0 && (flags & AFF_F_NEG_P)) sts_fp(my_pair + 48, y1);   // read back at the end of the step
                        dinv = aff_mul(sub(y2, y1), dinv);                                // the slope
                    }
                    sts_fp<512>(my_lam, dinv);
                    const fp_t l2 = aff_sqr(dinv);
                    const fp_t x1 = lds_fp(my_pair);
                    const fp_t x2 = pad ? x1 : lds_fp(my_pair + 96);
                    fp_t x3 = pad ? x1 : aff_sub(aff_sub(l2, x1), x2);   // a padded pair: the sum is P
                    sts_fp(my_pair + 144, aff_sub(x1, x3));              // x1 - x3, parked where Q.y was
                    x3 = aff_canon(x3);
                    if (x3.is_zero()) flags |= AFF_F_X_ZERO;
                    st_fp(dst_fp(0), x3);
                }
                // every lane has read x1, x2 for the last time
                __syncwarp();
                if (j > 0) {
                    const uint4 mN = staged_meta(my_meta);
                    keyN = mN.x;
                    flagsN = flags_of(mN);
                    vPN = mN.y;
                    vQN = (flagsN & AFF_F_PAD) ? mN.y : mN.w;
                    if (j > 1) stage_meta(p - 2 * PS, my_meta);
                    stage_dest(keyN);
                    stage_points(p - lane - PS, vPN, vQN, 0, 0);
                    // the prefix needed at step j - 1 is the product after pair j - 2 (own slots: no other lane reads them)
                    if (j > 1 && j - 1 < nj) stage_prefix(j - 2);
                }
                cp_async_commit();
                if (plain) {
                    const fp_t lt = aff_mul(lds_fp<512>(my_lam), lds_fp(my_pair + 144));
                    const fp_t y1 = lds_fp(my_pair + 48);
                    fp_t y3 = aff_canon(aff_sub(lt, y1));
                    if (pad) y3 = y1;
                    st_fp(dst_fp(1), y3);
                    if (flags & AFF_F_FINISHED) {          // as xyzz_t::from_affine: (x, y, 1, 1), the identity has ZZ = 0
                        const fp_t zz = (flags & AFF_F_X_ZERO) && y3.is_zero() ? fp_t::zero() : fp_t::one();
                        st_fp(dst_fp(2), zz);
                        st_fp(dst_fp(3), zz);
                    } else {
                        a.keys_out[slot] = key;
                        if (flags & AFF_F_MARK_PAD) a.keys_out[slot + 1] = INVALID_KEY;
                    }
                }
                // .. and y1, x1 - x3
                __syncwarp();
                if (j > 0) stage_points(p - lane - PS, vPN, vQN, 1, 1);
                cp_async_commit();
            }
        }
    }
}

// The degenerate pairs of a level, flagged by the level kernel's forward pass (an identity operand, or equal x: a doubling
// or opposite points): each is added on its own, with its own inversion, and placed like the others.  Clears the flags.
// Random inputs have none; an SRS whose points coincide (the reference's own tests) goes through here entirely.
template <bool LEVEL0>
__global__ void __launch_bounds__(128) msm_affine_rare_kernel(const __grid_constant__ AffLevelArgs a) {
    const uint32_t S = a.totals[a.level] >> 1;
    const uint32_t words = (S + 31) >> 5;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += gridDim.x * blockDim.x) {
        uint32_t bits = a.rare_bits[w];
        if (bits == 0) continue;
        a.rare_bits[w] = 0;
        while (bits) {
            const uint32_t p = (w << 5) + (uint32_t)__ffs((int)bits) - 1;
            bits &= bits - 1;
            uint32_t key;
            affine_t P, Q;
            if (LEVEL0) {
                const uint4 m = reinterpret_cast<const uint4*>(a.kv0)[p];
                key = m.x;
                P = ld_affine(a.pts_in + (m.y & 0x7fffffffu));
                Q = ld_affine(a.pts_in + (m.w & 0x7fffffffu));
                if (m.y >> 31) P.y = neg(P.y);
                if (m.w >> 31) Q.y = neg(Q.y);
            } else {
                key = a.keys_in[2 * (size_t)p];
                P = ld_affine(a.pts_in + 2 * (size_t)p);
                Q = ld_affine(a.pts_in + 2 * (size_t)p + 1);
            }
            fp_t den;
            const int kind = affine_add_prepare(P, Q, den);
            // (a point with x = 0 that is not the identity -- (0, +-2) lies on the curve -- arrives here as a plain addition)
            const affine_t R = affine_add_finish(kind, P, Q, (kind == AFF_DBL || kind == AFF_ADD) ? inv(den) : den);
            const uint32_t cn = (a.cnt0[key] + ((2u << a.level) - 1u)) >> (a.level + 1);
            if (cn <= 1) {
                st_xyzz(a.buckets + key, xyzz_t::from_affine(R));
            } else {
                const uint32_t rel = p - (a.off_in[key] >> 1);
                const uint32_t slot = a.off_out[key] + rel;
                st_affine(a.pts_out + slot, R);
                a.keys_out[slot] = key;
                if (!a.out_is_tail && (cn & 1u) && rel == cn - 1) a.keys_out[slot + 1] = INVALID_KEY;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// K5: chunked bucket accumulation in XYZZ coordinates (the tail of the affine tree, or the whole accumulation when no
// affine level runs).  The list is sorted by bucket and dense; its length comes from the layout scan.
// ------------------------------------------------------------------------------------------
struct TailSrc {
    const uint2* kv0;         // level 0 list (points gathered through val), or null
    const uint32_t* keys;     // level >= 1 list (point e of pts)
    const affine_t* pts;
    const uint32_t* total;    // number of entries (device)
};
template <bool LEVEL0>
__device__ __forceinline__ uint32_t tail_key(const TailSrc& s, size_t e) {
    return LEVEL0 ? s.kv0[e].x : s.keys[e];
}
template <bool LEVEL0>
__device__ __forceinline__ affine_t tail_point(const TailSrc& s, size_t e) {
    if (LEVEL0) {
        const uint32_t v = s.kv0[e].y;
        affine_t p = ld_affine(s.pts + (v & 0x7fffffffu));
        if (v >> 31) p.y = neg(p.y);
        return p;
    }
    return ld_affine(s.pts + e);
}

#ifndef BPK_ACC_MINBLOCKS
#define BPK_ACC_MINBLOCKS 4  // 4 x 128 threads / SM = 128 registers per thread: ~0.5 KB of spills, but 16 warps hide the
                             // dependent-issue waits better than 12 (2^24: 74.6 -> 73.2 ms; 2 CTAs: 77.0, 5 CTAs: 77.4)
#endif
template <bool LEVEL0>
__global__ void __launch_bounds__(128, BPK_ACC_MINBLOCKS) msm_accumulate_kernel(TailSrc src, uint32_t chunk, size_t num_chunks,
                                                              xyzz_t* __restrict__ buckets,
                                                              uint32_t* __restrict__ pkeys,
                                                              xyzz_t* __restrict__ pvals) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_chunks) return;
    const size_t M = *src.total;
    const size_t a = t * chunk;
    uint32_t pk0 = INVALID_KEY, pk1 = INVALID_KEY;
    if (a < M) {
        const size_t b = (a + chunk < M) ? a + chunk : M;
        uint32_t cur = INVALID_KEY;
        bool left_open = false;
        xyzz_t acc = xyzz_t::inf();
        // software pipeline: the point of entry e + 1 is fetched while entry e is added
        uint32_t k_next = tail_key<LEVEL0>(src, a);
        affine_t p_next = tail_point<LEVEL0>(src, a);
        for (size_t e = a; e < b; e++) {
            const uint32_t k = k_next;
            const affine_t p = p_next;
            if (e + 1 < b) {
                k_next = tail_key<LEVEL0>(src, e + 1);
                p_next = tail_point<LEVEL0>(src, e + 1);
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
                left_open = (e == a) && (a > 0) && (tail_key<LEVEL0>(src, a - 1) == k);
                acc = xyzz_t::from_affine(p);
            } else {
                xyzz_madd(acc, p);
            }
        }
        const bool right_open = (b < M) && (tail_key<LEVEL0>(src, b) == cur);
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

// K5b: a run that crosses chunk borders starts as slot 1 of some chunk t0 (the head) and continues as slot 0
// of the chunks t0+1 .. t1 whose first entry has the same key.  The thread that owns the head finds t1 by a
// binary search over the first keys of the chunks (sorted), adds short runs itself and queues long runs
// (heavy buckets: skewed scalars, narrow top windows) for a warp- or block-wide tree reduction, so that no scalar
// distribution can serialise the merge.
constexpr uint32_t MERGE_SERIAL_MAX = 8;    // runs up to this many partials: added by the head's thread
constexpr uint32_t MERGE_WARP_MAX = 256;    // up to this many: one warp per run; longer: one block per run
struct LongRun {
    uint32_t key, t0, t1, pad;
};

template <bool LEVEL0>
__global__ void __launch_bounds__(128, 3) msm_merge_partials_kernel(const uint32_t* __restrict__ pkeys,
                                                                  const xyzz_t* __restrict__ pvals,
                                                                  size_t num_chunks, xyzz_t* __restrict__ buckets,
                                                                  TailSrc src, uint32_t chunk,
                                                                  LongRun* __restrict__ long_runs,
                                                                  uint32_t* __restrict__ long_count,
                                                                  LongRun* __restrict__ warp_runs,
                                                                  uint32_t* __restrict__ warp_count) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= num_chunks) return;
    uint32_t k = pkeys[2 * t + 1];
    if (k == INVALID_KEY) return;
    // chunk t+1 starts with key k (the head is right-open); find the last chunk that does
    const size_t M = *src.total;
    const size_t live = (M + chunk - 1) / chunk;  // chunks that hold entries
    size_t lo = t + 1, hi = live;
    while (lo + 1 < hi) {
        size_t mid = lo + (hi - lo) / 2;
        if (tail_key<LEVEL0>(src, mid * chunk) <= k)
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
__device__ __forceinline__ xyzz_t shfl_down_xyzz(const xyzz_t& v, int delta) {
    xyzz_t o;
    o.X = shfl_down_fp(v.X, delta);
    o.Y = shfl_down_fp(v.Y, delta);
    o.ZZ = shfl_down_fp(v.ZZ, delta);
    o.ZZZ = shfl_down_fp(v.ZZZ, delta);
    return o;
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
            xyzz_t o = shfl_down_xyzz(acc, delta);
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
// K6: bucket reduction by bit planes.  With B_k the bucket of digit k + 1 (k < half = 2^L),
//     sum_k (k + 1) B_k = sum_k B_k + sum_{p < L} 2^p T_p,      T_p = sum_{k : bit p of k set} B_k.
// One pairwise tree carries the plane sums along: a level-k node covers 2^k consecutive buckets and holds
// k + 1 points [S, T_0 .. T_{k-1}] restricted to its range.  Merging the children (c0, c1) of a node:
//     S' = S(c0) + S(c1),   T_p' = T_p(c0) + T_p(c1) for p < k - 1,   T_{k-1}' = S(c1)
// (the upper child is exactly the set of buckets whose bit k - 1 is set).  One thread per (node, slot): every
// level is one launch of independent additions -- 2 unweighted additions per bucket in total, no doublings, no
// running-sum chains; the wide levels run at arithmetic throughput, each narrow level costs one addition of
// latency.  2^p is applied once, in the final Horner pass over the root's L plane sums.
// ------------------------------------------------------------------------------------------
// (msm_plane_tree_level_kernel, the wide levels: msm_tree.cu -- the one kernel of the MSM that is faster with its
// products inlined)

// The narrow top of the same tree in ONE block: from level k_first on, a level has at most 512 (node, slot) additions, so
// a launch per level is all latency (launch gap + one addition each).  The block keeps the levels in its ping-pong
// buffers and synchronises between them.
constexpr int TREE_TOP_THREADS = 512;
__global__ void __launch_bounds__(TREE_TOP_THREADS) msm_plane_tree_top_kernel(const xyzz_t* __restrict__ in, xyzz_t* __restrict__ buf_a,
                                                                               xyzz_t* __restrict__ buf_b, uint32_t k_first,
                                                                               uint32_t k_last, size_t nodes_first) {
    const uint32_t t = threadIdx.x;
    const xyzz_t* src = in;
    size_t nodes = nodes_first;
    for (uint32_t k = k_first; k <= k_last; k++) {
        xyzz_t* dst = (k & 1) ? buf_a : buf_b;
        const uint32_t slots = k + 1;
        if (t < nodes * slots) {
            const size_t node = t / slots;
            const uint32_t s = (uint32_t)(t - node * slots);
            const xyzz_t* c0 = src + 2 * node * k;
            const xyzz_t* c1 = c0 + k;
            if (s == k) {
                st_xyzz(dst + t, ld_xyzz(c1));
            } else {
                xyzz_t a = ld_xyzz(c0 + s);
                xyzz_t b = ld_xyzz(c1 + s);
                xyzz_add(a, b);
                st_xyzz(dst + t, a);
            }
        }
        __threadfence_block();
        __syncthreads();
        src = dst;
        nodes >>= 1;
    }
}

// ------------------------------------------------------------------------------------------
// K7: plane / window Horner + output
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

__global__ void msm_write_identity_kernel(uint64_t* out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) write_projective(out, xyzz_t::inf(), true);
}

// The L plane sums of a window are folded by Horner in 2.  With few windows (precomputed levels: one) the planes are
// split into groups of GROUP consecutive planes, one thread per group, and the group values are folded in 2^GROUP --
// the same doublings, but a third of the additions on the serial path.
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

// The same for at most four windows (precomputed levels: one), by one warp per window: lane p takes plane p and doubles
// it p times -- all lanes in step, so the serial path is L - 1 doublings instead of L doublings and L additions --, lane
// 31 takes the plain bucket sum, and a shuffle tree adds the lanes (5 additions).
__global__ void __launch_bounds__(128) msm_finalize_planes_warp_kernel(const xyzz_t* __restrict__ roots, uint32_t W,
                                                                        uint32_t L, uint32_t c, int normalise,
                                                                        uint64_t* out) {
    __shared__ xyzz_t win[4];
    const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const xyzz_t* root = roots + (size_t)w * (L + 1);   // [S, T_0 .. T_{L-1}], L <= 31
    xyzz_t acc = xyzz_t::inf();
    if (lane < L) acc = ld_xyzz(root + 1 + lane);
    else if (lane == 31) acc = ld_xyzz(root);
#pragma unroll 1
    for (uint32_t i = 0; i + 1 < L; i++)
        if (i < lane && lane < L) xyzz_dbl(acc);
    __syncwarp();
#pragma unroll 1
    for (int delta = 16; delta >= 1; delta >>= 1) {
        xyzz_t o = shfl_down_xyzz(acc, delta);
        xyzz_add(acc, o);
        __syncwarp();
    }
    if (lane == 0) win[w] = acc;
    __syncthreads();
    if (threadIdx.x != 0) return;
    acc = win[W - 1];
    for (int k = (int)W - 2; k >= 0; k--) {
        for (uint32_t i = 0; i < c; i++) xyzz_dbl(acc);
        xyzz_t v = win[k];
        xyzz_add(acc, v);
    }
    write_projective(out, acc, normalise != 0);
}

// sum of n homogeneous projective points (X:Y:Z), normalised output: one warp, lane-strided partial sums and a
// shuffle tree (the post-gather step of the sharded MSM adds one partial per rank)
__global__ void __launch_bounds__(32) g1_sum_kernel(const uint64_t* pts, uint32_t n, uint64_t* out) {
    const uint32_t lane = threadIdx.x;
    xyzz_t acc = xyzz_t::inf();
    for (uint32_t i = lane; i < n; i += 32) {
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
    __syncwarp();
#pragma unroll 1
    for (int delta = 16; delta >= 1; delta >>= 1) {
        if ((uint32_t)delta < n) {  // uniform across the warp
            xyzz_t o = shfl_down_xyzz(acc, delta);
            xyzz_add(acc, o);
        }
        __syncwarp();
    }
    if (lane == 0) write_projective(out, acc, true);
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
        // pair additions + bucket-tree additions (full adds ~2.3x a batched-affine add, 2 per bucket) + a latency
        // term for the serial depth of the tree / Horner tail expressed in pair-addition equivalents
        double cost = (double)W * ((double)n + 5.0 * (double)(1u << (c - 1))) + 2000.0 * c;
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
    int L;           // affine tree levels
};

static size_t level_ub(size_t M, uint32_t nb, int l) {  // slots of level l: entries + one pad per unfinished bucket
    const size_t e = M >> l;
    return e + 2 * (e < nb ? e : (size_t)nb) + 2;
}

// number of affine tree levels for M entries in nb buckets
static int msm_pick_levels(bpk_ctx* ctx, size_t M, uint32_t nb) {
    if (ctx->opt_msm_affine_levels >= 0)
        return (int)(ctx->opt_msm_affine_levels > MSM_MAX_LEVELS ? MSM_MAX_LEVELS : ctx->opt_msm_affine_levels);
    if (M < (size_t)ctx->opt_msm_min_pairs) return 0;
    // uniform digits: the fullest bucket holds about avg + 6 sqrt(avg) entries; one more level than its depth costs an
    // empty launch, one fewer leaves work to the (slower) XYZZ tail
    const double avg = (double)M / (double)nb;
    const double top = avg + 6.0 * sqrt(avg) + 1.0;
    int L = 0;
    while (L < MSM_MAX_LEVELS && (double)((size_t)1 << L) < top) L++;
    // a level with few pairs is all latency (one inversion per thread): leave those to the tail
    while (L > 0 && (M >> L) < (size_t)ctx->opt_msm_min_pairs) L--;
    return L;
}

static int msm_make_plan(bpk_ctx* ctx, const MsmPoints& pts, size_t n, MsmPlan* plan) {
    if (n >= ((size_t)1 << 31)) return BPK_ERR_TOO_LARGE;
    plan->pre = pts.pre_c != 0;
    plan->c = plan->pre ? pts.pre_c : pick_window(ctx, n);
    plan->W = plan->pre ? pts.pre_W : (256 + plan->c - 1) / plan->c;
    plan->half = 1u << (plan->c - 1);
    plan->WB = plan->pre ? 1 : plan->W;
    plan->nb_total = plan->WB * plan->half;
    if ((size_t)plan->W * n + plan->nb_total >= ((size_t)1 << 31)) return BPK_ERR_TOO_LARGE;
    if (plan->pre && (size_t)plan->W * pts.level_stride >= ((size_t)1 << 31)) return BPK_ERR_TOO_LARGE;
    plan->L = msm_pick_levels(ctx, (size_t)plan->W * n, plan->nb_total);
    // the tree's level buffers (96 B per slot at levels 1 and 2) must fit the budget; beyond it (2^26-point inputs next
    // to a precomputed SRS) everything goes through the XYZZ chunks, which need no per-entry storage
    if (plan->L >= 1) {
        const size_t M = (size_t)plan->W * n;
        const size_t bytes = (level_ub(M, plan->nb_total, 1) + (plan->L >= 2 ? level_ub(M, plan->nb_total, 2) : 0)) * 100;
        if (bytes > (size_t)ctx->opt_msm_level_mib << 20) plan->L = 0;
    }
    return BPK_OK;
}

static int msm_empty_result(bpk_ctx* ctx, uint64_t* d_out) {  // empty sum: identity (0, R, 0)
    msm_write_identity_kernel<<<1, 1, 0, ctx->stream>>>(d_out);
    count_launch(ctx);
    BPK_CUDA(cudaGetLastError());
    return BPK_OK;
}

static uint32_t msm_chunk(bpk_ctx* ctx, size_t M) {  // sorted entries per accumulate thread
    uint32_t chunk = (uint32_t)ctx->opt_msm_chunk;
    if (chunk == 0) {
        // about 2048 threads per SM, at most 256 entries each; then shrink the chunk so that the grid is a whole
        // number of waves of resident threads (threads take equal time, a nearly empty last wave costs a full one)
        size_t target = M / ((size_t)ctx->sm_count * 2048);
        target = target < 16 ? 16 : (target > 256 ? 256 : target);
        const size_t resident = (size_t)ctx->sm_count * BPK_ACC_MINBLOCKS * 128;
        const size_t waves = (M + target * resident - 1) / (target * resident);
        size_t c = (M + waves * resident - 1) / (waves * resident);
        chunk = (uint32_t)(c < 16 ? 16 : c);
    }
    return chunk;
}

// ---- workspace of one bucket fill ----
struct MsmWork {
    size_t M;                 // upper bound of the entries (W n)
    uint32_t nb, nblk;
    int L;
    uint32_t *cnt0, *cursor, *off, *blocksums, *totals;
    unsigned long long* stats;
    uint2* kv0;
    uint32_t* digits;         // [W][n] recoded windows for the phased scatter of large inputs (null: one-pass scatter)
    uint32_t phases;
    uint32_t* keys_lvl[2];    // [odd levels, even levels >= 2]
    affine_t* pts_lvl[2];
    uint4* scratch;
    uint32_t* rare_bits;        // flags of the degenerate pairs of the level in flight
    size_t rare_words;
    uint32_t aff_threads, bmax;
    size_t tail_ub;           // upper bound of the tail list
    uint32_t chunk;
    size_t num_chunks;
    uint32_t* pkeys;
    xyzz_t* pvals;
    LongRun *long_runs, *warp_runs;
    uint32_t *long_count, *warp_count;
};

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

// reserves (grow-only) and carves the buffers; growing later would synchronise the stream, so callers that run
// several fills back to back reserve for the largest first
static int msm_workspace(bpk_ctx* ctx, const MsmPlan& pl, size_t n, MsmWork* w) {
    w->M = (size_t)pl.W * n;
    w->nb = pl.nb_total;
    w->nblk = (w->nb + SCAN_TILE - 1) / SCAN_TILE;
    w->L = pl.L;
    const int L = pl.L;
    {   // tables
        const size_t b_cnt = align256((size_t)w->nb * 4), b_off = align256((size_t)(L + 1) * w->nb * 4);
        const size_t b_sums = align256((size_t)(L + 1) * w->nblk * 4), b_tot = 256, b_stats = 256;
        char* base;
        BPK_TRY(ws_reserve(ctx, 3, 2 * b_cnt + b_off + b_sums + b_tot + b_stats, (void**)&base));
        w->cnt0 = (uint32_t*)base;
        w->cursor = (uint32_t*)(base + b_cnt);
        w->off = (uint32_t*)(base + 2 * b_cnt);
        w->blocksums = (uint32_t*)(base + 2 * b_cnt + b_off);
        w->totals = (uint32_t*)(base + 2 * b_cnt + b_off + b_sums);
        w->stats = (unsigned long long*)(base + 2 * b_cnt + b_off + b_sums + b_tot);
    }
    const size_t kv_bytes = (level_ub(w->M, w->nb, 0) + 2) * sizeof(uint2);
    BPK_TRY(ws_reserve(ctx, 2, kv_bytes, (void**)&w->kv0));
    // the level-0 list is scattered in as many phases (bucket ranges) as it takes for one phase's slice to stay in L2
    w->digits = nullptr;
    w->phases = 1;
    const size_t slice = (size_t)(ctx->opt_msm_scatter_l2_mib < 1 ? 1 : ctx->opt_msm_scatter_l2_mib) << 20;
    if (kv_bytes > slice) {
        w->phases = (uint32_t)((kv_bytes + slice - 1) / slice);
        if (w->phases > 64) w->phases = 64;
        BPK_TRY(ws_reserve(ctx, 20, w->M * sizeof(uint32_t), (void**)&w->digits));
    }
    w->keys_lvl[0] = w->keys_lvl[1] = nullptr;
    w->pts_lvl[0] = w->pts_lvl[1] = nullptr;
    w->scratch = nullptr;
    w->rare_bits = nullptr;
    w->rare_words = 0;
    w->aff_threads = (uint32_t)ctx->sm_count * AFF_WARPS_PER_SM * 32;
    w->bmax = (uint32_t)(ctx->opt_msm_batch < 1 ? 1 : ctx->opt_msm_batch);
    if (L >= 1) {
        // + 66: the cooperative copies of the last warp of a level read up to 31 pairs past its end
        const size_t u1 = level_ub(w->M, w->nb, 1) + 66, u2 = L >= 2 ? level_ub(w->M, w->nb, 2) + 66 : 0;
        char* base;
        BPK_TRY(ws_reserve(ctx, 16, align256(u1 * 4) + align256(u2 * 4), (void**)&base));
        w->keys_lvl[0] = (uint32_t*)base;
        w->keys_lvl[1] = (uint32_t*)(base + align256(u1 * 4));
        BPK_TRY(ws_reserve(ctx, 17, u1 * sizeof(affine_t), (void**)&w->pts_lvl[0]));
        if (u2) BPK_TRY(ws_reserve(ctx, 18, u2 * sizeof(affine_t), (void**)&w->pts_lvl[1]));
        // batches never exceed what one round over the largest level needs
        const size_t pairs0 = level_ub(w->M, w->nb, 0) / 2;
        size_t need = (pairs0 + w->aff_threads - 1) / w->aff_threads;
        if (need < w->bmax) w->bmax = (uint32_t)(need < 1 ? 1 : need);
        BPK_TRY(ws_reserve(ctx, 19, (size_t)w->bmax * 3 * w->aff_threads * sizeof(uint4), (void**)&w->scratch));
        w->rare_words = pairs0 / 32 + 2 + MSM_MAX_LEVELS + 1;   // + one claim counter per level
        BPK_TRY(ws_reserve(ctx, 21, w->rare_words * sizeof(uint32_t), (void**)&w->rare_bits));
    }
    w->tail_ub = L == 0 ? w->M : level_ub(w->M, w->nb, L);
    w->chunk = msm_chunk(ctx, w->tail_ub);
    w->num_chunks = (w->tail_ub + w->chunk - 1) / w->chunk;
    {
        char* base;
        const size_t pv_bytes = 2 * w->num_chunks * sizeof(xyzz_t);
        BPK_TRY(ws_reserve(ctx, 5, pv_bytes + 2 * w->num_chunks * sizeof(uint32_t), (void**)&base));
        w->pvals = (xyzz_t*)base;
        w->pkeys = (uint32_t*)(base + pv_bytes);
    }
    {
        char* base;
        const size_t cap = w->num_chunks / MERGE_SERIAL_MAX + 2;  // a queued run covers > MERGE_SERIAL_MAX chunks
        BPK_TRY(ws_reserve(ctx, 11, 2 * cap * sizeof(LongRun) + 16, (void**)&base));
        w->long_count = (uint32_t*)base;
        w->warp_count = w->long_count + 1;
        w->long_runs = (LongRun*)(base + 16);
        w->warp_runs = w->long_runs + cap;
    }
    return BPK_OK;
}

template <bool LEVEL0>
static int msm_launch_tail_accumulate(bpk_ctx* ctx, const MsmWork& w, const TailSrc& src, xyzz_t* buckets) {
    const unsigned grid = (unsigned)((w.num_chunks + 127) / 128);
    msm_accumulate_kernel<LEVEL0><<<grid, 128, 0, ctx->stream>>>(src, w.chunk, w.num_chunks, buckets, w.pkeys, w.pvals);
    count_launch(ctx);
    BPK_CUDA(cudaGetLastError());
    return BPK_OK;
}
template <bool LEVEL0>
static int msm_launch_tail_merge(bpk_ctx* ctx, const MsmWork& w, const TailSrc& src, xyzz_t* buckets) {
    const unsigned grid = (unsigned)((w.num_chunks + 127) / 128);
    msm_merge_partials_kernel<LEVEL0><<<grid, 128, 0, ctx->stream>>>(w.pkeys, w.pvals, w.num_chunks, buckets, src, w.chunk,
                                                                    w.long_runs, w.long_count, w.warp_runs, w.warp_count);
    msm_merge_warp_runs_kernel<<<(unsigned)ctx->sm_count * 4, 128, 0, ctx->stream>>>(w.warp_runs, w.warp_count, w.pvals,
                                                                                     buckets);
    msm_merge_long_runs_kernel<<<(unsigned)ctx->sm_count * 2, MERGE_BLOCK, 0, ctx->stream>>>(w.long_runs, w.long_count,
                                                                                          w.pvals, buckets);
    count_launch(ctx, 3);
    BPK_CUDA(cudaGetLastError());
    return BPK_OK;
}

static int msm_configure_kernels(bpk_ctx* ctx) {  // opt in to > 48 KB of dynamic shared memory: per device, so per context
    if (ctx->msm_kernels_configured) return BPK_OK;
    BPK_CUDA(cudaFuncSetAttribute(msm_affine_level_kernel<true, AFF_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AFF_SMEM_BYTES));
    BPK_CUDA(cudaFuncSetAttribute(msm_affine_level_kernel<false, AFF_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AFF_SMEM_BYTES));
    BPK_CUDA(cudaFuncSetAttribute(msm_affine_level_kernel<true, AFF_THREADS_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AFF_SMEM_BYTES_SMALL));
    BPK_CUDA(cudaFuncSetAttribute(msm_affine_level_kernel<false, AFF_THREADS_SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AFF_SMEM_BYTES_SMALL));
    ctx->msm_kernels_configured = true;
    return BPK_OK;
}

// phase 1: count, layout, scatter, affine tree, tail: `buckets` (nb_total entries) receives the bucket sums of the n pairs
static int msm_fill_buckets(bpk_ctx* ctx, const MsmPlan& pl, const MsmPoints& pts, const fr_t* d_scalars, size_t n,
                            unsigned rshift, xyzz_t* buckets) {
    MsmWork w;
    BPK_TRY(msm_workspace(ctx, pl, n, &w));
    BPK_TRY(msm_configure_kernels(ctx));
    const int L = pl.L;
    ctx->last_c = pl.c;
    ctx->last_W = pl.W;
    ctx->last_chunk = w.chunk;
    ctx->last_buckets = pl.nb_total;
    ctx->last_levels = (unsigned)L;
    ctx->last_batch = w.bmax;
    ctx->last_stats_dev = w.stats;

    RecodeArgs ra;
    ra.scalars = d_scalars;
    ra.digits = w.digits;
    ra.n = (uint32_t)n;
    ra.c = pl.c;
    ra.W = pl.W;
    ra.rshift = rshift;
    ra.key_stride = pl.pre ? 0u : pl.half;
    ra.val_stride = pl.pre ? (uint32_t)pts.level_stride : 0u;
    const unsigned rgrid = (unsigned)((n + 255) / 256);
    {
        StageTimer t(ctx, "msm.recode");  // digits + histogram
        BPK_CUDA(cudaMemsetAsync(w.cnt0, 0, (size_t)w.nb * 4, ctx->stream));
        BPK_CUDA(cudaMemsetAsync(w.stats, 0, 32, ctx->stream));
        // many entries per bucket (few buckets): combine the lanes of a warp before every cursor / counter update
        const bool dense = w.M / w.nb > 256;
        if (dense)
            msm_count_kernel<true><<<rgrid, 256, 0, ctx->stream>>>(ra, w.cnt0);
        else
            msm_count_kernel<false><<<rgrid, 256, 0, ctx->stream>>>(ra, w.cnt0);
        count_launch(ctx);
        BPK_CUDA(cudaGetLastError());
        t.end();
    }
    {
        StageTimer t(ctx, "msm.sort");  // layout scan + scatter
        msm_level_sums_kernel<<<w.nblk, SCAN_THREADS, 0, ctx->stream>>>(w.cnt0, w.nb, L, w.nblk, w.blocksums, w.stats);
        msm_level_scan_kernel<<<L + 1, 1024, 0, ctx->stream>>>(w.blocksums, w.nblk, w.totals);
        msm_level_offsets_kernel<<<w.nblk, SCAN_THREADS, 0, ctx->stream>>>(w.cnt0, w.nb, L, w.nblk, w.blocksums, w.off,
                                                                        w.cursor, w.kv0);
        count_launch(ctx, 3);
        if (w.digits) {
            const uint32_t* skew = reinterpret_cast<const uint32_t*>(w.stats + 3);
            for (uint32_t ph = 0; ph < w.phases; ph++) {
                const uint32_t lo = (uint32_t)((uint64_t)w.nb * ph / w.phases), hi = (uint32_t)((uint64_t)w.nb * (ph + 1) / w.phases);
                const size_t threads = (w.M + RANGE_ITEMS - 1) / RANGE_ITEMS;
                msm_scatter_range_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(ra, w.cursor, skew, lo, hi,
                                                                                               w.kv0);
                count_launch(ctx);
            }
        } else {
            msm_scatter_kernel<<<rgrid, 256, 0, ctx->stream>>>(ra, w.cursor, reinterpret_cast<const uint32_t*>(w.stats + 3),
                                                               w.M / w.nb > 256 ? 1u : 0u, w.kv0);
            count_launch(ctx);
        }
        BPK_CUDA(cudaGetLastError());
        t.end();
    }
    TailSrc src;
    src.kv0 = L == 0 ? w.kv0 : nullptr;
    src.keys = L == 0 ? nullptr : w.keys_lvl[(L & 1) ? 0 : 1];
    src.pts = L == 0 ? pts.base : w.pts_lvl[(L & 1) ? 0 : 1];
    src.total = w.totals + L;
    {
        StageTimer t(ctx, "msm.accumulate");
        BPK_CUDA(cudaMemsetAsync(buckets, 0, (size_t)w.nb * sizeof(xyzz_t), ctx->stream));
        if (L > 0) BPK_CUDA(cudaMemsetAsync(w.rare_bits, 0, w.rare_words * sizeof(uint32_t), ctx->stream));
        for (int l = 0; l < L; l++) {
            AffLevelArgs a;
            a.kv0 = w.kv0;
            a.keys_in = l == 0 ? nullptr : w.keys_lvl[(l & 1) ? 0 : 1];
            a.pts_in = l == 0 ? pts.base : w.pts_lvl[(l & 1) ? 0 : 1];
            a.keys_out = w.keys_lvl[((l + 1) & 1) ? 0 : 1];
            a.pts_out = w.pts_lvl[((l + 1) & 1) ? 0 : 1];
            a.cnt0 = w.cnt0;
            a.off_in = w.off + (size_t)l * w.nb;
            a.off_out = w.off + (size_t)(l + 1) * w.nb;
            a.totals = w.totals;
            a.buckets = buckets;
            a.scratch = w.scratch;
            a.rare_bits = w.rare_bits;
            a.claim = w.rare_bits + (w.rare_words - MSM_MAX_LEVELS - 1) + l;
            a.level = (uint32_t)l;
            a.bmax = w.bmax;
            a.out_is_tail = l + 1 == L ? 1u : 0u;
            // the CTA shape by the level's expected size: large = a warp's even share is at least two full batches (where
            // the kernel claims guided batch sizes)
            const size_t pairs_ub = level_ub(w.M, w.nb, l) / 2;
            const bool large = ctx->opt_msm_cta_shape == 0 ? pairs_ub / (32 * (w.aff_threads / 32)) >= 2 * (size_t)w.bmax
                                                           : ctx->opt_msm_cta_shape == 1;
            const unsigned grid = w.aff_threads / (large ? AFF_THREADS : AFF_THREADS_SMALL);
            if (l == 0) {
                if (large) msm_affine_level_kernel<true, AFF_THREADS><<<grid, AFF_THREADS, AFF_SMEM_BYTES, ctx->stream>>>(a);
                else msm_affine_level_kernel<true, AFF_THREADS_SMALL><<<grid, AFF_THREADS_SMALL, AFF_SMEM_BYTES_SMALL, ctx->stream>>>(a);
                msm_affine_rare_kernel<true><<<(unsigned)ctx->sm_count * 4, 128, 0, ctx->stream>>>(a);
            } else {
                if (large) msm_affine_level_kernel<false, AFF_THREADS><<<grid, AFF_THREADS, AFF_SMEM_BYTES, ctx->stream>>>(a);
                else msm_affine_level_kernel<false, AFF_THREADS_SMALL><<<grid, AFF_THREADS_SMALL, AFF_SMEM_BYTES_SMALL, ctx->stream>>>(a);
                msm_affine_rare_kernel<false><<<(unsigned)ctx->sm_count * 4, 128, 0, ctx->stream>>>(a);
            }
            count_launch(ctx, 2);
        }
        // the XYZZ tail of the tree (everything, when no affine level runs)
        BPK_CUDA(cudaMemsetAsync(w.long_count, 0, 2 * sizeof(uint32_t), ctx->stream));
        if (L == 0)
            BPK_TRY(msm_launch_tail_accumulate<true>(ctx, w, src, buckets));
        else
            BPK_TRY(msm_launch_tail_accumulate<false>(ctx, w, src, buckets));
        t.end();
    }
    {
        StageTimer t(ctx, "msm.merge");  // partial sums of the runs that cross chunk borders
        if (L == 0)
            BPK_TRY(msm_launch_tail_merge<true>(ctx, w, src, buckets));
        else
            BPK_TRY(msm_launch_tail_merge<false>(ctx, w, src, buckets));
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
    const uint32_t c = pl.c, half = pl.half, WB = pl.WB;
    // bit-plane reduction (see K6): one tree whose nodes carry the plane sums, then Horner over the planes
    const uint32_t L = c - 1;  // half == 1 << L
    // ping-pong level buffers: level k holds WB * (half >> k) * (k + 1) points, largest at k = 1 and k = 2
    const size_t buf_a = (size_t)WB * half + TREE_TOP_THREADS;                    // odd levels  (k = 1: WB * half points)
    const size_t buf_b = (size_t)WB * (half / 4 + 1) * 3 + TREE_TOP_THREADS;      // even levels (k = 2: 3/4 WB * half points)
    xyzz_t* lvl;
    BPK_TRY(ws_reserve(ctx, 6, (buf_a + buf_b + 1) * sizeof(xyzz_t), (void**)&lvl));
    const xyzz_t* roots = buckets;  // L == 0: the single bucket of each window is its own root
    {
        StageTimer t(ctx, "msm.reduce");
        const xyzz_t* in = buckets;
        uint32_t k = 1;
        for (; k <= L; k++) {
            const size_t nodes = (size_t)WB * (half >> k);
            const size_t threads = nodes * (k + 1);
            // every later level is narrower still (nodes halve, slots grow by one): finish in one block
            if (ctx->opt_msm_tree_top && threads <= TREE_TOP_THREADS) break;
            xyzz_t* out = (k & 1) ? lvl : lvl + buf_a;
            msm_launch_plane_tree_level(ctx->stream, in, out, k, nodes);
            count_launch(ctx);
            in = out;
        }
        if (k <= L) {
            msm_plane_tree_top_kernel<<<1, TREE_TOP_THREADS, 0, ctx->stream>>>(in, lvl, lvl + buf_a, k, L,
                                                                                (size_t)WB * (half >> k));
            count_launch(ctx);
            in = (L & 1) ? lvl : lvl + buf_a;
        }
        roots = in;
        BPK_CUDA(cudaGetLastError());
        t.end();
    }
    {
        StageTimer t(ctx, "msm.finalize");
        uint32_t groups = (L + FIN_GROUP - 1) / FIN_GROUP;
        if (groups == 0 || WB * groups > 128) groups = 1;
        if (WB <= 4 && L >= 1 && L <= 31)
            msm_finalize_planes_warp_kernel<<<1, 32 * WB, 0, ctx->stream>>>(roots, WB, L, c, normalise ? 1 : 0, d_out);
        else
            msm_finalize_planes_kernel<<<1, 128, 0, ctx->stream>>>(roots, WB, L, c, groups, normalise ? 1 : 0, d_out);
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
        if (n) BPK_TRY(upload_host(ctx, d_stage, h_scalars, n * sizeof(fr_t), ctx->stream));
        return msm_run(ctx, pts, d_stage, n, rshift, normalise, d_out);
    }
    // both slices use the geometry (window, tree depth) of the larger one, and its workspace
    MsmPlan pl;
    BPK_TRY(msm_make_plan(ctx, pts, n - head, &pl));
    xyzz_t *buckets, *buckets_b;
    BPK_TRY(ws_reserve(ctx, 4, (size_t)pl.nb_total * sizeof(xyzz_t), (void**)&buckets));
    BPK_TRY(ws_reserve(ctx, 15, (size_t)pl.nb_total * sizeof(xyzz_t), (void**)&buckets_b));
    {
        MsmWork w;
        BPK_TRY(msm_workspace(ctx, pl, n - head, &w));
    }
    if (!ctx->copy_stream) BPK_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!ctx->copy_done) BPK_CUDA(cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming));
    if (!ctx->lane_fork) BPK_CUDA(cudaEventCreateWithFlags(&ctx->lane_fork, cudaEventDisableTiming));
    // the staging buffer may still be read by earlier work on the main stream
    BPK_CUDA(cudaEventRecord(ctx->lane_fork, ctx->stream));
    BPK_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->lane_fork, 0));
    BPK_TRY(upload_host(ctx, d_stage, h_scalars, head * sizeof(fr_t), ctx->stream));
    BPK_TRY(msm_fill_buckets(ctx, pl, pts, d_stage, head, rshift, buckets));
    BPK_TRY(upload_host(ctx, d_stage + head, h_scalars + 4 * head, (n - head) * sizeof(fr_t), ctx->copy_stream));
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
    g1_sum_kernel<<<1, 32, 0, ctx->stream>>>(d_points_xyz, (uint32_t)n, d_out_xyz);
    count_launch(ctx);
    BPK_CUDA(cudaGetLastError());
    t.end();
    return BPK_OK;
}

// counters of the most recent bucket fill: {entries, entries left to the XYZZ tail, non-empty buckets, 0}
int msm_read_stats(bpk_ctx* ctx, uint64_t out[4]) {
    out[0] = out[1] = out[2] = out[3] = 0;
    if (!ctx->last_stats_dev) return BPK_OK;
    BPK_CUDA(cudaStreamSynchronize(ctx->stream));
    BPK_CUDA(cudaMemcpy(out, ctx->last_stats_dev, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return BPK_OK;
}

}  // namespace bpk
