/* OPTIMISED CPU BASELINE + large-size checker (test infrastructure only; same rules as ref_cpu.c: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load liboracle.so).
 *
 * ref_cpu.c restates the REFERENCE's algorithms (naive O(n^2) DFT, 64 x 15-bucket MSM, single thread).  This file is
 * what BASELINE.md 3.2 asks for next to it: the textbook fast algorithms on every host core, so that the GPU numbers
 * are not quoted against a toy only --
 *   oracle_fr_ntt_fast      iterative radix-2 Cooley-Tukey (bit reversal + log2 n butterfly stages), pthreads;
 *                           same contract as ntt_381 / i_ntt_381 (src/utils.rs:63-129): natural order in and out,
 *                           w = ROOT_OF_UNITY^(2^32 / n), inverse scales by n^-1
 *   oracle_msm_pippenger_mt signed-digit Pippenger, window c ~ log2(n) - 3, XYZZ buckets with affine addends,
 *                           running-sum bucket reduction, point range split over the threads;
 *                           same value as BucketMSM::bucket_msm(points, scalars, 256, 4) (src/msm.rs:76-118)
 * and the O(n log n) pieces the trapdoor verifier check of 2^20+-gate proofs needs (tests/, bench.py):
 *   oracle_fr_poly_at       p(x) for p given by its values on the n-th roots of unity (inverse NTT, then Horner).
 * The field and curve arithmetic is ref_cpu.c's (u128 Montgomery as fp.rs / scalar.rs), included textually.
 */
#include "ref_cpu.c"

/* ------------------------------------------------------------------ thread helper */
typedef struct {
    void (*fn)(void*, int, int);
    void* arg;
    int tid, nthreads;
} par_job;
static void* par_tramp(void* p) {
    par_job* j = (par_job*)p;
    j->fn(j->arg, j->tid, j->nthreads);
    return NULL;
}
static void par_run(void (*fn)(void*, int, int), void* arg, int threads) {
    if (threads <= 1) {
        fn(arg, 0, 1);
        return;
    }
    pthread_t* th = (pthread_t*)calloc(threads, sizeof(pthread_t));
    par_job* jobs = (par_job*)calloc(threads, sizeof(par_job));
    for (int t = 0; t < threads; t++) {
        jobs[t].fn = fn;
        jobs[t].arg = arg;
        jobs[t].tid = t;
        jobs[t].nthreads = threads;
        pthread_create(&th[t], NULL, par_tramp, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

/* ------------------------------------------------------------------ radix-2 NTT */
static inline fr fr_sub(fr a, fr b) { fr r; f_sub(&FR, r.l, a.l, b.l); return r; }

typedef struct {
    fr* data;
    const fr* tw; /* tw[k] = w^k, k < n/2 */
    size_t n, half; /* current stage: butterflies span `half` */
    fr scale;
    int do_scale;
} ntt_stage;

static void ntt_stage_worker(void* arg, int tid, int nt) {
    ntt_stage* s = (ntt_stage*)arg;
    const size_t nb = s->n / 2; /* butterflies per stage */
    const size_t lo = nb * tid / nt, hi = nb * (tid + 1) / nt;
    const size_t half = s->half, stride = (s->n / 2) / half;
    for (size_t b = lo; b < hi; b++) {
        size_t grp = b / half, j = b % half;
        size_t i0 = grp * 2 * half + j, i1 = i0 + half;
        fr u = s->data[i0];
        fr v = fr_mul(s->data[i1], s->tw[j * stride]);
        s->data[i0] = fr_add(u, v);
        s->data[i1] = fr_sub(u, v);
    }
}
static void ntt_scale_worker(void* arg, int tid, int nt) {
    ntt_stage* s = (ntt_stage*)arg;
    const size_t lo = s->n * tid / nt, hi = s->n * (tid + 1) / nt;
    for (size_t i = lo; i < hi; i++) s->data[i] = fr_mul(s->data[i], s->scale);
}

int oracle_fr_ntt_fast(const u64* in, u64* out, size_t n, int inverse, int threads) {
    if (n == 0 || (n & (n - 1)) || n > ((size_t)1 << 32)) return -1;
    int lg = 0;
    while (((size_t)1 << lg) < n) lg++;
    fr* d = (fr*)out;
    const fr* src = (const fr*)in;
    /* bit-reversal permutation (in may alias out) */
    if (in != out) {
        for (size_t i = 0; i < n; i++) {
            size_t r = 0;
            for (int b = 0; b < lg; b++) r |= ((i >> b) & 1) << (lg - 1 - b);
            d[r] = src[i];
        }
    } else {
        for (size_t i = 0; i < n; i++) {
            size_t r = 0;
            for (int b = 0; b < lg; b++) r |= ((i >> b) & 1) << (lg - 1 - b);
            if (r > i) {
                fr t = d[i];
                d[i] = d[r];
                d[r] = t;
            }
        }
    }
    if (n == 1) return 0;
    /* w = ROOT^(2^32 / n) by repeated squaring (utils.rs:39-43) */
    fr w = inverse ? ROOT_OF_UNITY_INV : ROOT_OF_UNITY;
    for (int i = lg; i < 32; i++) w = fr_mul(w, w);
    fr* tw = (fr*)malloc(sizeof(fr) * (n / 2));
    tw[0] = fr_one();
    for (size_t k = 1; k < n / 2; k++) tw[k] = fr_mul(tw[k - 1], w);
    ntt_stage st;
    st.data = d;
    st.tw = tw;
    st.n = n;
    st.do_scale = 0;
    for (size_t half = 1; half < n; half <<= 1) {
        st.half = half;
        par_run(ntt_stage_worker, &st, n >= 4096 ? threads : 1);
    }
    if (inverse) {
        st.scale = fr_invert(fr_from_u64(n));
        par_run(ntt_scale_worker, &st, n >= 4096 ? threads : 1);
    }
    free(tw);
    return 0;
}

/* p(x) where `values` are p's evaluations on the n-th roots of unity (Lagrange basis, polynomial.rs:47-55):
 * coefficients by the fast inverse transform, then Horner.  x and the result are Montgomery. */
int oracle_fr_poly_at(const u64* values, size_t n, const u64* x, int threads, u64* out) {
    u64* coeffs = (u64*)malloc(sizeof(fr) * (n ? n : 1));
    int rc = oracle_fr_ntt_fast(values, coeffs, n, 1, threads);
    if (rc == 0) oracle_fr_horner(coeffs, n, x, out);
    free(coeffs);
    return rc;
}

/* ------------------------------------------------------------------ Pippenger MSM */
typedef struct { fp X, Y, ZZ, ZZZ; } xyzz;
static inline int fp_is_zero(fp a) { return f_is_zero(&FP, a.l); }
static inline fp fp_neg(fp a) { return fp_sub(fp_zero(), a); }
static inline fp fp_dbl(fp a) { return fp_add(a, a); }
static inline int fp_eq(fp a, fp b) { return memcmp(a.l, b.l, 48) == 0; }
static inline xyzz xyzz_inf(void) { xyzz r; memset(&r, 0, sizeof r); return r; }

/* EFD dbl-2008-s-1 (a = 0) */
static xyzz xyzz_dbl(xyzz p) {
    if (fp_is_zero(p.ZZ)) return p;
    fp U = fp_dbl(p.Y), V = fp_sqr(U), W = fp_mul(U, V), S = fp_mul(p.X, V);
    fp X2 = fp_sqr(p.X), M = fp_add(fp_dbl(X2), X2);
    xyzz r;
    r.X = fp_sub(fp_sqr(M), fp_dbl(S));
    r.Y = fp_sub(fp_mul(M, fp_sub(S, r.X)), fp_mul(W, p.Y));
    r.ZZ = fp_mul(V, p.ZZ);
    r.ZZZ = fp_mul(W, p.ZZZ);
    return r;
}
/* EFD add-2008-s, with the degenerate cases */
static xyzz xyzz_add(xyzz p, xyzz q) {
    if (fp_is_zero(q.ZZ)) return p;
    if (fp_is_zero(p.ZZ)) return q;
    fp U1 = fp_mul(p.X, q.ZZ), U2 = fp_mul(q.X, p.ZZ), S1 = fp_mul(p.Y, q.ZZZ), S2 = fp_mul(q.Y, p.ZZZ);
    fp Pd = fp_sub(U2, U1), Rd = fp_sub(S2, S1);
    if (fp_is_zero(Pd)) return fp_is_zero(Rd) ? xyzz_dbl(p) : xyzz_inf();
    fp PP = fp_sqr(Pd), PPP = fp_mul(Pd, PP), Qv = fp_mul(U1, PP);
    xyzz r;
    r.X = fp_sub(fp_sub(fp_sqr(Rd), PPP), fp_dbl(Qv));
    r.Y = fp_sub(fp_mul(Rd, fp_sub(Qv, r.X)), fp_mul(S1, PPP));
    r.ZZ = fp_mul(fp_mul(p.ZZ, q.ZZ), PP);
    r.ZZZ = fp_mul(fp_mul(p.ZZZ, q.ZZZ), PPP);
    return r;
}
/* EFD madd-2008-s: q affine (x, y), not infinity */
static xyzz xyzz_madd(xyzz p, fp qx, fp qy) {
    if (fp_is_zero(p.ZZ)) {
        xyzz r = {qx, qy, fp_one(), fp_one()};
        return r;
    }
    fp U2 = fp_mul(qx, p.ZZ), S2 = fp_mul(qy, p.ZZZ), Pd = fp_sub(U2, p.X), Rd = fp_sub(S2, p.Y);
    if (fp_is_zero(Pd)) {
        if (!fp_is_zero(Rd)) return xyzz_inf();
        xyzz a = {qx, qy, fp_one(), fp_one()};
        return xyzz_dbl(a);
    }
    fp PP = fp_sqr(Pd), PPP = fp_mul(Pd, PP), Qv = fp_mul(p.X, PP);
    xyzz r;
    r.X = fp_sub(fp_sub(fp_sqr(Rd), PPP), fp_dbl(Qv));
    r.Y = fp_sub(fp_mul(Rd, fp_sub(Qv, r.X)), fp_mul(p.Y, PPP));
    r.ZZ = fp_mul(p.ZZ, PP);
    r.ZZZ = fp_mul(p.ZZZ, PPP);
    return r;
}
/* XYZZ (x = X/ZZ, y = Y/ZZZ) -> homogeneous projective (X ZZZ : Y ZZ : ZZ ZZZ) */
static g1p xyzz_to_proj(xyzz p) {
    if (fp_is_zero(p.ZZ)) return g1_identity();
    g1p r = {fp_mul(p.X, p.ZZZ), fp_mul(p.Y, p.ZZ), fp_mul(p.ZZ, p.ZZZ)};
    return r;
}

typedef struct {
    const g1p* points; /* Z == R (normalised) or Z == 0 (identity) */
    const fr* scalars;
    size_t n;
    int c;
    g1p* partial; /* one per thread */
    int bad_input;
} pip_job;

static void pip_worker(void* arg, int tid, int nt) {
    pip_job* J = (pip_job*)arg;
    const size_t lo = J->n * tid / nt, hi = J->n * (tid + 1) / nt, m = hi - lo;
    const int c = J->c, W = (255 + c) / c + 1; /* +1: the carry of the signed recoding */
    const size_t half = (size_t)1 << (c - 1);
    /* signed digits of every scalar of the slice */
    int32_t* digits = (int32_t*)malloc(sizeof(int32_t) * m * W);
    const fp one = fp_one();
    for (size_t i = 0; i < m; i++) {
        uint8_t bytes[40];
        memset(bytes, 0, sizeof bytes);
        fr_to_bytes(J->scalars[lo + i], bytes);
        int carry = 0;
        for (int w = 0; w < W; w++) {
            size_t bit = (size_t)w * c;
            u64 chunk = 0;
            if (bit < 256) {
                memcpy(&chunk, bytes + bit / 8, 8); /* c <= 24: a window spans at most 4 bytes after the shift */
                chunk = (chunk >> (bit % 8)) & (((u64)1 << c) - 1);
            }
            int64_t d = (int64_t)chunk + carry;
            if ((size_t)d > half) {
                d -= (int64_t)1 << c;
                carry = 1;
            } else
                carry = 0;
            digits[i * W + w] = (int32_t)d;
        }
        const g1p* P = &J->points[lo + i];
        if (!fp_is_zero(P->z) && !fp_eq(P->z, one)) J->bad_input = 1;
    }
    xyzz* buckets = (xyzz*)malloc(sizeof(xyzz) * half);
    xyzz total = xyzz_inf();
    for (int w = W - 1; w >= 0; w--) {
        for (int k = 0; k < c; k++) total = xyzz_dbl(total);
        memset(buckets, 0, sizeof(xyzz) * half);
        for (size_t i = 0; i < m; i++) {
            int32_t d = digits[i * W + w];
            const g1p* P = &J->points[lo + i];
            if (d == 0 || fp_is_zero(P->z)) continue;
            if (d > 0)
                buckets[d - 1] = xyzz_madd(buckets[d - 1], P->x, P->y);
            else
                buckets[-d - 1] = xyzz_madd(buckets[-d - 1], P->x, fp_neg(P->y));
        }
        xyzz run = xyzz_inf(), sum = xyzz_inf();
        for (size_t b = half; b-- > 0;) {
            run = xyzz_add(run, buckets[b]);
            sum = xyzz_add(sum, run);
        }
        total = xyzz_add(total, sum);
    }
    J->partial[tid] = xyzz_to_proj(total);
    free(buckets);
    free(digits);
}

/* returns 0; -1 on bad arguments; -2 if a point is neither normalised (Z == R) nor the identity */
int oracle_msm_pippenger_mt(const u64* points_xyz, size_t n_points, const u64* scalars, size_t n_scalars, int threads,
                            int window, u64* out_xyz) {
    size_t n = n_points < n_scalars ? n_points : n_scalars;
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    size_t per = n / threads + 1;
    int c = window;
    if (c <= 0) {
        int lg = 0;
        while (((size_t)2 << lg) <= per) lg++;
        c = lg - 3;
    }
    if (c < 2) c = 2;
    if (c > 20) c = 20;
    pip_job J;
    J.points = (const g1p*)points_xyz;
    J.scalars = (const fr*)scalars;
    J.n = n;
    J.c = c;
    J.bad_input = 0;
    J.partial = (g1p*)calloc(threads, sizeof(g1p));
    par_run(pip_worker, &J, threads);
    g1p acc = g1_identity();
    for (int t = 0; t < threads; t++) acc = g1_add(&acc, &J.partial[t]);
    memcpy(out_xyz, &acc, sizeof acc);
    free(J.partial);
    return J.bad_input ? -2 : 0;
}
