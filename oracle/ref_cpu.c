/* CPU ORACLE in C (test infrastructure only): a restatement of the reference's own algorithms for
 * the MSM / NTT hot path, operation for operation, used (a) to cross-check the Python oracle at
 * sizes Python is too slow for and (b) as the timed CPU baseline ("port") in bench.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * liboracle.so; the product (baby-plonk-rust_b200/) never does.
 *
 * It is a restatement in C, NOT rustc output: the reference's Rust toolchain is absent from this
 * image.  Parity: pinned -- tests/test_oracle_c.py checks it against the reference's
 * Fp / Fr known-answer vectors (tests/golden/reference_kats.json) and against oracle/bls12_381.py,
 * which is itself pinned against the 1000-entry G1 vectors.
 *
 * What is restated, with the reference lines (paths relative to the reference root):
 *   mac/adc/sbb                       lib/bls12_381/src/util.rs:3-20
 *   Fp mul / montgomery_reduce        lib/bls12_381/src/fp.rs:565-609, 487-562 (schoolbook + HAC 14.32)
 *   Fp add / sub / neg / subtract_p   lib/bls12_381/src/fp.rs:361-423
 *   Scalar mul / montgomery_reduce    lib/bls12_381/src/scalar.rs:562-586, 514-558
 *   Scalar add / sub / pow / to_bytes lib/bls12_381/src/scalar.rs:590-635, 381-392, 292-304
 *   G1Projective add (Alg 7), double (Alg 9), mul_by_3b   lib/bls12_381/src/g1.rs:597-712
 *   BucketMSM::bucket_msm, c_bit_msm, get_c_bit_chunk     src/msm.rs:23-139
 *   ntt_381 / i_ntt_381 (naive DFT, n x n matrix, pow per term)   src/utils.rs:63-129
 * The reference is single-threaded; oracle_bucket_msm_mt splits the pair list over T threads and adds
 * the partial results, for a baseline that may use every host core.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

/* util.rs:3-20 */
static inline u64 adc(u64 a, u64 b, u64 carry, u64* carry_out) {
    u128 r = (u128)a + b + carry;
    *carry_out = (u64)(r >> 64);
    return (u64)r;
}
static inline u64 sbb(u64 a, u64 b, u64 borrow, u64* borrow_out) {
    u128 r = (u128)a - b - (borrow >> 63);
    *borrow_out = (u64)(r >> 64);
    return (u64)r;
}
static inline u64 mac(u64 a, u64 b, u64 c, u64 carry, u64* carry_out) {
    u128 r = (u128)a + (u128)b * c + carry;
    *carry_out = (u64)(r >> 64);
    return (u64)r;
}

/* ------------------------------------------------------------------ generic N-limb Montgomery field */
typedef struct {
    int n;
    u64 mod[6];
    u64 inv;
    u64 one[6]; /* R mod p */
} field_t;

static const field_t FP = {6,
                           {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull,
                            0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull},
                           0x89f3fffcfffcfffdull,
                           {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull, 0x77ce585370525745ull,
                            0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull}};
static const field_t FR = {4,
                           {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull, 0, 0},
                           0xfffffffeffffffffull,
                           {0x00000001fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full, 0, 0}};

/* subtract_p / final conditional subtraction (fp.rs:361-383, scalar.rs:555-557) */
static inline void f_sub_mod_if_ge(const field_t* F, u64* r) {
    u64 t[6], borrow = 0;
    for (int i = 0; i < F->n; i++) t[i] = sbb(r[i], F->mod[i], borrow, &borrow);
    if (!(borrow >> 63))
        for (int i = 0; i < F->n; i++) r[i] = t[i];
}
static inline void f_add(const field_t* F, u64* r, const u64* a, const u64* b) {
    u64 carry = 0;
    for (int i = 0; i < F->n; i++) r[i] = adc(a[i], b[i], carry, &carry);
    f_sub_mod_if_ge(F, r);
}
static inline void f_sub(const field_t* F, u64* r, const u64* a, const u64* b) {
    u64 borrow = 0, t[6];
    for (int i = 0; i < F->n; i++) t[i] = sbb(a[i], b[i], borrow, &borrow);
    u64 mask = (borrow >> 63) ? ~(u64)0 : 0, carry = 0;
    for (int i = 0; i < F->n; i++) r[i] = adc(t[i], F->mod[i] & mask, carry, &carry);
}
/* schoolbook product followed by montgomery_reduce (fp.rs:565-609 + 487-562) */
static inline void f_mul(const field_t* F, u64* r, const u64* a, const u64* b) {
    const int n = F->n;
    u64 t[13];
    memset(t, 0, sizeof t);
    for (int i = 0; i < n; i++) {
        u64 carry = 0;
        for (int j = 0; j < n; j++) t[i + j] = mac(t[i + j], a[i], b[j], carry, &carry);
        t[i + n] = carry;
    }
    u64 carry2 = 0;
    for (int i = 0; i < n; i++) {
        u64 k = t[i] * F->inv, carry = 0;
        (void)mac(t[i], k, F->mod[0], 0, &carry);
        for (int j = 1; j < n; j++) t[i + j] = mac(t[i + j], k, F->mod[j], carry, &carry);
        t[i + n] = adc(t[i + n], carry2, carry, &carry2);
    }
    u64 out[6];
    for (int i = 0; i < n; i++) out[i] = t[i + n];
    f_sub_mod_if_ge(F, out);
    for (int i = 0; i < n; i++) r[i] = out[i];
}
static inline int f_is_zero(const field_t* F, const u64* a) {
    u64 acc = 0;
    for (int i = 0; i < F->n; i++) acc |= a[i];
    return acc == 0;
}

/* ------------------------------------------------------------------ Fp / Fr front-ends */
typedef struct { u64 l[6]; } fp;
typedef struct { u64 l[4]; } fr;
static inline fp fp_add(fp a, fp b) { fp r; f_add(&FP, r.l, a.l, b.l); return r; }
static inline fp fp_sub(fp a, fp b) { fp r; f_sub(&FP, r.l, a.l, b.l); return r; }
static inline fp fp_mul(fp a, fp b) { fp r; f_mul(&FP, r.l, a.l, b.l); return r; }
static inline fp fp_sqr(fp a) { return fp_mul(a, a); }
static inline fr fr_add(fr a, fr b) { fr r; f_add(&FR, r.l, a.l, b.l); return r; }
static inline fr fr_mul(fr a, fr b) { fr r; f_mul(&FR, r.l, a.l, b.l); return r; }
static inline fr fr_one(void) { fr r; memcpy(r.l, FR.one, 32); return r; }
static inline fp fp_one(void) { fp r; memcpy(r.l, FP.one, 48); return r; }
static inline fp fp_zero(void) { fp r; memset(r.l, 0, 48); return r; }

/* Scalar::pow (scalar.rs:381-392): all 256 bits, always square and multiply, conditional assign */
static fr fr_pow(fr base, const u64 by[4]) {
    fr res = fr_one();
    for (int w = 3; w >= 0; w--)
        for (int i = 63; i >= 0; i--) {
            res = fr_mul(res, res);
            fr tmp = fr_mul(res, base);
            if ((by[w] >> i) & 1) res = tmp;
        }
    return res;
}
/* Scalar::invert: the reference uses a fixed addition chain (scalar.rs:416-511, ~300 multiplications);
 * restated as a^(q-2) by square-and-multiply (255 squarings + 128 multiplications) */
static fr fr_invert(fr a) {
    u64 e[4] = {FR.mod[0] - 2, FR.mod[1], FR.mod[2], FR.mod[3]};
    fr res = fr_one();
    for (int w = 3; w >= 0; w--)
        for (int i = 63; i >= 0; i--) {
            res = fr_mul(res, res);
            if ((e[w] >> i) & 1) res = fr_mul(res, a);
        }
    return res;
}
/* Scalar::to_bytes (scalar.rs:292-304): Montgomery reduce to canonical, little-endian bytes */
static void fr_to_bytes(fr a, uint8_t out[32]) {
    fr one_raw = {{1, 0, 0, 0}};
    fr c = fr_mul(a, one_raw);
    memcpy(out, c.l, 32);
}
static fr fr_from_u64(u64 v) { /* Scalar::from(u64): v * R2 reduced == to Montgomery */
    static const fr R2 = {{0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull}};
    fr x = {{v, 0, 0, 0}};
    return fr_mul(x, R2);
}

/* ------------------------------------------------------------------ G1Projective (g1.rs) */
typedef struct { fp x, y, z; } g1p;
static inline g1p g1_identity(void) { g1p r; r.x = fp_zero(); r.y = fp_one(); r.z = fp_zero(); return r; }
static inline fp mul_by_3b(fp a) { /* g1.rs:597-601 */
    a = fp_add(a, a);
    a = fp_add(a, a);
    return fp_add(fp_add(a, a), a);
}
/* g1.rs:638-667, Algorithm 9 of ePrint 2015/1060 */
static g1p g1_double(const g1p* p) {
    fp t0 = fp_sqr(p->y);
    fp z3 = fp_add(t0, t0);
    z3 = fp_add(z3, z3);
    z3 = fp_add(z3, z3);
    fp t1 = fp_mul(p->y, p->z);
    fp t2 = fp_sqr(p->z);
    t2 = mul_by_3b(t2);
    fp x3 = fp_mul(t2, z3);
    fp y3 = fp_add(t0, t2);
    z3 = fp_mul(t1, z3);
    t1 = fp_add(t2, t2);
    t2 = fp_add(t1, t2);
    t0 = fp_sub(t0, t2);
    y3 = fp_mul(t0, y3);
    y3 = fp_add(x3, y3);
    t1 = fp_mul(p->x, p->y);
    x3 = fp_mul(t0, t1);
    x3 = fp_add(x3, x3);
    g1p r = {x3, y3, z3};
    if (f_is_zero(&FP, p->z.l)) return g1_identity();
    return r;
}
/* g1.rs:670-712, Algorithm 7 */
static g1p g1_add(const g1p* a, const g1p* b) {
    fp t0 = fp_mul(a->x, b->x);
    fp t1 = fp_mul(a->y, b->y);
    fp t2 = fp_mul(a->z, b->z);
    fp t3 = fp_add(a->x, a->y);
    fp t4 = fp_add(b->x, b->y);
    t3 = fp_mul(t3, t4);
    t4 = fp_add(t0, t1);
    t3 = fp_sub(t3, t4);
    t4 = fp_add(a->y, a->z);
    fp x3 = fp_add(b->y, b->z);
    t4 = fp_mul(t4, x3);
    x3 = fp_add(t1, t2);
    t4 = fp_sub(t4, x3);
    x3 = fp_add(a->x, a->z);
    fp y3 = fp_add(b->x, b->z);
    x3 = fp_mul(x3, y3);
    y3 = fp_add(t0, t2);
    y3 = fp_sub(x3, y3);
    x3 = fp_add(t0, t0);
    t0 = fp_add(x3, t0);
    t2 = mul_by_3b(t2);
    fp z3 = fp_add(t1, t2);
    t1 = fp_sub(t1, t2);
    y3 = mul_by_3b(y3);
    x3 = fp_mul(t4, y3);
    t2 = fp_mul(t3, t1);
    x3 = fp_sub(t2, x3);
    y3 = fp_mul(y3, t0);
    t1 = fp_mul(t1, z3);
    y3 = fp_add(t1, y3);
    t0 = fp_mul(t0, t3);
    z3 = fp_mul(z3, t4);
    z3 = fp_add(z3, t0);
    g1p r = {x3, y3, z3};
    return r;
}

/* ------------------------------------------------------------------ src/msm.rs */
/* get_c_bit_chunk (msm.rs:119-139): to_bytes, reverse, explode into a heap Vec<bool>, slice, fold */
static u64 get_c_bit_chunk(const fr* scalar, size_t chunk_index, size_t chunk_size) {
    size_t start_bit = chunk_index * chunk_size, end_bit = start_bit + chunk_size;
    uint8_t bytes[32], rev[32];
    fr_to_bytes(*scalar, bytes);
    for (int i = 0; i < 32; i++) rev[i] = bytes[31 - i];
    uint8_t* bits = (uint8_t*)malloc(256); /* u8_to_bool_array allocates a Vec (msm.rs:64-75) */
    for (int i = 0; i < 32; i++)
        for (int j = 0; j < 8; j++) bits[8 * i + j] = (rev[i] >> (7 - j)) & 1;
    uint8_t* chunk = (uint8_t*)malloc(chunk_size ? chunk_size : 1); /* .to_vec() (msm.rs:133) */
    memcpy(chunk, bits + start_bit, end_bit - start_bit);
    u64 res = 0;
    for (size_t i = 0; i < chunk_size; i++)
        if (chunk[i]) res |= (u64)1 << (chunk_size - 1 - i);
    free(chunk);
    free(bits);
    return res;
}
/* c_bit_msm (msm.rs:23-49) */
static g1p c_bit_msm(const g1p* points, size_t n_points, const u64* digits, size_t n_digits, size_t c) {
    size_t nb = ((size_t)1 << c) - 1;
    g1p* buckets = (g1p*)malloc(sizeof(g1p) * (nb ? nb : 1));
    for (size_t i = 0; i < nb; i++) buckets[i] = g1_identity();
    size_t n = n_points < n_digits ? n_points : n_digits; /* zip */
    for (size_t i = 0; i < n; i++)
        if (digits[i] != 0) buckets[digits[i] - 1] = g1_add(&buckets[digits[i] - 1], &points[i]);
    g1p acc = g1_identity(), res = g1_identity();
    for (size_t i = nb; i-- > 0;) {
        acc = g1_add(&acc, &buckets[i]);
        res = g1_add(&res, &acc);
    }
    free(buckets);
    return res;
}
/* bucket_msm (msm.rs:76-118).  returns 0, or -1 where the reference panics */
int oracle_bucket_msm(const u64* points_xyz, size_t n_points, const u64* scalars, size_t n_scalars, size_t b,
                      size_t c, u64* out_xyz) {
    if (c == 0 || c > 24) return -1;
    size_t k = b / c;
    if (k == 0 || k * c > 256) return -1;
    const g1p* points = (const g1p*)points_xyz;
    const fr* sc = (const fr*)scalars;
    g1p* t_points = (g1p*)malloc(sizeof(g1p) * k);
    for (size_t i = 0; i < k; i++) {
        u64* digits = (u64*)calloc(n_scalars ? n_scalars : 1, sizeof(u64)); /* vec![0; scalars.len()] (msm.rs:91) */
        for (size_t j = 0; j < n_scalars; j++) digits[j] = get_c_bit_chunk(&sc[j], i, c);
        t_points[i] = c_bit_msm(points, n_points, digits, n_scalars, c);
        free(digits);
    }
    g1p result = t_points[0];
    for (size_t j = 1; j < k; j++) {
        for (size_t d = 0; d < c; d++) result = g1_double(&result);
        result = g1_add(&result, &t_points[j]);
    }
    free(t_points);
    memcpy(out_xyz, &result, sizeof(g1p));
    return 0;
}

typedef struct {
    const u64 *points, *scalars;
    size_t n, b, c;
    g1p out;
    int rc;
} msm_job;
static void* msm_worker(void* arg) {
    msm_job* j = (msm_job*)arg;
    j->rc = oracle_bucket_msm(j->points, j->n, j->scalars, j->n, j->b, j->c, (u64*)&j->out);
    return NULL;
}
/* the same algorithm on `threads` slices of the pair list, partial results added */
int oracle_bucket_msm_mt(const u64* points_xyz, size_t n_points, const u64* scalars, size_t n_scalars, size_t b,
                         size_t c, int threads, u64* out_xyz) {
    size_t n = n_points < n_scalars ? n_points : n_scalars;
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    msm_job* jobs = (msm_job*)calloc(threads, sizeof(msm_job));
    pthread_t* th = (pthread_t*)calloc(threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++) {
        size_t lo = n * t / threads, hi = n * (t + 1) / threads;
        jobs[t].points = points_xyz + 18 * lo;
        jobs[t].scalars = scalars + 4 * lo;
        jobs[t].n = hi - lo;
        jobs[t].b = b;
        jobs[t].c = c;
        pthread_create(&th[t], NULL, msm_worker, &jobs[t]);
    }
    g1p acc = g1_identity();
    int rc = 0;
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
        acc = g1_add(&acc, &jobs[t].out);
    }
    memcpy(out_xyz, &acc, sizeof(g1p));
    free(jobs);
    free(th);
    return rc;
}

/* ------------------------------------------------------------------ src/utils.rs */
static const fr ROOT_OF_UNITY = {{0xb9b58d8c5f0e466aull, 0x5b1b4c801819d7ecull, 0x0af53ae352a31e64ull, 0x5bf3adda19e9b27bull}};
static const fr ROOT_OF_UNITY_INV = {{0x4256481adcf3219aull, 0x45f37b7f96b6cad3ull, 0xf9c3f1d75f7a3b27ull, 0x2d2fc049658afd43ull}};

/* ntt_381 / i_ntt_381 (utils.rs:63-81, 106-129): n x n u64 matrix of x*y, one 256-bit pow per term.
 * rows [row_lo, row_hi) only, so a bounded sample of a large transform can be timed. */
static int dft_rows(const u64* in, u64* out, size_t n, int inverse, size_t row_lo, size_t row_hi) {
    if (n == 0 || (n & (n - 1))) return -1; /* assert!(is_power_of_two(n)) */
    const fr* e = (const fr*)in;
    fr* o = (fr*)out;
    const fr root = inverse ? ROOT_OF_UNITY_INV : ROOT_OF_UNITY;
    u64 step = ((u64)1 << 32) / n;
    fr ninv = fr_one();
    for (size_t x = row_lo; x < row_hi; x++) {
        u64* row = (u64*)malloc(sizeof(u64) * n); /* one row of the reference's n x n matrix */
        for (size_t y = 0; y < n; y++) row[y] = (u64)x * y;
        fr sum = {{0, 0, 0, 0}};
        for (size_t y = 0; y < n; y++) {
            u64 by[4] = {row[y] * step, 0, 0, 0};
            sum = fr_add(sum, fr_mul(e[y], fr_pow(root, by)));
        }
        if (inverse) {
            ninv = fr_invert(fr_from_u64(n)); /* per output, as the reference (utils.rs:126) */
            sum = fr_mul(sum, ninv);
        }
        o[x] = sum;
        free(row);
    }
    return 0;
}
int oracle_ntt_381(const u64* in, u64* out, size_t n) { return dft_rows(in, out, n, 0, 0, n); }
int oracle_i_ntt_381(const u64* in, u64* out, size_t n) { return dft_rows(in, out, n, 1, 0, n); }
int oracle_ntt_381_rows(const u64* in, u64* out, size_t n, int inverse, size_t row_lo, size_t row_hi) {
    return dft_rows(in, out, n, inverse, row_lo, row_hi);
}

/* ------------------------------------------------------------------ src/polynomial.rs */
/* roots_of_unity (utils.rs:45-52): [1, w, w^2, ...] by sequential multiplication, w = ROOT^(2^32 / n) by pow */
static void roots_of_unity_seq(size_t n, fr* out) {
    u64 by[4] = {((u64)1 << 32) / n, 0, 0, 0};
    fr w = fr_pow(ROOT_OF_UNITY, by), cur = fr_one();
    for (size_t i = 0; i < n; i++) {
        out[i] = cur;
        cur = fr_mul(cur, w);
    }
}
/* coeffs_evaluate (polynomial.rs:34-45): sum_i c_i * x.pow([i,0,0,0]) -- one 256-bit pow per term */
static fr coeffs_evaluate(const fr* c, size_t len, fr x) {
    fr res = {{0, 0, 0, 0}};
    for (size_t i = 0; i < len; i++) {
        u64 by[4] = {(u64)i, 0, 0, 0};
        res = fr_add(res, fr_mul(c[i], fr_pow(x, by)));
    }
    return res;
}
/* impl Mul for Polynomial, Monomial branch (polynomial.rs:241-273): evaluate both operands on the D-th roots of
 * unity (D = find_next_power_of_two, utils.rs:54-61), multiply pointwise, i_ntt_381, keep la + lb - 1 coefficients */
int oracle_poly_mul(const u64* a, size_t la, const u64* b, size_t lb, u64* out) {
    if (la == 0 || lb == 0) return -1;
    size_t target = la + lb - 1, D = 1;
    while (D < target) D <<= 1;
    fr* roots = (fr*)malloc(sizeof(fr) * D);
    fr* prod = (fr*)malloc(sizeof(fr) * D);
    fr* coef = (fr*)malloc(sizeof(fr) * D);
    roots_of_unity_seq(D, roots);
    for (size_t i = 0; i < D; i++)
        prod[i] = fr_mul(coeffs_evaluate((const fr*)a, la, roots[i]), coeffs_evaluate((const fr*)b, lb, roots[i]));
    int rc = dft_rows((const u64*)prod, (u64*)coef, D, 1, 0, D);
    memcpy(out, coef, sizeof(fr) * target);
    free(roots);
    free(prod);
    free(coef);
    return rc;
}

/* ------------------------------------------------------------------ helpers for the tests */
void oracle_fp_mul(const u64* a, const u64* b, u64* r) { f_mul(&FP, r, a, b); }
void oracle_fr_mul(const u64* a, const u64* b, u64* r) { f_mul(&FR, r, a, b); }
void oracle_fp_add(const u64* a, const u64* b, u64* r) { f_add(&FP, r, a, b); }
void oracle_fp_sub(const u64* a, const u64* b, u64* r) { f_sub(&FP, r, a, b); }
void oracle_g1_add(const u64* a, const u64* b, u64* r) { g1p t = g1_add((const g1p*)a, (const g1p*)b); memcpy(r, &t, sizeof t); }
void oracle_g1_double(const u64* a, u64* r) { g1p t = g1_double((const g1p*)a); memcpy(r, &t, sizeof t); }
/* sum_i s_i tau^i (Horner) on Montgomery scalars -> Montgomery result: closed-form MSM exponent */
void oracle_fr_horner(const u64* scalars, size_t n, const u64* tau, u64* out) {
    const fr* s = (const fr*)scalars;
    fr t, acc = {{0, 0, 0, 0}};
    memcpy(&t, tau, 32);
    for (size_t i = n; i-- > 0;) acc = fr_add(fr_mul(acc, t), s[i]);
    memcpy(out, &acc, 32);
}
/* [1]G, [2]G, ... [n]G as (non-normalised) G1Projective, by repeated complete addition of the
 * generator (g1.rs:199-214): cheap valid points for the CPU timing legs */
void oracle_g1_iota(size_t n, u64* out_xyz) {
    g1p g;
    static const u64 gx[6] = {0x5cb38790fd530c16ull, 0x7817fc679976fff5ull, 0x154f95c7143ba1c1ull,
                              0xf0ae6acdf3d0e747ull, 0xedce6ecc21dbf440ull, 0x120177419e0bfb75ull};
    static const u64 gy[6] = {0xbaac93d50ce72271ull, 0x8c22631a7918fd8eull, 0xdd595f13570725ceull,
                              0x51ac582950405194ull, 0x0e1c8c3fad0059c0ull, 0x0bbc3efc5008a26aull};
    memcpy(g.x.l, gx, 48);
    memcpy(g.y.l, gy, 48);
    g.z = fp_one();
    g1p cur = g;
    for (size_t i = 0; i < n; i++) {
        memcpy(out_xyz + 18 * i, &cur, sizeof cur);
        cur = g1_add(&cur, &g);
    }
}
/* P, [k]P, [k^2]P, ... (n points) for a small multiplier k by double-and-add with the complete formulas:
 * the synthetic SRS [tau^i]G of the benchmark workload (Setup::generate_srs, setup.rs:12-31, multiplies by tau
 * the same way) for the CPU timing legs, continued from any starting point so that threads can split the range */
void oracle_g1_powers_small(const u64* start_xyz, u64 k, size_t n, u64* out_xyz) {
    g1p cur;
    memcpy(&cur, start_xyz, sizeof cur);
    for (size_t i = 0; i < n; i++) {
        memcpy(out_xyz + 18 * i, &cur, sizeof cur);
        g1p acc = g1_identity();
        for (int b = 63; b >= 0; b--) {
            acc = g1_double(&acc);
            if ((k >> b) & 1) acc = g1_add(&acc, &cur);
        }
        cur = acc;
    }
}
