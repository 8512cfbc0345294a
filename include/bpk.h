/* bpk.h -- C ABI of the B200-native MSM / NTT library for baby-plonk-rust.
 *
 * The reference (ChainUpZero/baby-plonk-rust) has no FFI; its hot path is reached through three
 * plain Rust call surfaces.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference root).  A Rust maintainer binds these
 * with `extern "C"` exactly as INTEGRATION.md shows.
 *
 * Data conventions (identical to what Rust holds in memory, so `as_ptr()` is passed unchanged):
 *   Scalar (Fr)      4 x u64 little-endian limbs, Montgomery form, R = 2^256, value < q
 *                    (lib/bls12_381/src/scalar.rs:22)
 *   Fp               6 x u64 limbs, Montgomery, R = 2^384, value < p  (lib/bls12_381/src/fp.rs:15)
 *   G1Projective     18 x u64: X | Y | Z homogeneous projective, Z == 0 <=> identity
 *                    (lib/bls12_381/src/g1.rs:442-446, identity (0,1,0) g1.rs:605-611)
 * Every G1 output is the affine-normalised representative (x, y, 1) or the identity (0, 1, 0), all
 * in Montgomery form; G1Projective equality in the reference is projective equivalence
 * (g1.rs:479-496), so Rust-side `==`, the transcript bytes and the verifier see no difference.
 *
 * Ownership: the caller owns every buffer; the library never keeps a host pointer after return.
 * Threading: every entry point takes the context's own lock for its whole duration, so calls on one context from
 * several threads (e.g. `cargo test`, which runs the reference's tests in parallel) are serialised by the library;
 * distinct contexts may be used concurrently.  bpk_destroy must not race with other calls on the same context.
 * Field elements handed over as scalars (tau, coset shifts, evaluation points) must be canonical (< modulus),
 * otherwise BPK_ERR_INVALID_ARG; coordinates of points are not validated (garbage in, garbage out, but no call hangs).
 * Errors: every function returns 0 on success or a negative bpk_status; nothing throws or aborts.
 * The Rust shim turns a non-zero status into `panic!`, preserving the reference's error behaviour.
 * There is no CPU fallback: without a CUDA device bpk_init fails with BPK_ERR_NO_DEVICE.
 */
#ifndef BPK_H
#define BPK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bpk_ctx bpk_ctx;

enum bpk_status {
    BPK_OK = 0,
    BPK_ERR_NO_DEVICE = -1,    /* no usable CUDA device / wrong architecture */
    BPK_ERR_CUDA = -2,         /* a CUDA runtime call failed; see bpk_last_error */
    BPK_ERR_INVALID_ARG = -3,  /* null pointer, bad handle, ... */
    BPK_ERR_NOT_POW2 = -4,     /* NTT length is not a power of two (utils.rs:65,108 assert!) */
    BPK_ERR_TOO_LARGE = -5,    /* size beyond the supported range */
    BPK_ERR_WINDOW = -6,       /* bucket_msm (b, c) the reference itself panics on */
    BPK_ERR_OOM = -7
};

/* ---- context ------------------------------------------------------------------------------- */
/* One context per GPU (one process per GPU in multi-GPU runs).  `device` is the CUDA ordinal. */
int bpk_init(bpk_ctx** out, int device);
void bpk_destroy(bpk_ctx* ctx);
const char* bpk_strerror(int status);
const char* bpk_last_error(bpk_ctx* ctx); /* text of the last CUDA error seen by this context */
int bpk_abi_version(void);
/* Use an existing CUDA stream (cudaStream_t) for all work; NULL = legacy default stream. */
int bpk_set_stream(bpk_ctx* ctx, void* cuda_stream);
int bpk_synchronize(bpk_ctx* ctx);

/* ---- SRS (Setup.powers_of_x, src/setup.rs:7-10) --------------------------------------------- */
/* Upload n projective points (18 u64 each) once per Setup; they are kept on the device in affine
 * form.  Replaces passing `&self.powers_of_x` on every commit (setup.rs:36). */
int bpk_srs_load(bpk_ctx* ctx, const uint64_t* points_xyz, size_t n, uint64_t* handle_out);
/* Setup::generate_srs powers_of_x (setup.rs:12-31): [tau^i] G for i < n, built on the device. */
int bpk_srs_generate(bpk_ctx* ctx, const uint64_t tau_mont[4], size_t n, uint64_t* handle_out);
/* The slice [tau^i] G for first <= i < first + n: the shard of the SRS one rank of a multi-GPU job owns. */
int bpk_srs_generate_range(bpk_ctx* ctx, const uint64_t tau_mont[4], size_t first, size_t n,
                           uint64_t* handle_out);
/* Optional: trade HBM for work.  Stores [2^(c w)] P_i for every c-bit window w next to the SRS (W = ceil(256/c)
 * times the SRS size), so that every later MSM on this handle uses one shared bucket set, fewer windows and no
 * final doubling chain.  window_bits == 0 picks c from the SRS length.  Results are unchanged (bit-exact). */
int bpk_srs_precompute(bpk_ctx* ctx, uint64_t handle, unsigned window_bits);
/* Read points [first, first+count) back as normalised G1Projective limbs. */
int bpk_srs_read(bpk_ctx* ctx, uint64_t handle, size_t first, size_t count, uint64_t* out_xyz);
int bpk_srs_len(bpk_ctx* ctx, uint64_t handle, size_t* n_out);
/* HBM bytes the handle's point table occupies (the SRS, or W x that after bpk_srs_precompute). */
int bpk_srs_table_bytes(bpk_ctx* ctx, uint64_t handle, size_t* bytes_out);
int bpk_srs_free(bpk_ctx* ctx, uint64_t handle);

/* ---- MSM ------------------------------------------------------------------------------------ */
/* BucketMSM::bucket_msm(points, scalars, b, c) (src/msm.rs:76-118) with points = the SRS behind
 * `handle`.  Semantics of the reference, exactly: k = b / c windows; the result is
 * sum_i (s_i >> (256 - k*c)) * P_i over the first min(n_points, n_scalars) pairs (zip truncation,
 * msm.rs:29); k == 0 or k*c > 256 is a panic in the reference (msm.rs:105,133) and
 * BPK_ERR_WINDOW here.  For every c dividing 256 -- the only in-tree call is (256, 4),
 * setup.rs:36 -- this is the plain MSM.  The GPU chooses its own window size. */
int bpk_bucket_msm(bpk_ctx* ctx, uint64_t handle, const uint64_t* scalars_mont, size_t n_scalars,
                   size_t b, size_t c, uint64_t out_xyz[18]);
/* Setup::commit (src/setup.rs:32-37) == bpk_bucket_msm(.., 256, 4). */
int bpk_msm_g1(bpk_ctx* ctx, uint64_t handle, const uint64_t* scalars_mont, size_t n_scalars,
               uint64_t out_xyz[18]);
/* Same with the points passed on every call, as the reference signature does (no device cache). */
int bpk_msm_g1_points(bpk_ctx* ctx, const uint64_t* points_xyz, size_t n_points,
                      const uint64_t* scalars_mont, size_t n_scalars, uint64_t out_xyz[18]);
/* Device-resident variant: scalars already in HBM (device pointer, n x 4 u64 Montgomery), result
 * written to a device buffer of 18 u64.  `first` selects the SRS slice [first, first + n) so that a
 * rank of a multi-GPU job can own a shard of the points.  normalise == 0 returns an un-normalised
 * projective representative (a valid G1Projective) -- the per-rank partial sum that is gathered
 * over NCCL and added by bpk_g1_sum. */
int bpk_msm_g1_dev(bpk_ctx* ctx, uint64_t handle, size_t first, const void* d_scalars_mont,
                   size_t n, int normalise, void* d_out_xyz);
/* Scalars in HOST memory, result on the device (normalise as in bpk_msm_g1_dev).  For n >= 2^22 the upload is
 * pipelined with the computation: a head slice of n / 8 scalars is uploaded and accumulated while the copy engine
 * brings the rest over on a second stream; with pinned host memory only the head's upload stays exposed.
 * bpk_bucket_msm / bpk_msm_g1 (the Setup::commit path, src/setup.rs:32-37) use the same pipeline. */
int bpk_msm_g1_from_host(bpk_ctx* ctx, uint64_t handle, size_t first, const uint64_t* scalars_mont, size_t n,
                         int normalise, void* d_out_xyz);
/* `count` independent commitments on the same SRS in one call (the three wire commitments of round 1, the three
 * quotient pieces of round 3, the two opening proofs of round 5: src/prover.rs:253-262, 487-497, 640-646).
 * d_scalars_mont, first and n are host arrays of `count` device pointers / slice starts / lengths; d_out_xyz
 * receives count x 18 u64.  The MSMs run on separate streams, so the latency-bound stages of one overlap the
 * accumulation of the others; results are those of `count` bpk_msm_g1_dev calls. */
int bpk_msm_g1_dev_batch(bpk_ctx* ctx, uint64_t handle, size_t count, const void* const* d_scalars_mont,
                         const size_t* first, const size_t* n, int normalise, void* d_out_xyz);
/* Sum of n projective points (host pointers, 18 u64 each), normalised: the post-gather step. */
int bpk_g1_sum(bpk_ctx* ctx, const uint64_t* points_xyz, size_t n, uint64_t out_xyz[18]);
/* Same with device pointers (the gathered partials stay in HBM). */
int bpk_g1_sum_dev(bpk_ctx* ctx, const void* d_points_xyz, size_t n, void* d_out_xyz);

/* ---- NTT ------------------------------------------------------------------------------------ */
/* ntt_381 (src/utils.rs:63-81): out[i] = sum_j in[j] w^(ij), w = ROOT_OF_UNITY^(2^32 / n), natural
 * order in and out; `batch` independent transforms stored back to back.  n must be a power of two
 * (else BPK_ERR_NOT_POW2, the reference asserts).  in == out is allowed. */
int bpk_ntt_fr(bpk_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n, size_t batch);
/* i_ntt_381 (src/utils.rs:106-129): inverse transform, scaled by n^-1. */
int bpk_intt_fr(bpk_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n, size_t batch);
/* Coset variants (no reference counterpart; used by the quotient pipeline that replaces the
 * evaluate-on-domain loop of impl Mul, src/polynomial.rs:241-273): forward evaluates the
 * polynomial on shift * w^i; inverse interpolates from those evaluations. */
int bpk_coset_ntt_fr(bpk_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n, size_t batch,
                     const uint64_t shift_mont[4]);
int bpk_coset_intt_fr(bpk_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n, size_t batch,
                      const uint64_t shift_mont[4]);
/* Device-resident NTT: d_in / d_out are device pointers (may be equal).
 * flags: bit 0 = inverse, bit 1 = coset (d_shift_mont: HOST pointer to 4 u64, may be NULL). */
int bpk_ntt_fr_dev(bpk_ctx* ctx, const void* d_in, void* d_out, size_t n, size_t batch, int flags,
                   const uint64_t* shift_mont);
/* impl Mul for Polynomial, Monomial x Monomial (src/polynomial.rs:189-273): out has la + lb - 1
 * coefficients (trailing zeros kept, polynomial.rs:272). la, lb >= 1. */
int bpk_poly_mul_fr(bpk_ctx* ctx, const uint64_t* a, size_t la, const uint64_t* b, size_t lb,
                    uint64_t* out);

/* Same with device pointers (la, lb >= 1; d_out holds la + lb - 1 coefficients; may not alias the inputs). */
int bpk_poly_mul_fr_dev(bpk_ctx* ctx, const void* d_a, size_t la, const void* d_b, size_t lb, void* d_out);

/* ---- device memory (for hosts without a CUDA binding of their own: the Rust / C++ callers of the d_* entry points) ----
 * Transfers and fills are ordered on the context's stream; bpk_dev_download returns when the bytes have arrived;
 * bpk_dev_upload from pageable memory returns when the source may be reused.  Freed blocks are cached by size and
 * reused (stream-ordered, no device synchronisation); everything is released by bpk_destroy. */
int bpk_dev_alloc(bpk_ctx* ctx, size_t bytes, void** d_out);
int bpk_dev_free(bpk_ctx* ctx, void* d_ptr);
int bpk_dev_upload(bpk_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
int bpk_dev_download(bpk_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
int bpk_dev_copy(bpk_ctx* ctx, void* d_dst, const void* d_src, size_t bytes);
int bpk_dev_zero(bpk_ctx* ctx, void* d_dst, size_t bytes);

/* ---- device-resident Fr vector / polynomial primitives (the callers' O(n) work) ----
 * All pointers named d_* are device pointers to n x 4 u64 Montgomery limbs; scalars are host pointers to
 * 4 u64.  These keep the prover's polynomials in HBM between the NTTs and the commitments. */
/* op: 0 a+b, 1 a-b, 2 a*b (pointwise), 3 a*s, 4 a + s*b, 5 a+s on every element, 6 a - s*b.
 * Polynomial Add / Sub / Mul<Scalar> of src/polynomial.rs:57-187 on equal-length operands. d_out may alias. */
int bpk_fr_vec_op(bpk_ctx* ctx, int op, const void* d_a, const void* d_b, const uint64_t* scalar_mont,
                  void* d_out, size_t n);
/* out[i] = a[i] * c0 * g^i: z(X) -> z(wX) (src/prover.rs:661-674), coset shifts. */
int bpk_fr_scale_powers(bpk_ctx* ctx, const void* d_a, const uint64_t g_mont[4], const uint64_t c0_mont[4],
                        void* d_out, size_t n);
/* Polynomial::coeffs_evaluate (src/polynomial.rs:34-45): sum_i c_i x^i. */
int bpk_fr_poly_eval(bpk_ctx* ctx, const void* d_coeffs, size_t n, const uint64_t x_mont[4], uint64_t out_mont[4]);
/* The same for `count` (<= 64) polynomials at one point (round 4 evaluates a, b, c, s1, s2 at zeta, src/prover.rs:502-541):
 * d_coeffs and n are host arrays of device pointers / lengths, out_mont receives count x 4 u64. */
int bpk_fr_poly_eval_many(bpk_ctx* ctx, size_t count, const void* const* d_coeffs, const size_t* n,
                          const uint64_t x_mont[4], uint64_t* out_mont);
/* impl Div by the linear polynomial X - root (src/polynomial.rs:314-380 as used at src/prover.rs:623-638):
 * n coefficients in, n - 1 quotient coefficients out, remainder dropped. */
int bpk_fr_poly_div_linear(bpk_ctx* ctx, const void* d_coeffs, size_t n, const uint64_t root_mont[4],
                           void* d_quotient);
/* impl Div by Z_H = X^n - 1 (src/prover.rs:450): len coefficients in, len - n quotient coefficients out. */
int bpk_fr_poly_div_vanishing(bpk_ctx* ctx, const void* d_coeffs, size_t len, size_t n, void* d_quotient);
/* Round 2 grand product (src/prover.rs:286-317): d_z receives n + 1 values Z_0 = 1 .. Z_n (Z_n == 1 for a
 * satisfied permutation; the reference asserts it).  Inputs are the wire and sigma columns on H. */
int bpk_plonk_grand_product(bpk_ctx* ctx, const void* d_a, const void* d_b, const void* d_c, const void* d_s1,
                            const void* d_s2, const void* d_s3, size_t n, const uint64_t beta[4],
                            const uint64_t gamma[4], const uint64_t k1[4], const uint64_t k2[4], void* d_z);

/* Round 3 quotient (src/prover.rs:370-452) in evaluation form.  Inputs are evaluations on the coset
 * g * <w_domain> (produce them with bpk_ntt_fr_dev and a shift), rows `domain` values apart:
 *   d_witness_evals  5 rows, per proof:    a, b, c, z, PI
 *   d_circuit_evals 10 rows, per circuit:  ql, qr, qm, qo, qc, s1, s2, s3, L1, X   (may stay resident)
 * d_out receives the evaluations of
 *   t = [a ql + b qr + a b qm + c qo + PI + qc + alpha * perm + alpha^2 (z - 1) L1] / Z_H
 * on the same coset; zh_inv_mont holds the domain / n (<= 64) values 1 / Z_H(g w_domain^i), i < domain / n. */
int bpk_plonk_quotient_evals(bpk_ctx* ctx, const void* d_witness_evals, const void* d_circuit_evals, size_t domain,
                             size_t n, const uint64_t beta[4], const uint64_t gamma[4], const uint64_t alpha[4],
                             const uint64_t k1[4], const uint64_t k2[4], const uint64_t* zh_inv_mont, void* d_out);

/* Multi-GPU form of round 3 (the independent transforms of the round dealt over the ranks): the quotient
 * domain g <w_4n> is the union of R sub-cosets (g w_4n^j) <w_m>, m = 4n / R; rank j evaluates t on its own one.
 *   bpk_fr_fold             out[r][k] = sum_q in[r][k + q m] s^q: coefficients of p mod (X^m - s), s = (g w_4n^j)^m; a coset
 *                           NTT of size m with shift g w_4n^j (bpk_ntt_fr_dev) then yields p on the sub-coset.  `rows`
 *                           polynomials of `len` coefficients, stored row_stride apart.
 *   bpk_plonk_quotient_evals_shard   as bpk_plonk_quotient_evals on `points` = m points, witness rows a, b, c, z, PI and
 *                           z(wX) (a sixth row: on a sub-coset it is not a rotation of z's), zh_period = max(1, 4 / R)
 *                           values 1 / Z_H(g w_4n^j w_m^i), i < zh_period.
 * The ranks' parts are gathered (NCCL all-gather of m x 32 B per rank), interleaved (point j + R i) and interpolated by
 * one inverse coset NTT of size 4n. */
int bpk_fr_fold(bpk_ctx* ctx, const void* d_in, size_t rows, size_t row_stride, size_t len, size_t m,
                const uint64_t s_mont[4], void* d_out);
int bpk_plonk_quotient_evals_shard(bpk_ctx* ctx, const void* d_witness_evals, const void* d_circuit_evals, size_t points,
                                   unsigned zh_period, const uint64_t beta[4], const uint64_t gamma[4],
                                   const uint64_t alpha[4], const uint64_t k1[4], const uint64_t k2[4],
                                   const uint64_t* zh_inv_mont, void* d_out);

/* ---- host utility ---------------------------------------------------------------------------------
 * keccak-f[1600] on 25 little-endian 64-bit lanes (lane x + 5 y), in place; runs on the host.  The Fiat-Shamir
 * transcript of the reference (src/transcript.rs over merlin / STROBE-128) sits on it; the Python host layer calls
 * it instead of permuting in the interpreter. */
void bpk_keccak_f1600(uint64_t lanes[25]);

/* The synthetic benchmark circuit of SURVEY 8d (C4: row 0 "out public", rows alternate c_k <== c_{k-1} * y_k and
 * c_k <== c_{k-1} + y_k, `gates` rows used of n) as the pre-processed columns src/program.rs:51-147 would produce:
 * columns[0..4] = QL QR QM QO QC, [5..7] = S1 S2 S3, [8..10] = the witness columns A B C (n x 4 u64 Montgomery each,
 * caller-allocated); public_out = the public input, canonical.  Runs on the host. */
int bpk_synthetic_chain_circuit(size_t n, size_t gates, uint64_t seed, uint64_t* const columns[11], uint64_t public_out[4]);

/* ---- instrumentation (bench.py, tests) -------------------------------------------------------- */
/* When enabled, every kernel stage is bracketed by CUDA events on the context's stream. */
int bpk_profile_enable(bpk_ctx* ctx, int on);
int bpk_profile_reset(bpk_ctx* ctx);
/* Accumulated device time (ms) and number of kernel launches of stage `name`
 * ("msm.recode", "msm.sort", "msm.accumulate", "msm.reduce", "msm.finalize", "ntt.pass", ...). */
int bpk_profile_get(bpk_ctx* ctx, const char* name, double* ms_out, uint64_t* launches_out);
/* Total kernel launches issued by this context since bpk_profile_reset. */
uint64_t bpk_launch_count(bpk_ctx* ctx);
/* Plan of the most recent MSM: {window bits c, windows W, pairs per accumulate thread, buckets}. */
int bpk_msm_last_plan(bpk_ctx* ctx, unsigned out[4]);
/* Counters of the most recent MSM (synchronises the stream): {entries = non-zero digits, additions done in affine
 * coordinates by the pairwise tree, additions left to the XYZZ tail, non-empty buckets, tree levels launched,
 * additions per shared inversion}. */
int bpk_msm_last_stats(bpk_ctx* ctx, uint64_t out[6]);
/* Register-only IMAD.WIDE throughput probe: returns 32x32+64 multiply-adds per second.  "imad.mode" selects the form: 0
 * independent accumulates, 1 two 12-limb carry chains, 2 fused accumulates with a 64-bit addend, 3 the Fp product the MSM
 * kernels call (dependent chain), 4 the same product inlined, 5 two independent inlined products; "imad.warps_per_sm"
 * (4..64, multiple of 4, default 64) the resident warps per SM -- the pair that showed the multiplier saturates with two
 * warps per scheduler (profiles/r2_affine_v2.md). */
int bpk_imad_peak(bpk_ctx* ctx, double* wide_imad_per_s_out, double* seconds_out);
/* Tunables (A/B measurements and tests; the defaults are the measured optima):
 *   "msm.window" window bits of the non-precomputed MSM (0 = auto), "msm.chunk" pairs per accumulate thread (0 = auto),
 *   "msm.affine_levels" levels of the batched-affine pairwise tree (-1 = from the expected bucket load, 0 = XYZZ chunks only),
 *   "msm.min_pairs" a tree level expected to hold fewer pairs is left to the XYZZ tail, "msm.batch" additions per shared
 *   inversion, "msm.level_mib" memory budget of the tree's level buffers, "msm.cta_shape" level kernel as four CTAs of 4 warps
 *   per SM (1), one CTA of 16 warps (2) or by the level's size (0, default), "msm.tree_top" 0 = one launch per level of the
 *   bucket-reduction tree, "msm.scatter_l2_mib" the sorted list is scattered in phases over bucket ranges of at most this size,
 *   "msm.lanes" 1..3 concurrent MSMs of bpk_msm_g1_dev_batch, "msm.host_slices" 0 = no upload / compute overlap,
 *   "ntt.tile_log2" log2 of the R x C tile per CTA (default 10), "ntt.max_radix_log2" (0 = auto), "ntt.threads",
 *   "ntt.kernel" 0 = auto, 1 = one radix-2 stage per barrier, 2 = register-blocked radix-8 steps, 3 = 4 rows per
 *   thread with two products per call (auto uses it below 2^18 elements),
 *   "ntt.scratch_mib" scratch bound of a batched transform (larger batches go a few rows at a time),
 *   "ntt.direct_max_log2" largest per-size inter-pass twiddle table, "ntt.direct_budget_mib" HBM budget of all
 *   such tables together (a cache: dropped and rebuilt on demand beyond it, or when an allocation fails), "imad.mode" / "imad.warps_per_sm" form and occupancy of bpk_imad_peak,
 *   "host.stage_threads" threads that stage PAGEABLE host inputs through pinned buffers (0 = leave it to the driver).
 * Unknown keys and out-of-range values return BPK_ERR_INVALID_ARG. */
int bpk_set_option(bpk_ctx* ctx, const char* key, long value);

#ifdef __cplusplus
}
#endif
#endif /* BPK_H */
