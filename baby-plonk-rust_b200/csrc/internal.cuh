// Internal declarations shared by the translation units of libbpk.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/bpk.h"
#include "ec.cuh"

namespace bpk {

constexpr int NTT_MAX_LOG = 28;   // largest transform: 2^28 elements (8 GiB)
constexpr int TW_LO_BITS = 13;    // two-level twiddle table split
constexpr int TW_HI_BITS = NTT_MAX_LOG - TW_LO_BITS;
constexpr int MSM_LANES = 3;      // concurrent MSMs of bpk_msm_g1_dev_batch (one stream + one workspace bank each)
constexpr int WS_SLOTS = 24;      // workspace slots per bank

struct DeviceBuffer {
    void* ptr = nullptr;
    size_t bytes = 0;
};

struct SrsEntry {
    // n affine points, Montgomery, (0,0) = infinity.  After bpk_srs_precompute the buffer holds pre_W
    // levels of n points each: level w = [2^(pre_c * w)] P_i (level 0 = the SRS itself), so that all windows
    // of a scalar fall into ONE set of 2^(pre_c - 1) buckets and no window Horner pass is needed.
    affine_t* points = nullptr;
    size_t n = 0;
    uint32_t pre_c = 0;
    uint32_t pre_W = 0;
};

struct MsmPoints {
    const affine_t* base;  // first point of this call's slice (level 0)
    size_t level_stride;   // distance between precomputed levels, in points (0 if not precomputed)
    uint32_t pre_c, pre_W;
};

struct StageStat {
    double ms = 0;
    uint64_t launches = 0;
};

struct PendingEvent {
    std::string name;
    cudaEvent_t start, stop;
    uint64_t launches;
};

}  // namespace bpk

struct bpk_ctx {
    // every extern "C" entry point holds this for its whole duration: calls on one context from several threads
    // (cargo test runs the reference's tests in parallel) are serialised here, not by the caller
    std::recursive_mutex mutex;
    int device = 0;
    bool msm_kernels_configured = false;   // cudaFuncSetAttribute of the level kernels done on this device
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::string last_error;

    // NTT tables: powers of the primitive 2^NTT_MAX_LOG-th root (forward / inverse)
    bpk::fr_t* tw_lo[2] = {nullptr, nullptr};
    bpk::fr_t* tw_hi[2] = {nullptr, nullptr};
    // per-size inter-pass twiddle tables w_{2^lg}^e, e < 2^lg (built on first use, forward / inverse)
    bpk::fr_t* tw_direct[2][bpk::NTT_MAX_LOG + 1] = {};
    size_t tw_direct_bytes = 0;
    // coset tables for the last shift used (lo: g^i, hi: g^(i << TW_LO_BITS)), forward / inverse
    bpk::fr_t* coset_lo[2] = {nullptr, nullptr};
    bpk::fr_t* coset_hi[2] = {nullptr, nullptr};
    bpk::fr_t coset_shift[2];
    bool coset_valid[2] = {false, false};
    size_t coset_n = 0;

    // workspaces (grown on demand, never shrunk): bank 0 for serial calls, banks 1.. for the MSM lanes
    bpk::DeviceBuffer ws[bpk::WS_SLOTS * (bpk::MSM_LANES + 1)];
    int ws_bank = 0;
    cudaStream_t lane_stream[bpk::MSM_LANES] = {};
    cudaEvent_t lane_done[bpk::MSM_LANES] = {};
    cudaEvent_t lane_fork = nullptr;
    cudaStream_t copy_stream = nullptr;  // uploads the tail of a host-resident MSM input under the head's accumulation
    cudaEvent_t copy_done = nullptr;
    // ring of pinned buffers through which PAGEABLE host memory (a Rust Vec) is uploaded at copy-engine speed
    static constexpr int STAGE_RING = 3;
    static constexpr size_t STAGE_CHUNK = (size_t)32 << 20;
    void* stage_buf[STAGE_RING] = {};
    cudaEvent_t stage_done[STAGE_RING] = {};
    long opt_host_stage_threads = 6;   // 0: hand pageable memory to cudaMemcpyAsync (the driver stages it, ~12 GB/s)

    // bpk_dev_alloc / bpk_dev_free: freed blocks are kept by size and handed out again (a prover allocates the same
    // sizes for every proof; cudaMalloc / cudaFree would synchronise the device each time)
    std::map<size_t, std::vector<void*>> dev_pool;
    std::map<void*, size_t> dev_live;

    std::map<uint64_t, bpk::SrsEntry> srs;
    uint64_t next_handle = 1;

    // fixed-base table for bpk_srs_generate (built lazily)
    bpk::affine_t* gen_table = nullptr;

    // options
    long opt_msm_window = 0;
    long opt_msm_chunk = 0;
    long opt_msm_lanes = bpk::MSM_LANES;  // 1: bpk_msm_g1_dev_batch runs its MSMs one after the other
    long opt_msm_host_slices = 1;  // 0: upload all scalars before the MSM starts
    long opt_msm_affine_levels = -1;  // levels of the batched-affine pairwise tree (-1: from the expected bucket load, 0: XYZZ only)
    long opt_msm_cta_shape = 0;        // level kernel: 0 = by level size, 1 = four CTAs of 4 warps per SM, 2 = one CTA of 16 (A/B)
    long opt_msm_min_pairs = 1 << 18;  // a tree level expected to hold fewer pairs is left to the XYZZ tail (r2_msm_plan_sweep.md)
    long opt_msm_batch = 256;          // additions that share one inversion (per thread)
    long opt_msm_level_mib = 48 << 10; // budget of the tree's level buffers
    long opt_msm_tree_top = 1;         // narrow top of the bucket-reduction tree in one block
    long opt_msm_scatter_l2_mib = 400; // the level-0 list is scattered in phases over bucket ranges of at most this size
    long opt_ntt_tile_log2 = 10;  // R x C elements per CTA tile (32 KiB): best of the sweep in profiles/
    long opt_ntt_max_radix_log2 = 0;  // 0 = auto
    long opt_ntt_threads = 0;
    long opt_ntt_kernel = 0;  // 0: auto, 1: one radix-2 stage per barrier, 2: register-blocked radix-8 steps
    long opt_ntt_direct_max_log2 = 25;  // largest direct twiddle table (2^25 x 32 B = 1 GiB)
    long opt_ntt_scratch_mib = 4096;    // scratch of one batched transform; larger batches are transformed a few rows at a time
    long opt_ntt_direct_budget_mib = 3072;  // all direct tables together; beyond it they are dropped and rebuilt on demand
    long opt_imad_mode = 0;
    long opt_imad_warps_per_sm = 64;

    // plan of the most recent MSM (window bits, windows, pairs per accumulate thread, buckets)
    unsigned last_c = 0, last_W = 0, last_chunk = 0, last_buckets = 0, last_levels = 0, last_batch = 0;
    unsigned long long* last_stats_dev = nullptr;  // counters of the most recent bucket fill (device)

    // instrumentation
    bool profiling = false;
    uint64_t launches = 0;
    std::map<std::string, bpk::StageStat> stats;
    std::vector<bpk::PendingEvent> pending;
};

namespace bpk {

int cuda_fail(bpk_ctx* ctx, cudaError_t e, const char* what, const char* file, int line);

#define BPK_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t _e = (call);                                                \
        if (_e != cudaSuccess) return bpk::cuda_fail(ctx, _e, #call, __FILE__, __LINE__); \
    } while (0)

#define BPK_TRY(call)            \
    do {                         \
        int _s = (call);         \
        if (_s != BPK_OK) return _s; \
    } while (0)

// grow-only device workspace slot
int ws_reserve(bpk_ctx* ctx, int slot, size_t bytes, void** out);

// stage instrumentation: end() hands the event pair to the context; a timer that goes out of scope without end()
// (an error return between the two) destroys its events
struct StageTimer {
    bpk_ctx* ctx;
    const char* name;
    uint64_t launches_before;
    cudaEvent_t start = nullptr, stop = nullptr;
    StageTimer(bpk_ctx* c, const char* n);
    ~StageTimer();
    StageTimer(const StageTimer&) = delete;
    StageTimer& operator=(const StageTimer&) = delete;
    void end();
};
int profile_collect(bpk_ctx* ctx);
// host -> device copy, ordered on `stream`; on return the source may be reused.  Pageable sources are staged through
// the context's pinned ring by several threads (the driver's own staging of pageable memory is single-threaded).
int upload_host(bpk_ctx* ctx, void* d_dst, const void* h_src, size_t bytes, cudaStream_t stream);

inline void count_launch(bpk_ctx* ctx, uint64_t n = 1) { ctx->launches += n; }

// ---- ntt.cu ----
int ntt_init_tables(bpk_ctx* ctx);
int ntt_drop_direct_tables(bpk_ctx* ctx);
int ntt_run(bpk_ctx* ctx, const fr_t* d_in, fr_t* d_out, size_t n, size_t batch, bool inverse,
            const fr_t* shift /* host, Montgomery, or null */);
int pointwise_mul(bpk_ctx* ctx, fr_t* d_a, const fr_t* d_b, size_t n);
// out[i] = pre * base^(i << shift), i < count
int launch_pow_table(bpk_ctx* ctx, fr_t* d_out, const fr_t& base, const fr_t& pre, uint32_t count, uint32_t shift);

// ---- polyops.cu ----
int fr_vec_op(bpk_ctx* ctx, int op, const fr_t* a, const fr_t* b, const fr_t& s, fr_t* out, size_t n);
int fr_scale_powers(bpk_ctx* ctx, const fr_t* a, const fr_t& g, const fr_t& c0, fr_t* out, size_t n);
int fr_poly_eval(bpk_ctx* ctx, const fr_t* c, size_t n, const fr_t& x, fr_t* d_out);
int fr_poly_eval_many(bpk_ctx* ctx, size_t count, const fr_t* const* c, const size_t* n, const fr_t& x, fr_t* d_out);
int fr_poly_div_linear(bpk_ctx* ctx, const fr_t* c, size_t n, const fr_t& root, fr_t* q);
int fr_poly_div_vanishing(bpk_ctx* ctx, const fr_t* c, size_t len, size_t n, fr_t* q);
int plonk_grand_product(bpk_ctx* ctx, const fr_t* A, const fr_t* B, const fr_t* C, const fr_t* s1, const fr_t* s2,
                        const fr_t* s3, size_t n, const fr_t& beta, const fr_t& gamma, const fr_t& k1, const fr_t& k2,
                        fr_t* Z);
int plonk_quotient_evals(bpk_ctx* ctx, const fr_t* wv, const fr_t* cv, size_t D, size_t n, const fr_t& beta,
                         const fr_t& gamma, const fr_t& alpha, const fr_t& k1, const fr_t& k2, const fr_t* zh_inv_host, fr_t* out);

int plonk_quotient_evals_shard(bpk_ctx* ctx, const fr_t* wv, const fr_t* cv, size_t m, uint32_t period, const fr_t& beta,
                               const fr_t& gamma, const fr_t& alpha, const fr_t& k1, const fr_t& k2, const fr_t* zh_inv_host,
                               fr_t* out);
int fr_fold(bpk_ctx* ctx, const fr_t* in, size_t rows, size_t stride, size_t len, size_t m, const fr_t& s, fr_t* out);

// ---- msm.cu ----
int msm_run(bpk_ctx* ctx, const MsmPoints& pts, const fr_t* d_scalars, size_t n, unsigned rshift,
            bool normalise, uint64_t* d_out_xyz /* 18 u64 on device */);
int msm_run_from_host(bpk_ctx* ctx, const MsmPoints& pts, const uint64_t* h_scalars, fr_t* d_stage, size_t n,
                      unsigned rshift, bool normalise, uint64_t* d_out_xyz);
int g1_sum_run(bpk_ctx* ctx, const uint64_t* d_points_xyz, size_t n, uint64_t* d_out_xyz);
int msm_read_stats(bpk_ctx* ctx, uint64_t out[4]);
// one wide level of the bucket-reduction tree (msm_tree.cu): nodes_out nodes of k + 1 points from 2 nodes_out of k
void msm_launch_plane_tree_level(cudaStream_t stream, const xyzz_t* in, xyzz_t* out, uint32_t k, size_t nodes_out);

// ---- srs.cu ----
int srs_from_projective(bpk_ctx* ctx, const uint64_t* d_xyz, size_t n, affine_t* d_out);
int srs_to_projective(bpk_ctx* ctx, const affine_t* d_pts, size_t n, uint64_t* d_xyz);
int srs_generate(bpk_ctx* ctx, const fr_t& tau, size_t first, size_t n, affine_t* d_out);
// level w of the precomputed table from level w-1: out[i] = [2^c] in[i]
int srs_precompute_level(bpk_ctx* ctx, const affine_t* d_in, affine_t* d_out, size_t n, uint32_t c);

// ---- misc ----
int imad_peak_run(bpk_ctx* ctx, double* rate, double* seconds);

}  // namespace bpk
