// C ABI of libbpk.so (see include/bpk.h): context, staging of host buffers, error mapping.
#include "internal.cuh"

namespace bpk {

int cuda_fail(bpk_ctx* ctx, cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s) at %s:%d", what, cudaGetErrorName(e), cudaGetErrorString(e), file, line);
    if (ctx) ctx->last_error = buf;
    cudaGetLastError();  // clear the sticky-less error state
    return e == cudaErrorMemoryAllocation ? BPK_ERR_OOM : BPK_ERR_CUDA;
}

int ws_reserve(bpk_ctx* ctx, int slot, size_t bytes, void** out) {
    DeviceBuffer& b = ctx->ws[slot + WS_SLOTS * ctx->ws_bank];
    if (bytes == 0) bytes = 16;
    if (b.bytes < bytes) {
        if (b.ptr) {
            BPK_CUDA(cudaStreamSynchronize(ctx->stream));
            BPK_CUDA(cudaFree(b.ptr));
            b.ptr = nullptr;
            b.bytes = 0;
        }
        size_t want = bytes + bytes / 8;  // a little slack so slowly growing sizes do not realloc every call
        BPK_CUDA(cudaMalloc(&b.ptr, want));
        b.bytes = want;
    }
    *out = b.ptr;
    return BPK_OK;
}

StageTimer::StageTimer(bpk_ctx* c, const char* n) : ctx(c), name(n), launches_before(c->launches) {
    if (!ctx->profiling) return;
    if (cudaEventCreate(&start) != cudaSuccess || cudaEventCreate(&stop) != cudaSuccess) {
        start = stop = nullptr;
        return;
    }
    cudaEventRecord(start, ctx->stream);
}
void StageTimer::end() {
    if (!ctx->profiling || !start) return;
    cudaEventRecord(stop, ctx->stream);
    PendingEvent pe;
    pe.name = name;
    pe.start = start;
    pe.stop = stop;
    pe.launches = ctx->launches - launches_before;
    ctx->pending.push_back(pe);
    start = stop = nullptr;  // owned by the context now
}
StageTimer::~StageTimer() {
    if (start) cudaEventDestroy(start);
    if (stop) cudaEventDestroy(stop);
}
int profile_collect(bpk_ctx* ctx) {
    for (auto& pe : ctx->pending) {
        float ms = 0;
        cudaEventSynchronize(pe.stop);
        cudaEventElapsedTime(&ms, pe.start, pe.stop);
        StageStat& s = ctx->stats[pe.name];
        s.ms += ms;
        s.launches += pe.launches;
        cudaEventDestroy(pe.start);
        cudaEventDestroy(pe.stop);
    }
    ctx->pending.clear();
    return BPK_OK;
}

// ---- uploads from pageable host memory ------------------------------------------------------------
int upload_host(bpk_ctx* ctx, void* d_dst, const void* h_src, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return BPK_OK;
    bool pageable = false;
    if (bytes >= ((size_t)8 << 20) && ctx->opt_host_stage_threads > 0) {
        cudaPointerAttributes attr;
        cudaError_t e = cudaPointerGetAttributes(&attr, h_src);
        if (e != cudaSuccess) cudaGetLastError();
        pageable = e == cudaSuccess && attr.type == cudaMemoryTypeUnregistered;
    }
    if (!pageable) {
        BPK_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, stream));
        return BPK_OK;
    }
    for (int i = 0; i < bpk_ctx::STAGE_RING; i++) {
        if (!ctx->stage_buf[i]) {
            BPK_CUDA(cudaHostAlloc(&ctx->stage_buf[i], bpk_ctx::STAGE_CHUNK, cudaHostAllocDefault));
            BPK_CUDA(cudaEventCreateWithFlags(&ctx->stage_done[i], cudaEventDisableTiming));
            BPK_CUDA(cudaEventRecord(ctx->stage_done[i], stream));
        }
    }
    const int nthreads = (int)(ctx->opt_host_stage_threads > 16 ? 16 : ctx->opt_host_stage_threads);
    const char* src = static_cast<const char*>(h_src);
    char* dst = static_cast<char*>(d_dst);
    size_t k = 0;
    for (size_t off = 0; off < bytes; off += bpk_ctx::STAGE_CHUNK, k++) {
        const size_t len = bytes - off < bpk_ctx::STAGE_CHUNK ? bytes - off : bpk_ctx::STAGE_CHUNK;
        const int slot = (int)(k % bpk_ctx::STAGE_RING);
        BPK_CUDA(cudaEventSynchronize(ctx->stage_done[slot]));   // the copy that last used this buffer has drained
        char* stage = static_cast<char*>(ctx->stage_buf[slot]);
        std::thread workers[16];
        const size_t per = ((len + nthreads - 1) / nthreads + 4095) & ~(size_t)4095;
        int started = 0;
        for (int t = 1; t < nthreads; t++) {
            const size_t lo = per * t;
            if (lo >= len) break;
            const size_t cnt = len - lo < per ? len - lo : per;
            workers[started++] = std::thread([=] { memcpy(stage + lo, src + off + lo, cnt); });
        }
        memcpy(stage, src + off, len < per ? len : per);
        for (int t = 0; t < started; t++) workers[t].join();
        BPK_CUDA(cudaMemcpyAsync(dst + off, stage, len, cudaMemcpyHostToDevice, stream));
        BPK_CUDA(cudaEventRecord(ctx->stage_done[slot], stream));
    }
    return BPK_OK;
}

// ---- IMAD throughput probe ---------------------------------------------------------------------
// (hi:lo) += a * b: the mad.lo.cc / madc.hi pair that ptxas fuses into one IMAD.WIDE.U32 Rd, Ra, Rb, Rd
__device__ __forceinline__ void probe_mad_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}

// mode 0: 8 independent 32x32+64 multiply-adds per iteration (IMAD.WIDE.U32)
// mode 1: two independent 12-limb carry chains (IMAD.WIDE.U32.X), the shape the Fp multiplier issues
__global__ void __launch_bounds__(256) imad_probe_kernel(uint64_t* out, uint32_t iters, uint32_t seed, int mode) {
    uint32_t a = seed + threadIdx.x * 2654435761u + blockIdx.x;
    uint32_t b = a * 747796405u + 2891336453u;
    if (mode == 0) {
        uint64_t acc[8];
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = a + i;
        for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a + (uint32_t)i), "r"(b));
        }
        uint64_t s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s ^= acc[i];
        out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (mode == 2) {
        // fused accumulate: IMAD.WIDE.U32 Rd, Ra, Rb, Rd (what mul() issues), 14 independent columns
        uint32_t lo[14], hi[14], m[14];
#pragma unroll
        for (int i = 0; i < 14; i++) {
            lo[i] = a + i;
            hi[i] = b + i;
            m[i] = a * (i + 3);
        }
        for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 14; i++) probe_mad_wide(lo[i], hi[i], m[i], b);
        }
        uint64_t s = 0;
#pragma unroll
        for (int i = 0; i < 14; i++) s ^= ((uint64_t)hi[i] << 32) | lo[i];
        out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (mode == 3 || mode == 4) {
        // the real Fp multiplier on a dependent chain: 3 = the out-of-line body the MSM kernels call, 4 = inlined;
        // with imad.warps_per_sm this gives the multiplier's duty cycle at the MSM kernels' occupancy
        fp_t x, y;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            x.l[i] = (a + i) & 0x0fffffffu;
            y.l[i] = (b + i) & 0x0fffffffu;
        }
        for (uint32_t it = 0; it < iters; it++) x = mode == 3 ? mul_lazy(x, y) : mul_cc<FpParams, false>(x, y);
        uint64_t s = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) s ^= ((uint64_t)x.l[i] << 32) | y.l[i];
        out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (mode == 5) {
        // two independent inlined products per iteration (ILP 2 inside one thread)
        fp_t x, y, z;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            x.l[i] = (a + i) & 0x0fffffffu;
            y.l[i] = (b + i) & 0x0fffffffu;
            z.l[i] = (a ^ (b + i)) & 0x0fffffffu;
        }
        for (uint32_t it = 0; it < iters; it += 2) {
            x = mul_cc<FpParams, false>(x, y);
            z = mul_cc<FpParams, false>(z, y);
        }
        uint64_t s = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) s ^= ((uint64_t)x.l[i] << 32) | z.l[i];
        out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        uint32_t x[12], y[12], m[12];
#pragma unroll
        for (int i = 0; i < 12; i++) {
            x[i] = a + i;
            y[i] = b + i;
            m[i] = a * (i + 3);
        }
        for (uint32_t it = 0; it < iters; it++) {
            detail::cmad_n<12>(x, m, b);
            detail::cmad_n<12>(y, m, a);
        }
        uint64_t s = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) s ^= ((uint64_t)x[i] << 32) | y[i];
        out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

int imad_peak_run(bpk_ctx* ctx, double* rate, double* seconds) {
    const int mode = (int)ctx->opt_imad_mode;
    const uint32_t iters = mode >= 3 ? 1u << 10 : 1u << 14;
    const unsigned threads = 128;
    const unsigned blocks = (unsigned)ctx->sm_count * (unsigned)(ctx->opt_imad_warps_per_sm / 4);   // one wave, all resident
    uint64_t* d_out;
    BPK_TRY(ws_reserve(ctx, 7, (size_t)blocks * threads * sizeof(uint64_t), (void**)&d_out));
    cudaEvent_t e0, e1;
    BPK_CUDA(cudaEventCreate(&e0));
    BPK_CUDA(cudaEventCreate(&e1));
    imad_probe_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, iters / 16, 1, mode);  // warm-up
    BPK_CUDA(cudaEventRecord(e0, ctx->stream));
    imad_probe_kernel<<<blocks, threads, 0, ctx->stream>>>(d_out, iters, 2, mode);
    BPK_CUDA(cudaEventRecord(e1, ctx->stream));
    count_launch(ctx, 2);
    BPK_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    BPK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    // mode 1: 2 chains x 6; modes 3, 4: one Fp product = 144 + 144 wide IMADs
    double per_thread = mode == 0 ? 8.0 * iters : mode == 2 ? 14.0 * iters : mode >= 3 ? 288.0 * iters : 12.0 * iters;
    *rate = per_thread * threads * blocks / (ms * 1e-3);
    *seconds = ms * 1e-3;
    return BPK_OK;
}

static fr_t fr_from_host(const uint64_t v[4]) {
    fr_t r;
    for (int i = 0; i < 4; i++) {
        r.l[2 * i] = (uint32_t)v[i];
        r.l[2 * i + 1] = (uint32_t)(v[i] >> 32);
    }
    return r;
}

}  // namespace bpk

using namespace bpk;

// Calls on one context are serialised inside the library (include/bpk.h, "Threading").
#define BPK_LOCK(ctx) std::lock_guard<std::recursive_mutex> bpk_guard_((ctx)->mutex)

// Fp / Fr limbs that arrive over the ABI must be canonical (< modulus): the kernels' reduction steps assume it
template <class F>
static bool limbs_canonical(const uint64_t* v) {
    constexpr int N = F::N;
    for (int i = N - 1; i >= 0; i--) {
        const uint32_t w = (uint32_t)(v[i / 2] >> (32 * (i & 1)));
        const uint32_t m = F::params::mod(i);
        if (w != m) return w < m;
    }
    return false;  // equal to the modulus
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int bpk_abi_version(void) { return 1; }

extern "C" const char* bpk_strerror(int s) {
    switch (s) {
        case BPK_OK: return "ok";
        case BPK_ERR_NO_DEVICE: return "no usable CUDA device (sm_100 required; there is no CPU fallback)";
        case BPK_ERR_CUDA: return "CUDA runtime error (see bpk_last_error)";
        case BPK_ERR_INVALID_ARG: return "invalid argument";
        case BPK_ERR_NOT_POW2: return "NTT length is not a power of two";
        case BPK_ERR_TOO_LARGE: return "size beyond the supported range";
        case BPK_ERR_WINDOW: return "bucket_msm window parameters the reference panics on";
        case BPK_ERR_OOM: return "out of device memory";
        default: return "unknown status";
    }
}

extern "C" int bpk_init(bpk_ctx** out, int device) {
    if (!out) return BPK_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        cudaGetLastError();
        return BPK_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return BPK_ERR_NO_DEVICE;
    if (prop.major != 10) return BPK_ERR_NO_DEVICE;  // kernels are built for sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return BPK_ERR_NO_DEVICE;
    bpk_ctx* ctx = new bpk_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->stream = nullptr;  // legacy default stream: ordered with torch's default stream
    int s = ntt_init_tables(ctx);
    if (s != BPK_OK) {
        fprintf(stderr, "bpk_init: %s\n", ctx->last_error.c_str());
        delete ctx;
        return s;
    }
    *out = ctx;
    return BPK_OK;
}

extern "C" void bpk_destroy(bpk_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    profile_collect(ctx);
    for (auto& kv : ctx->srs) cudaFree(kv.second.points);
    for (auto& b : ctx->ws)
        if (b.ptr) cudaFree(b.ptr);
    for (int d = 0; d < 2; d++) {
        cudaFree(ctx->tw_lo[d]);
        cudaFree(ctx->tw_hi[d]);
        cudaFree(ctx->coset_lo[d]);
        cudaFree(ctx->coset_hi[d]);
        for (auto& t : ctx->tw_direct[d])
            if (t) cudaFree(t);
    }
    if (ctx->gen_table) cudaFree(ctx->gen_table);
    for (auto& kv : ctx->dev_pool)
        for (void* p : kv.second) cudaFree(p);
    for (auto& kv : ctx->dev_live) cudaFree(kv.first);
    for (int l = 0; l < MSM_LANES; l++) {
        if (ctx->lane_stream[l]) cudaStreamDestroy(ctx->lane_stream[l]);
        if (ctx->lane_done[l]) cudaEventDestroy(ctx->lane_done[l]);
    }
    if (ctx->lane_fork) cudaEventDestroy(ctx->lane_fork);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->copy_done) cudaEventDestroy(ctx->copy_done);
    for (int i = 0; i < bpk_ctx::STAGE_RING; i++) {
        if (ctx->stage_buf[i]) cudaFreeHost(ctx->stage_buf[i]);
        if (ctx->stage_done[i]) cudaEventDestroy(ctx->stage_done[i]);
    }
    delete ctx;
}

extern "C" const char* bpk_last_error(bpk_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

extern "C" int bpk_set_stream(bpk_ctx* ctx, void* s) {
    if (!ctx) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    BPK_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stream = (cudaStream_t)s;
    return BPK_OK;
}

extern "C" int bpk_synchronize(bpk_ctx* ctx) {
    if (!ctx) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    BPK_CUDA(cudaStreamSynchronize(ctx->stream));
    return BPK_OK;
}

extern "C" int bpk_set_option(bpk_ctx* ctx, const char* key, long value) {
    if (!ctx || !key) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    std::string k(key);
    if (k == "msm.window") ctx->opt_msm_window = value;
    else if (k == "msm.chunk") ctx->opt_msm_chunk = value;
    else if (k == "msm.affine_levels") {
        if (value < -1 || value > 29) return BPK_ERR_INVALID_ARG;
        ctx->opt_msm_affine_levels = value;
    } else if (k == "msm.min_pairs") {
        if (value < 0) return BPK_ERR_INVALID_ARG;
        ctx->opt_msm_min_pairs = value;
    } else if (k == "msm.batch") {
        if (value < 1 || value > 4096) return BPK_ERR_INVALID_ARG;
        ctx->opt_msm_batch = value;
    } else if (k == "msm.cta_shape") {
        if (value < 0 || value > 2) return BPK_ERR_INVALID_ARG;
        ctx->opt_msm_cta_shape = value;
    } else if (k == "msm.level_mib") {
        if (value < 0) return BPK_ERR_INVALID_ARG;
        ctx->opt_msm_level_mib = value;
    } else if (k == "msm.scatter_l2_mib") {
        if (value < 0) return BPK_ERR_INVALID_ARG;
        ctx->opt_msm_scatter_l2_mib = value;
    } else if (k == "msm.tree_top") ctx->opt_msm_tree_top = value;
    else if (k == "msm.lanes") ctx->opt_msm_lanes = value;
    else if (k == "msm.host_slices") ctx->opt_msm_host_slices = value;
    else if (k == "ntt.tile_log2") {
        if (value < 1 || value > 12) return BPK_ERR_INVALID_ARG;
        ctx->opt_ntt_tile_log2 = value;
    } else if (k == "ntt.max_radix_log2") {
        if (value < 0 || value > 12) return BPK_ERR_INVALID_ARG;
        ctx->opt_ntt_max_radix_log2 = value;
    } else if (k == "ntt.direct_max_log2") {
        if (value < 0 || value > 28) return BPK_ERR_INVALID_ARG;
        ctx->opt_ntt_direct_max_log2 = value;
    } else if (k == "ntt.scratch_mib") {
        if (value < 1) return BPK_ERR_INVALID_ARG;
        ctx->opt_ntt_scratch_mib = value;
    } else if (k == "ntt.direct_budget_mib") {
        if (value < 0) return BPK_ERR_INVALID_ARG;
        ctx->opt_ntt_direct_budget_mib = value;
    } else if (k == "ntt.threads") {
        if (value != 0 && (value < 32 || value > 1024 || (value & 31))) return BPK_ERR_INVALID_ARG;
        ctx->opt_ntt_threads = value;
    } else if (k == "ntt.kernel") ctx->opt_ntt_kernel = value;
    else if (k == "host.stage_threads") {
        if (value < 0 || value > 16) return BPK_ERR_INVALID_ARG;
        ctx->opt_host_stage_threads = value;
    } else if (k == "imad.mode") ctx->opt_imad_mode = value;
    else if (k == "imad.warps_per_sm") {
        if (value < 4 || value > 64 || value % 4) return BPK_ERR_INVALID_ARG;
        ctx->opt_imad_warps_per_sm = value;
    }
    else return BPK_ERR_INVALID_ARG;
    return BPK_OK;
}

// ------------------------------------------------------------------------------------------------
// SRS
// ------------------------------------------------------------------------------------------------
static int srs_register(bpk_ctx* ctx, affine_t* pts, size_t n, uint64_t* handle_out) {
    uint64_t h = ctx->next_handle++;
    SrsEntry e;
    e.points = pts;
    e.n = n;
    ctx->srs[h] = e;
    *handle_out = h;
    return BPK_OK;
}

extern "C" int bpk_srs_load(bpk_ctx* ctx, const uint64_t* points_xyz, size_t n, uint64_t* handle_out) {
    if (!ctx || !handle_out || (n && !points_xyz)) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    BPK_CUDA(cudaSetDevice(ctx->device));
    affine_t* pts = nullptr;
    BPK_CUDA(cudaMalloc(&pts, (n ? n : 1) * sizeof(affine_t)));
    if (n) {
        // stage in slices so the projective copy never needs more than 256 MiB of scratch
        const size_t slice = (size_t)1 << 20;
        uint64_t* d_xyz;
        int s = ws_reserve(ctx, 8, (n < slice ? n : slice) * 18 * sizeof(uint64_t), (void**)&d_xyz);
        if (s != BPK_OK) { cudaFree(pts); return s; }
        for (size_t off = 0; off < n; off += slice) {
            size_t cnt = n - off < slice ? n - off : slice;
            s = upload_host(ctx, d_xyz, points_xyz + 18 * off, cnt * 18 * sizeof(uint64_t), ctx->stream);
            if (s != BPK_OK) { cudaFree(pts); return s; }
            cudaError_t e;
            s = srs_from_projective(ctx, d_xyz, cnt, pts + off);
            if (s != BPK_OK) { cudaFree(pts); return s; }
            e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { cudaFree(pts); return cuda_fail(ctx, e, "srs convert", __FILE__, __LINE__); }
        }
    }
    return srs_register(ctx, pts, n, handle_out);
}

extern "C" int bpk_srs_generate(bpk_ctx* ctx, const uint64_t tau_mont[4], size_t n, uint64_t* handle_out) {
    return bpk_srs_generate_range(ctx, tau_mont, 0, n, handle_out);
}

extern "C" int bpk_srs_generate_range(bpk_ctx* ctx, const uint64_t tau_mont[4], size_t first, size_t n,
                                      uint64_t* handle_out) {
    if (!ctx || !handle_out || !tau_mont) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    if (!limbs_canonical<fr_t>(tau_mont)) return BPK_ERR_INVALID_ARG;
    BPK_CUDA(cudaSetDevice(ctx->device));
    affine_t* pts = nullptr;
    BPK_CUDA(cudaMalloc(&pts, (n ? n : 1) * sizeof(affine_t)));
    int s = srs_generate(ctx, fr_from_host(tau_mont), first, n, pts);
    if (s == BPK_OK) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) s = cuda_fail(ctx, e, "srs generate", __FILE__, __LINE__);
    }
    if (s != BPK_OK) { cudaFree(pts); return s; }
    return srs_register(ctx, pts, n, handle_out);
}

extern "C" int bpk_srs_len(bpk_ctx* ctx, uint64_t handle, size_t* n_out) {
    if (!ctx || !n_out) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    auto it = ctx->srs.find(handle);
    if (it == ctx->srs.end()) return BPK_ERR_INVALID_ARG;
    *n_out = it->second.n;
    return BPK_OK;
}

extern "C" int bpk_srs_read(bpk_ctx* ctx, uint64_t handle, size_t first, size_t count, uint64_t* out_xyz) {
    if (!ctx || (count && !out_xyz)) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    auto it = ctx->srs.find(handle);
    if (it == ctx->srs.end()) return BPK_ERR_INVALID_ARG;
    if (first > it->second.n || count > it->second.n - first) return BPK_ERR_INVALID_ARG;
    if (count == 0) return BPK_OK;
    BPK_CUDA(cudaSetDevice(ctx->device));
    uint64_t* d_xyz;
    BPK_TRY(ws_reserve(ctx, 8, count * 18 * sizeof(uint64_t), (void**)&d_xyz));
    BPK_TRY(srs_to_projective(ctx, it->second.points + first, count, d_xyz));
    BPK_CUDA(cudaMemcpyAsync(out_xyz, d_xyz, count * 18 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    BPK_CUDA(cudaStreamSynchronize(ctx->stream));
    return BPK_OK;
}

extern "C" int bpk_srs_free(bpk_ctx* ctx, uint64_t handle) {
    if (!ctx) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    auto it = ctx->srs.find(handle);
    if (it == ctx->srs.end()) return BPK_ERR_INVALID_ARG;
    BPK_CUDA(cudaSetDevice(ctx->device));
    BPK_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(it->second.points);
    ctx->srs.erase(it);
    return BPK_OK;
}

// Trade HBM for work: store [2^(c w)] P_i for every window w next to the SRS (W x the SRS size; a 2^24-point
// SRS with c = 22 is 12 x 1.5 GiB of a B200's 180 GB).  All windows of a scalar then share one set of
// 2^(c-1) buckets: fewer windows for the same bucket memory, and no Horner doubling chain at the end.
extern "C" int bpk_srs_precompute(bpk_ctx* ctx, uint64_t handle, unsigned window_bits) {
    if (!ctx) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    auto it = ctx->srs.find(handle);
    if (it == ctx->srs.end()) return BPK_ERR_INVALID_ARG;
    SrsEntry& e = it->second;
    if (e.pre_c != 0) return window_bits == 0 || window_bits == e.pre_c ? BPK_OK : BPK_ERR_INVALID_ARG;
    if (e.n == 0) return BPK_OK;
    unsigned c = window_bits;
    if (c == 0) {
        // measured optimum of the sweeps in profiles/r2_msm_plan_sweep.md (c = 8..21 at 2^16..2^22, c = 20..23 at 2^23, 2^24):
        // larger windows trade W n pair additions against 2^(c-1) bucket additions and tree depth.  Windows that leave the
        // top digit only 3 bits (c = 18, 21: 255 = 14 x 18 + 3 = 12 x 21 + 3) put all n entries of that digit into four
        // buckets -- contended histogram updates and long runs in the tail -- and lose to their neighbours.
        unsigned lg = 0;
        while (((size_t)2 << lg) <= e.n) lg++;          // floor(log2 n)
        if (e.n - ((size_t)1 << lg) >= ((size_t)1 << lg) / 2) lg++;  // nearest power of two
        c = lg <= 16 ? (lg < 9 ? 6 : lg - 3) : lg <= 18 ? 16 : lg == 19 ? 17 : lg <= 23 ? 20 : 22;
    }
    if (c < 2 || c > 24) return BPK_ERR_INVALID_ARG;
    const unsigned W = (256 + c - 1) / c;
    if ((size_t)W * e.n >= ((size_t)1 << 31)) return BPK_ERR_TOO_LARGE;
    BPK_CUDA(cudaSetDevice(ctx->device));
    affine_t* table = nullptr;
    BPK_CUDA(cudaMalloc(&table, (size_t)W * e.n * sizeof(affine_t)));
    cudaError_t err = cudaMemcpyAsync(table, e.points, e.n * sizeof(affine_t), cudaMemcpyDeviceToDevice, ctx->stream);
    int s = err == cudaSuccess ? BPK_OK : cuda_fail(ctx, err, "precompute copy", __FILE__, __LINE__);
    for (unsigned w = 1; w < W && s == BPK_OK; w++)
        s = srs_precompute_level(ctx, table + (size_t)(w - 1) * e.n, table + (size_t)w * e.n, e.n, c);
    if (s == BPK_OK) {
        err = cudaStreamSynchronize(ctx->stream);
        if (err != cudaSuccess) s = cuda_fail(ctx, err, "precompute", __FILE__, __LINE__);
    }
    if (s != BPK_OK) { cudaFree(table); return s; }
    cudaFree(e.points);
    e.points = table;
    e.pre_c = c;
    e.pre_W = W;
    return BPK_OK;
}

// ------------------------------------------------------------------------------------------------
// MSM
// ------------------------------------------------------------------------------------------------
static int window_to_shift(size_t b, size_t c, unsigned* rshift) {
    if (c == 0) return BPK_ERR_WINDOW;        // division by zero panics in Rust
    size_t k = b / c;
    if (k == 0) return BPK_ERR_WINDOW;        // t_points[0] out of bounds (msm.rs:105)
    if (k * c > 256) return BPK_ERR_WINDOW;   // bits[start..end] out of range (msm.rs:133)
    if (c > 63) return BPK_ERR_WINDOW;        // 1 << c overflow / bools_to_u64 shift overflow
    *rshift = (unsigned)(256 - k * c);
    return BPK_OK;
}

static MsmPoints msm_points_of(const SrsEntry& e, size_t first) {
    MsmPoints p;
    p.base = e.points + first;
    p.level_stride = e.pre_c ? e.n : 0;
    p.pre_c = e.pre_c;
    p.pre_W = e.pre_W;
    return p;
}

static int msm_host_scalars(bpk_ctx* ctx, const SrsEntry& srs, const uint64_t* scalars, size_t n_scalars,
                            unsigned rshift, uint64_t out_xyz[18]) {
    const size_t n_points = srs.n;
    const MsmPoints d_points = msm_points_of(srs, 0);
    size_t n = n_points < n_scalars ? n_points : n_scalars;  // zip truncation (msm.rs:29)
    fr_t* d_scalars;
    BPK_TRY(ws_reserve(ctx, 9, (n ? n : 1) * sizeof(fr_t), (void**)&d_scalars));
    uint64_t* d_out;
    BPK_TRY(ws_reserve(ctx, 10, 18 * sizeof(uint64_t), (void**)&d_out));
    BPK_TRY(msm_run_from_host(ctx, d_points, scalars, d_scalars, n, rshift, true, d_out));
    BPK_CUDA(cudaMemcpyAsync(out_xyz, d_out, 18 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    BPK_CUDA(cudaStreamSynchronize(ctx->stream));
    return BPK_OK;
}

extern "C" int bpk_bucket_msm(bpk_ctx* ctx, uint64_t handle, const uint64_t* scalars_mont, size_t n_scalars,
                              size_t b, size_t c, uint64_t out_xyz[18]) {
    if (!ctx || !out_xyz || (n_scalars && !scalars_mont)) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    unsigned rshift = 0;
    BPK_TRY(window_to_shift(b, c, &rshift));
    auto it = ctx->srs.find(handle);
    if (it == ctx->srs.end()) return BPK_ERR_INVALID_ARG;
    BPK_CUDA(cudaSetDevice(ctx->device));
    return msm_host_scalars(ctx, it->second, scalars_mont, n_scalars, rshift, out_xyz);
}

extern "C" int bpk_msm_g1(bpk_ctx* ctx, uint64_t handle, const uint64_t* scalars_mont, size_t n_scalars,
                          uint64_t out_xyz[18]) {
    return bpk_bucket_msm(ctx, handle, scalars_mont, n_scalars, 256, 4, out_xyz);
}

extern "C" int bpk_msm_g1_points(bpk_ctx* ctx, const uint64_t* points_xyz, size_t n_points,
                                 const uint64_t* scalars_mont, size_t n_scalars, uint64_t out_xyz[18]) {
    if (!ctx || !out_xyz || (n_points && !points_xyz) || (n_scalars && !scalars_mont)) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    size_t n = n_points < n_scalars ? n_points : n_scalars;
    uint64_t h = 0;
    BPK_TRY(bpk_srs_load(ctx, points_xyz, n, &h));
    int s = bpk_bucket_msm(ctx, h, scalars_mont, n, 256, 4, out_xyz);
    bpk_srs_free(ctx, h);
    return s;
}

extern "C" int bpk_msm_g1_dev(bpk_ctx* ctx, uint64_t handle, size_t first, const void* d_scalars_mont, size_t n,
                              int normalise, void* d_out_xyz) {
    if (!ctx || !d_out_xyz || (n && !d_scalars_mont)) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    auto it = ctx->srs.find(handle);
    if (it == ctx->srs.end()) return BPK_ERR_INVALID_ARG;
    if (first > it->second.n || n > it->second.n - first) return BPK_ERR_INVALID_ARG;
    BPK_CUDA(cudaSetDevice(ctx->device));
    return msm_run(ctx, msm_points_of(it->second, first), (const fr_t*)d_scalars_mont, n, 0, normalise != 0,
                   (uint64_t*)d_out_xyz);
}

extern "C" int bpk_msm_g1_from_host(bpk_ctx* ctx, uint64_t handle, size_t first, const uint64_t* scalars_mont,
                                    size_t n, int normalise, void* d_out_xyz) {
    if (!ctx || !d_out_xyz || (n && !scalars_mont)) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    auto it = ctx->srs.find(handle);
    if (it == ctx->srs.end()) return BPK_ERR_INVALID_ARG;
    if (first > it->second.n || n > it->second.n - first) return BPK_ERR_INVALID_ARG;
    BPK_CUDA(cudaSetDevice(ctx->device));
    fr_t* d_scalars;
    BPK_TRY(ws_reserve(ctx, 9, (n ? n : 1) * sizeof(fr_t), (void**)&d_scalars));
    return msm_run_from_host(ctx, msm_points_of(it->second, first), scalars_mont, d_scalars, n, 0, normalise != 0,
                             (uint64_t*)d_out_xyz);
}

extern "C" int bpk_msm_g1_dev_batch(bpk_ctx* ctx, uint64_t handle, size_t count, const void* const* d_scalars_mont,
                                    const size_t* first, const size_t* n, int normalise, void* d_out_xyz) {
    if (!ctx || !d_out_xyz || (count && (!d_scalars_mont || !first || !n))) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    auto it = ctx->srs.find(handle);
    if (it == ctx->srs.end()) return BPK_ERR_INVALID_ARG;
    for (size_t i = 0; i < count; i++) {
        if (n[i] && !d_scalars_mont[i]) return BPK_ERR_INVALID_ARG;
        if (first[i] > it->second.n || n[i] > it->second.n - first[i]) return BPK_ERR_INVALID_ARG;
    }
    BPK_CUDA(cudaSetDevice(ctx->device));
    uint64_t* out = (uint64_t*)d_out_xyz;
    // very large MSMs gain nothing from running side by side (their latency-bound stages are a per-mille of the
    // accumulation) and each lane would hold its own ~20 GB workspace: run them one after the other
    bool huge = false;
    for (size_t i = 0; i < count; i++) huge = huge || n[i] >= ((size_t)1 << 23);
    if (count <= 1 || ctx->opt_msm_lanes <= 1 || huge) {
        for (size_t i = 0; i < count; i++)
            BPK_TRY(msm_run(ctx, msm_points_of(it->second, first[i]), (const fr_t*)d_scalars_mont[i], n[i], 0,
                            normalise != 0, out + 18 * i));
        return BPK_OK;
    }
    // Each MSM goes to its own stream and workspace bank.  The accumulate kernels share the SMs (they are
    // arithmetic-bound, so nothing is lost), while the latency-bound stages of one MSM -- sort, merge, bucket
    // reduction, normalisation -- run under the accumulation of the others instead of idling the GPU.
    const int lanes = (int)(ctx->opt_msm_lanes < MSM_LANES ? ctx->opt_msm_lanes : MSM_LANES);
    if (!ctx->lane_fork) BPK_CUDA(cudaEventCreateWithFlags(&ctx->lane_fork, cudaEventDisableTiming));
    for (int l = 0; l < lanes; l++) {
        if (!ctx->lane_stream[l]) BPK_CUDA(cudaStreamCreateWithFlags(&ctx->lane_stream[l], cudaStreamNonBlocking));
        if (!ctx->lane_done[l]) BPK_CUDA(cudaEventCreateWithFlags(&ctx->lane_done[l], cudaEventDisableTiming));
    }
    cudaStream_t main_stream = ctx->stream;
    BPK_CUDA(cudaEventRecord(ctx->lane_fork, main_stream));
    int status = BPK_OK;
    for (size_t i = 0; i < count && status == BPK_OK; i++) {
        const int l = (int)(i % lanes);
        if (i < (size_t)lanes) {
            cudaError_t e = cudaStreamWaitEvent(ctx->lane_stream[l], ctx->lane_fork, 0);
            if (e != cudaSuccess) { status = cuda_fail(ctx, e, "lane fork", __FILE__, __LINE__); break; }
        }
        ctx->stream = ctx->lane_stream[l];
        ctx->ws_bank = l + 1;
        status = msm_run(ctx, msm_points_of(it->second, first[i]), (const fr_t*)d_scalars_mont[i], n[i], 0,
                         normalise != 0, out + 18 * i);
        ctx->stream = main_stream;
        ctx->ws_bank = 0;
    }
    for (int l = 0; l < lanes; l++) {  // join, also on the error path so that the caller's stream stays ordered
        cudaError_t e = cudaEventRecord(ctx->lane_done[l], ctx->lane_stream[l]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(main_stream, ctx->lane_done[l], 0);
        if (e != cudaSuccess && status == BPK_OK) status = cuda_fail(ctx, e, "lane join", __FILE__, __LINE__);
    }
    return status;
}

extern "C" int bpk_g1_sum(bpk_ctx* ctx, const uint64_t* points_xyz, size_t n, uint64_t out_xyz[18]) {
    if (!ctx || !out_xyz || (n && !points_xyz) || n > (1u << 20)) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    BPK_CUDA(cudaSetDevice(ctx->device));
    uint64_t* d_in;
    BPK_TRY(ws_reserve(ctx, 8, (n ? n : 1) * 18 * sizeof(uint64_t), (void**)&d_in));
    uint64_t* d_out;
    BPK_TRY(ws_reserve(ctx, 10, 18 * sizeof(uint64_t), (void**)&d_out));
    if (n) BPK_CUDA(cudaMemcpyAsync(d_in, points_xyz, n * 18 * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    BPK_TRY(g1_sum_run(ctx, d_in, n, d_out));
    BPK_CUDA(cudaMemcpyAsync(out_xyz, d_out, 18 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    BPK_CUDA(cudaStreamSynchronize(ctx->stream));
    return BPK_OK;
}

extern "C" int bpk_g1_sum_dev(bpk_ctx* ctx, const void* d_points_xyz, size_t n, void* d_out_xyz) {
    if (!ctx || !d_out_xyz || (n && !d_points_xyz) || n > (1u << 20)) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    BPK_CUDA(cudaSetDevice(ctx->device));
    return g1_sum_run(ctx, (const uint64_t*)d_points_xyz, n, (uint64_t*)d_out_xyz);
}

// ------------------------------------------------------------------------------------------------
// NTT
// ------------------------------------------------------------------------------------------------
static int ntt_host(bpk_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n, size_t batch, bool inverse,
                    const uint64_t* shift) {
    if (!ctx || !in || !out) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    if (shift && !limbs_canonical<fr_t>(shift)) return BPK_ERR_INVALID_ARG;
    if (n == 0 || (n & (n - 1)) != 0) return BPK_ERR_NOT_POW2;
    if (n > ((size_t)1 << NTT_MAX_LOG)) return BPK_ERR_TOO_LARGE;
    if (batch == 0) return BPK_OK;
    BPK_CUDA(cudaSetDevice(ctx->device));
    size_t bytes = n * batch * sizeof(fr_t);
    fr_t* d_buf;
    BPK_TRY(ws_reserve(ctx, 9, bytes, (void**)&d_buf));
    BPK_TRY(upload_host(ctx, d_buf, in, bytes, ctx->stream));
    fr_t sh;
    if (shift) sh = fr_from_host(shift);
    BPK_TRY(ntt_run(ctx, d_buf, d_buf, n, batch, inverse, shift ? &sh : nullptr));
    BPK_CUDA(cudaMemcpyAsync(out, d_buf, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    BPK_CUDA(cudaStreamSynchronize(ctx->stream));
    return BPK_OK;
}

extern "C" int bpk_ntt_fr(bpk_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n, size_t batch) {
    return ntt_host(ctx, in, out, n, batch, false, nullptr);
}
extern "C" int bpk_intt_fr(bpk_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n, size_t batch) {
    return ntt_host(ctx, in, out, n, batch, true, nullptr);
}
extern "C" int bpk_coset_ntt_fr(bpk_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n, size_t batch,
                                const uint64_t shift_mont[4]) {
    if (!shift_mont) return BPK_ERR_INVALID_ARG;
    return ntt_host(ctx, in, out, n, batch, false, shift_mont);
}
extern "C" int bpk_coset_intt_fr(bpk_ctx* ctx, const uint64_t* in, uint64_t* out, size_t n, size_t batch,
                                 const uint64_t shift_mont[4]) {
    if (!shift_mont) return BPK_ERR_INVALID_ARG;
    return ntt_host(ctx, in, out, n, batch, true, shift_mont);
}

extern "C" int bpk_ntt_fr_dev(bpk_ctx* ctx, const void* d_in, void* d_out, size_t n, size_t batch, int flags,
                              const uint64_t* shift_mont) {
    if (!ctx || !d_in || !d_out) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    if ((flags & 2) && (!shift_mont || !limbs_canonical<fr_t>(shift_mont))) return BPK_ERR_INVALID_ARG;
    BPK_CUDA(cudaSetDevice(ctx->device));
    fr_t sh;
    if (flags & 2) sh = fr_from_host(shift_mont);
    return ntt_run(ctx, (const fr_t*)d_in, (fr_t*)d_out, n, batch, (flags & 1) != 0, (flags & 2) ? &sh : nullptr);
}

extern "C" int bpk_poly_mul_fr(bpk_ctx* ctx, const uint64_t* a, size_t la, const uint64_t* b, size_t lb,
                               uint64_t* out) {
    if (!ctx || !a || !b || !out || la == 0 || lb == 0) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    // D = find_next_power_of_two(deg a, deg b) (utils.rs:54-61): smallest power of two >= la + lb - 1
    size_t target = la + lb - 1;
    size_t D = 1;
    while (D < target) D <<= 1;
    if (D > ((size_t)1 << NTT_MAX_LOG)) return BPK_ERR_TOO_LARGE;
    BPK_CUDA(cudaSetDevice(ctx->device));
    fr_t* d_buf;
    BPK_TRY(ws_reserve(ctx, 9, 2 * D * sizeof(fr_t), (void**)&d_buf));
    BPK_CUDA(cudaMemsetAsync(d_buf, 0, 2 * D * sizeof(fr_t), ctx->stream));
    BPK_TRY(upload_host(ctx, d_buf, a, la * sizeof(fr_t), ctx->stream));
    BPK_TRY(upload_host(ctx, d_buf + D, b, lb * sizeof(fr_t), ctx->stream));
    BPK_TRY(ntt_run(ctx, d_buf, d_buf, D, 2, false, nullptr));      // both operands to evaluation form
    BPK_TRY(pointwise_mul(ctx, d_buf, d_buf + D, D));               // polynomial.rs:262-266
    BPK_TRY(ntt_run(ctx, d_buf, d_buf, D, 1, true, nullptr));       // i_ntt_381 (polynomial.rs:270)
    BPK_CUDA(cudaMemcpyAsync(out, d_buf, target * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    BPK_CUDA(cudaStreamSynchronize(ctx->stream));
    return BPK_OK;
}

extern "C" int bpk_poly_mul_fr_dev(bpk_ctx* ctx, const void* d_a, size_t la, const void* d_b, size_t lb, void* d_out) {
    if (!ctx || !d_a || !d_b || !d_out || la == 0 || lb == 0) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    size_t target = la + lb - 1;
    size_t D = 1;
    while (D < target) D <<= 1;
    if (D > ((size_t)1 << NTT_MAX_LOG)) return BPK_ERR_TOO_LARGE;
    BPK_CUDA(cudaSetDevice(ctx->device));
    fr_t* d_buf;
    BPK_TRY(ws_reserve(ctx, 9, 2 * D * sizeof(fr_t), (void**)&d_buf));
    BPK_CUDA(cudaMemsetAsync(d_buf, 0, 2 * D * sizeof(fr_t), ctx->stream));
    BPK_CUDA(cudaMemcpyAsync(d_buf, d_a, la * sizeof(fr_t), cudaMemcpyDeviceToDevice, ctx->stream));
    BPK_CUDA(cudaMemcpyAsync(d_buf + D, d_b, lb * sizeof(fr_t), cudaMemcpyDeviceToDevice, ctx->stream));
    BPK_TRY(ntt_run(ctx, d_buf, d_buf, D, 2, false, nullptr));
    BPK_TRY(pointwise_mul(ctx, d_buf, d_buf + D, D));
    BPK_TRY(ntt_run(ctx, d_buf, d_buf, D, 1, true, nullptr));
    BPK_CUDA(cudaMemcpyAsync(d_out, d_buf, target * sizeof(fr_t), cudaMemcpyDeviceToDevice, ctx->stream));
    return BPK_OK;
}

// ------------------------------------------------------------------------------------------------
// device memory for callers without a CUDA runtime binding of their own (Rust / C++ hosts of the
// device-resident entry points); all transfers are ordered on the context's stream
// ------------------------------------------------------------------------------------------------
extern "C" int bpk_dev_alloc(bpk_ctx* ctx, size_t bytes, void** d_out) {
    if (!ctx || !d_out) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    BPK_CUDA(cudaSetDevice(ctx->device));
    *d_out = nullptr;
    bytes = (bytes + 255) / 256 * 256;
    if (bytes == 0) bytes = 256;
    auto it = ctx->dev_pool.find(bytes);
    if (it != ctx->dev_pool.end() && !it->second.empty()) {  // reuse is ordered by the context's stream
        *d_out = it->second.back();
        it->second.pop_back();
    } else {
        cudaError_t e = cudaMalloc(d_out, bytes);
        if (e == cudaErrorMemoryAllocation) {  // give the cached blocks back and retry once
            cudaGetLastError();
            for (auto& kv : ctx->dev_pool)
                for (void* p : kv.second) cudaFree(p);
            ctx->dev_pool.clear();
            e = cudaMalloc(d_out, bytes);
            if (e == cudaErrorMemoryAllocation) {
                cudaGetLastError();
                return BPK_ERR_OOM;
            }
        }
        BPK_CUDA(e);
    }
    ctx->dev_live[*d_out] = bytes;
    return BPK_OK;
}

extern "C" int bpk_dev_free(bpk_ctx* ctx, void* d_ptr) {
    if (!ctx) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    if (!d_ptr) return BPK_OK;
    auto it = ctx->dev_live.find(d_ptr);
    if (it == ctx->dev_live.end()) return BPK_ERR_INVALID_ARG;  // not from bpk_dev_alloc (or freed twice)
    ctx->dev_pool[it->second].push_back(d_ptr);
    ctx->dev_live.erase(it);
    return BPK_OK;
}

extern "C" int bpk_dev_upload(bpk_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
    if (!ctx || (bytes && (!d_dst || !h_src))) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    BPK_CUDA(cudaSetDevice(ctx->device));
    if (bytes) BPK_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return BPK_OK;
}

extern "C" int bpk_dev_download(bpk_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
    if (!ctx || (bytes && (!h_dst || !d_src))) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    BPK_CUDA(cudaSetDevice(ctx->device));
    if (bytes) BPK_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    BPK_CUDA(cudaStreamSynchronize(ctx->stream));
    return BPK_OK;
}

extern "C" int bpk_dev_copy(bpk_ctx* ctx, void* d_dst, const void* d_src, size_t bytes) {
    if (!ctx || (bytes && (!d_dst || !d_src))) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    BPK_CUDA(cudaSetDevice(ctx->device));
    if (bytes) BPK_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return BPK_OK;
}

extern "C" int bpk_dev_zero(bpk_ctx* ctx, void* d_dst, size_t bytes) {
    if (!ctx || (bytes && !d_dst)) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    BPK_CUDA(cudaSetDevice(ctx->device));
    if (bytes) BPK_CUDA(cudaMemsetAsync(d_dst, 0, bytes, ctx->stream));
    return BPK_OK;
}

// ------------------------------------------------------------------------------------------------
// device-resident Fr vector / polynomial primitives (prover rounds)
// ------------------------------------------------------------------------------------------------
extern "C" int bpk_fr_vec_op(bpk_ctx* ctx, int op, const void* d_a, const void* d_b, const uint64_t* scalar_mont,
                             void* d_out, size_t n) {
    if (!ctx || op < 0 || op > 6 || (n && (!d_a || !d_out))) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    const bool needs_b = op == 0 || op == 1 || op == 2 || op == 4 || op == 6;
    const bool needs_s = op >= 3;
    if (n && ((needs_b && !d_b) || (needs_s && !scalar_mont))) return BPK_ERR_INVALID_ARG;
    BPK_CUDA(cudaSetDevice(ctx->device));
    fr_t s = needs_s ? fr_from_host(scalar_mont) : fr_t::zero();
    return fr_vec_op(ctx, op, (const fr_t*)d_a, (const fr_t*)d_b, s, (fr_t*)d_out, n);
}

extern "C" int bpk_fr_scale_powers(bpk_ctx* ctx, const void* d_a, const uint64_t g_mont[4], const uint64_t c0_mont[4],
                                   void* d_out, size_t n) {
    if (!ctx || !g_mont || !c0_mont || (n && (!d_a || !d_out))) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    BPK_CUDA(cudaSetDevice(ctx->device));
    return fr_scale_powers(ctx, (const fr_t*)d_a, fr_from_host(g_mont), fr_from_host(c0_mont), (fr_t*)d_out, n);
}

extern "C" int bpk_fr_poly_eval(bpk_ctx* ctx, const void* d_coeffs, size_t n, const uint64_t x_mont[4],
                                uint64_t out_mont[4]) {
    if (!ctx || !x_mont || !out_mont || (n && !d_coeffs)) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    BPK_CUDA(cudaSetDevice(ctx->device));
    fr_t* d_out;
    BPK_TRY(ws_reserve(ctx, 10, sizeof(fr_t) * 8, (void**)&d_out));
    BPK_TRY(fr_poly_eval(ctx, (const fr_t*)d_coeffs, n, fr_from_host(x_mont), d_out));
    BPK_CUDA(cudaMemcpyAsync(out_mont, d_out, sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    BPK_CUDA(cudaStreamSynchronize(ctx->stream));
    return BPK_OK;
}

extern "C" int bpk_fr_poly_eval_many(bpk_ctx* ctx, size_t count, const void* const* d_coeffs, const size_t* n,
                                     const uint64_t x_mont[4], uint64_t* out_mont) {
    if (!ctx || !x_mont || (count && (!d_coeffs || !n || !out_mont)) || count > 64) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    for (size_t i = 0; i < count; i++)
        if (n[i] && !d_coeffs[i]) return BPK_ERR_INVALID_ARG;
    if (count == 0) return BPK_OK;
    BPK_CUDA(cudaSetDevice(ctx->device));
    fr_t* d_out;
    BPK_TRY(ws_reserve(ctx, 10, sizeof(fr_t) * 64, (void**)&d_out));
    BPK_TRY(fr_poly_eval_many(ctx, count, (const fr_t* const*)d_coeffs, n, fr_from_host(x_mont), d_out));
    BPK_CUDA(cudaMemcpyAsync(out_mont, d_out, count * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    BPK_CUDA(cudaStreamSynchronize(ctx->stream));
    return BPK_OK;
}

extern "C" int bpk_fr_poly_div_linear(bpk_ctx* ctx, const void* d_coeffs, size_t n, const uint64_t root_mont[4],
                                      void* d_quotient) {
    if (!ctx || !root_mont || (n >= 2 && (!d_coeffs || !d_quotient))) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    BPK_CUDA(cudaSetDevice(ctx->device));
    return fr_poly_div_linear(ctx, (const fr_t*)d_coeffs, n, fr_from_host(root_mont), (fr_t*)d_quotient);
}

extern "C" int bpk_fr_poly_div_vanishing(bpk_ctx* ctx, const void* d_coeffs, size_t len, size_t n, void* d_quotient) {
    if (!ctx || n == 0 || (len > n && (!d_coeffs || !d_quotient))) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    BPK_CUDA(cudaSetDevice(ctx->device));
    return fr_poly_div_vanishing(ctx, (const fr_t*)d_coeffs, len, n, (fr_t*)d_quotient);
}

extern "C" int bpk_plonk_grand_product(bpk_ctx* ctx, const void* d_a, const void* d_b, const void* d_c, const void* d_s1,
                                       const void* d_s2, const void* d_s3, size_t n, const uint64_t beta[4],
                                       const uint64_t gamma[4], const uint64_t k1[4], const uint64_t k2[4], void* d_z) {
    if (!ctx || !d_a || !d_b || !d_c || !d_s1 || !d_s2 || !d_s3 || !beta || !gamma || !k1 || !k2 || !d_z)
        return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    BPK_CUDA(cudaSetDevice(ctx->device));
    return plonk_grand_product(ctx, (const fr_t*)d_a, (const fr_t*)d_b, (const fr_t*)d_c, (const fr_t*)d_s1,
                               (const fr_t*)d_s2, (const fr_t*)d_s3, n, fr_from_host(beta), fr_from_host(gamma),
                               fr_from_host(k1), fr_from_host(k2), (fr_t*)d_z);
}

extern "C" int bpk_plonk_quotient_evals(bpk_ctx* ctx, const void* d_witness_evals, const void* d_circuit_evals,
                                        size_t domain, size_t n,
                                        const uint64_t beta[4], const uint64_t gamma[4], const uint64_t alpha[4],
                                        const uint64_t k1[4], const uint64_t k2[4], const uint64_t* zh_inv_mont,
                                        void* d_out) {
    if (!ctx || !d_witness_evals || !d_circuit_evals || !beta || !gamma || !alpha || !k1 || !k2 || !zh_inv_mont || !d_out) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    if (n == 0 || domain < n || domain / n > 64) return BPK_ERR_INVALID_ARG;
    BPK_CUDA(cudaSetDevice(ctx->device));
    fr_t zh[64];
    for (size_t i = 0; i < domain / n; ++i) zh[i] = fr_from_host(zh_inv_mont + 4 * i);
    return plonk_quotient_evals(ctx, (const fr_t*)d_witness_evals, (const fr_t*)d_circuit_evals, domain, n, fr_from_host(beta), fr_from_host(gamma),
                                fr_from_host(alpha), fr_from_host(k1), fr_from_host(k2), zh, (fr_t*)d_out);
}

extern "C" int bpk_plonk_quotient_evals_shard(bpk_ctx* ctx, const void* d_witness_evals, const void* d_circuit_evals,
                                              size_t points, unsigned zh_period, const uint64_t beta[4],
                                              const uint64_t gamma[4], const uint64_t alpha[4], const uint64_t k1[4],
                                              const uint64_t k2[4], const uint64_t* zh_inv_mont, void* d_out) {
    if (!ctx || !d_witness_evals || !d_circuit_evals || !beta || !gamma || !alpha || !k1 || !k2 || !zh_inv_mont || !d_out)
        return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    if (points == 0 || zh_period == 0 || zh_period > 64 || (zh_period & (zh_period - 1))) return BPK_ERR_INVALID_ARG;
    BPK_CUDA(cudaSetDevice(ctx->device));
    fr_t zh[64];
    for (unsigned i = 0; i < zh_period; ++i) zh[i] = fr_from_host(zh_inv_mont + 4 * i);
    return plonk_quotient_evals_shard(ctx, (const fr_t*)d_witness_evals, (const fr_t*)d_circuit_evals, points, zh_period,
                                      fr_from_host(beta), fr_from_host(gamma), fr_from_host(alpha), fr_from_host(k1),
                                      fr_from_host(k2), zh, (fr_t*)d_out);
}

extern "C" int bpk_fr_fold(bpk_ctx* ctx, const void* d_in, size_t rows, size_t row_stride, size_t len, size_t m,
                           const uint64_t s_mont[4], void* d_out) {
    if (!ctx || !s_mont || (rows && m && (!d_in || !d_out))) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    if (len > row_stride || (rows && m && len == 0)) return BPK_ERR_INVALID_ARG;
    BPK_CUDA(cudaSetDevice(ctx->device));
    return fr_fold(ctx, (const fr_t*)d_in, rows, row_stride, len, m, fr_from_host(s_mont), (fr_t*)d_out);
}

// ------------------------------------------------------------------------------------------------
// host utility: keccak-f[1600] for the Fiat-Shamir transcript of the host layers (merlin / STROBE-128 sits on it;
// a few dozen permutations per proof, which cost ~10 ms when done in interpreted Python)
// ------------------------------------------------------------------------------------------------
extern "C" void bpk_keccak_f1600(uint64_t s[25]) {
    static const int rho[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    uint64_t lfsr = 1;
    for (int round = 0; round < 24; round++) {
        uint64_t rc = 0;
        for (int j = 0; j < 7; j++) {
            if (lfsr & 1) rc |= (uint64_t)1 << ((1 << j) - 1);
            lfsr <<= 1;
            if (lfsr & 0x100) lfsr ^= 0x171;
        }
        uint64_t c[5], m[25];
        for (int x = 0; x < 5; x++) c[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
        for (int x = 0; x < 5; x++) {
            uint64_t r = c[(x + 1) % 5];
            uint64_t d = c[(x + 4) % 5] ^ ((r << 1) | (r >> 63));
            for (int y = 0; y < 25; y += 5) s[x + y] ^= d;
        }
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) {
                int k = rho[x + 5 * y];
                uint64_t v = s[x + 5 * y];
                m[y + 5 * ((2 * x + 3 * y) % 5)] = k ? (v << k) | (v >> (64 - k)) : v;
            }
        for (int y = 0; y < 25; y += 5)
            for (int x = 0; x < 5; x++) s[x + y] = m[x + y] ^ (~m[(x + 1) % 5 + y] & m[(x + 2) % 5 + y]);
        s[0] ^= rc;
    }
}

// ------------------------------------------------------------------------------------------------
// host utility: the synthetic benchmark circuit (SURVEY 8d C4 family) as pre-processed columns.  Pure host code --
// the interpreted generator in synthetic.py needs ~90 s for 2^24 rows, this one ~2 s; tests/test_synthetic_cpu.py
// checks that both produce the same columns.
// ------------------------------------------------------------------------------------------------
namespace {
typedef unsigned __int128 u128h;
struct HFr {  // 4 x u64 Montgomery residue mod q (scalar.rs:22), host arithmetic
    uint64_t l[4];
};
const uint64_t HQ[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
const uint64_t HQ_INV = 0xfffffffeffffffffull;  // -q^-1 mod 2^64 (scalar.rs:164)
const HFr HR2 = {{0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull}};
inline void hfr_cond_sub(uint64_t r[4]) {
    uint64_t t[4], borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128h d = (u128h)r[i] - HQ[i] - borrow;
        t[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    if (!borrow)
        for (int i = 0; i < 4; i++) r[i] = t[i];
}
inline HFr hfr_mul(const HFr& a, const HFr& b) {
    uint64_t t[9] = {0};
    for (int i = 0; i < 4; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < 4; j++) {
            u128h p = (u128h)a.l[j] * b.l[i] + t[i + j] + carry;
            t[i + j] = (uint64_t)p;
            carry = (uint64_t)(p >> 64);
        }
        u128h s = (u128h)t[i + 4] + carry;
        t[i + 4] = (uint64_t)s;
        t[i + 5] += (uint64_t)(s >> 64);
        const uint64_t m = t[i] * HQ_INV;
        carry = 0;
        for (int j = 0; j < 4; j++) {
            u128h p = (u128h)m * HQ[j] + t[i + j] + carry;
            t[i + j] = (uint64_t)p;
            carry = (uint64_t)(p >> 64);
        }
        s = (u128h)t[i + 4] + carry;
        t[i + 4] = (uint64_t)s;
        t[i + 5] += (uint64_t)(s >> 64);
    }
    HFr r = {{t[4], t[5], t[6], t[7]}};
    hfr_cond_sub(r.l);
    return r;
}
inline HFr hfr_add(const HFr& a, const HFr& b) {
    HFr r;
    uint64_t carry = 0;
    for (int i = 0; i < 4; i++) {
        u128h s = (u128h)a.l[i] + b.l[i] + carry;
        r.l[i] = (uint64_t)s;
        carry = (uint64_t)(s >> 64);
    }
    hfr_cond_sub(r.l);  // 2q < 2^256: no carry out
    return r;
}
inline HFr hfr_from_u64(uint64_t v) {
    HFr x = {{v, 0, 0, 0}};
    return hfr_mul(x, HR2);
}
inline HFr hfr_neg_one() {  // q - 1 in Montgomery form = -(R mod q)
    HFr one = hfr_from_u64(1), r;
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128h d = (u128h)HQ[i] - one.l[i] - borrow;
        r.l[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return r;
}
}  // namespace

// columns (each n x 4 u64, Montgomery): [0..4] QL QR QM QO QC, [5..7] S1 S2 S3, [8..10] A B C; public_out: the public
// input (canonical limbs).  Same circuit as synthetic.chain_circuit(n, gates, seed).
extern "C" int bpk_synthetic_chain_circuit(size_t n, size_t gates, uint64_t seed, uint64_t* const columns[11],
                                           uint64_t public_out[4]) {
    if (!columns || !public_out || n < 2 || (n & (n - 1)) || gates < 2 || gates > n || n > ((size_t)1 << 32))
        return BPK_ERR_INVALID_ARG;
    for (int k = 0; k < 11; k++)
        if (!columns[k]) return BPK_ERR_INVALID_ARG;
    const size_t m = gates - 1;
    HFr* col[11];
    for (int k = 0; k < 11; k++) {
        col[k] = reinterpret_cast<HFr*>(columns[k]);
        memset(col[k], 0, n * sizeof(HFr));
    }
    HFr *ql = col[0], *qr = col[1], *qm = col[2], *qo = col[3], *s1 = col[5], *s2 = col[6], *s3 = col[7], *A = col[8],
        *B = col[9], *C = col[10];
    // roots of unity: w = ROOT_OF_UNITY^(2^32 / n) (utils.rs:39-43), then sequential powers (utils.rs:45-52)
    HFr w = {{0xb9b58d8c5f0e466aull, 0x5b1b4c801819d7ecull, 0x0af53ae352a31e64ull, 0x5bf3adda19e9b27bull}};
    for (size_t k = n; k < ((size_t)1 << 32); k <<= 1) w = hfr_mul(w, w);
    std::vector<HFr> roots(n);
    roots[0] = hfr_from_u64(1);
    for (size_t i = 1; i < n; i++) roots[i] = hfr_mul(roots[i - 1], w);
    const HFr one = hfr_from_u64(1), minus_one = hfr_neg_one(), two = hfr_from_u64(2), three = hfr_from_u64(3);
    uint64_t state = seed * 0x9E3779B97F4A7C15ull + 1;
    HFr c_prev = hfr_from_u64(3 + seed % 1000);
    for (size_t k = 1; k <= m; k++) {
        state = state * 6364136223846793005ull + 1442695040888963407ull;
        const HFr y = hfr_from_u64((state >> 20) + 2);
        A[k] = c_prev;
        B[k] = y;
        if (k & 1) {
            c_prev = hfr_mul(c_prev, y);
            qm[k] = minus_one;
            qo[k] = one;
        } else {
            c_prev = hfr_add(c_prev, y);
            ql[k] = minus_one;
            qr[k] = minus_one;
            qo[k] = one;
        }
        C[k] = c_prev;
    }
    A[0] = c_prev;
    ql[0] = one;
    {   // canonical form of the public input
        HFr raw = {{1, 0, 0, 0}};
        HFr canon = hfr_mul(c_prev, raw);
        memcpy(public_out, canon.l, 32);
    }
    // sigma: identity labels (col + 1) w^row, then the copy-constraint cycles
    for (size_t i = 0; i < n; i++) {
        s1[i] = roots[i];
        s2[i] = hfr_mul(roots[i], two);
        s3[i] = hfr_mul(roots[i], three);
    }
    HFr* s[3] = {s1, s2, s3};
    for (size_t k = 1; k < m; k++) {   // c_k: (O, k) <-> (L, k + 1)
        s[0][k + 1] = hfr_mul(roots[k], three);
        s[2][k] = roots[k + 1];
    }
    s[0][0] = hfr_mul(roots[m], three);   // out: (L, 0) <-> (O, m)
    s[2][m] = roots[0];
    // all unused cells form one cycle in the order (R,0), (O,0), then rows m+1.. in row-major order
    const size_t unused = 2 + 3 * (n - 1 - m);
    auto cell = [&](size_t i, int& c, size_t& r) {
        if (i < 2) {
            c = (int)i + 1;
            r = 0;
        } else {
            c = (int)((i - 2) % 3);
            r = m + 1 + (i - 2) / 3;
        }
    };
    const HFr mult[3] = {one, two, three};
    for (size_t i = 0; i < unused; i++) {
        int c, nc;
        size_t r, nr;
        cell(i, c, r);
        cell((i + 1) % unused, nc, nr);
        s[nc][nr] = hfr_mul(roots[r], mult[c]);
    }
    return BPK_OK;
}

// ------------------------------------------------------------------------------------------------
// instrumentation
// ------------------------------------------------------------------------------------------------
extern "C" int bpk_profile_enable(bpk_ctx* ctx, int on) {
    if (!ctx) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    profile_collect(ctx);
    ctx->profiling = on != 0;
    return BPK_OK;
}
extern "C" int bpk_profile_reset(bpk_ctx* ctx) {
    if (!ctx) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    profile_collect(ctx);
    ctx->stats.clear();
    ctx->launches = 0;
    return BPK_OK;
}
extern "C" int bpk_profile_get(bpk_ctx* ctx, const char* name, double* ms_out, uint64_t* launches_out) {
    if (!ctx || !name) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    profile_collect(ctx);
    auto it = ctx->stats.find(name);
    double ms = 0;
    uint64_t l = 0;
    if (it != ctx->stats.end()) {
        ms = it->second.ms;
        l = it->second.launches;
    }
    if (ms_out) *ms_out = ms;
    if (launches_out) *launches_out = l;
    return BPK_OK;
}
extern "C" uint64_t bpk_launch_count(bpk_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int bpk_msm_last_plan(bpk_ctx* ctx, unsigned out[4]) {
    if (!ctx || !out) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    out[0] = ctx->last_c;
    out[1] = ctx->last_W;
    out[2] = ctx->last_chunk;
    out[3] = ctx->last_buckets;
    return BPK_OK;
}

extern "C" int bpk_msm_last_stats(bpk_ctx* ctx, uint64_t out[6]) {
    if (!ctx || !out) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    BPK_CUDA(cudaSetDevice(ctx->device));
    uint64_t st[4];
    BPK_TRY(msm_read_stats(ctx, st));
    out[0] = st[0];                 // (bucket, point) entries of the last bucket fill
    out[1] = st[0] - st[1];         // additions done in affine coordinates (pairwise tree)
    out[2] = st[1] - st[2];         // additions left to the XYZZ tail
    out[3] = st[2];                 // non-empty buckets
    out[4] = ctx->last_levels;      // tree levels launched
    out[5] = ctx->last_batch;       // additions per shared inversion (upper bound)
    return BPK_OK;
}

extern "C" int bpk_srs_table_bytes(bpk_ctx* ctx, uint64_t handle, size_t* bytes_out) {
    if (!ctx || !bytes_out) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    auto it = ctx->srs.find(handle);
    if (it == ctx->srs.end()) return BPK_ERR_INVALID_ARG;
    const SrsEntry& e = it->second;
    *bytes_out = (e.pre_c ? (size_t)e.pre_W : 1) * e.n * sizeof(affine_t);
    return BPK_OK;
}

extern "C" int bpk_imad_peak(bpk_ctx* ctx, double* rate, double* seconds) {
    if (!ctx || !rate || !seconds) return BPK_ERR_INVALID_ARG;
    BPK_LOCK(ctx);
    BPK_CUDA(cudaSetDevice(ctx->device));
    return imad_peak_run(ctx, rate, seconds);
}
