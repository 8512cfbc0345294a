// Wide levels of the bucket-reduction tree (msm.cu, K6), in a translation unit of their own: here every Fp product is
// INLINED.  The level kernel is 14 products per thread on registers and nothing else, so the argument moves of an
// out-of-line product (msm.cu) cost more than the code size (2^21 buckets: 2.44 -> 2.05 ms); the batched-affine level
// kernel of msm.cu measures the other way (51.6 against 52.7-53.6 ms).
#include "internal.cuh"

namespace bpk {

namespace {
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
}  // namespace

#ifndef BPK_TREE_MINBLOCKS
#define BPK_TREE_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(128, BPK_TREE_MINBLOCKS) msm_plane_tree_level_kernel(const xyzz_t* __restrict__ in,
                                                                    xyzz_t* __restrict__ out, uint32_t k,
                                                                    size_t nodes_out) {
    // one thread per addition, slot-major (t = s nodes + node): the lanes of a warp do the same kind of work.  The copy
    // T_{k-1}' = S(c1) is the second operand of the node's s = 0 addition and is written by that thread.  (One thread per
    // output slot, node-major, left 1 / (k + 1) of the lanes idle: 50 % at the widest level.)
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nodes_out * k) return;
    const uint32_t s = (uint32_t)(t / nodes_out);
    const size_t node = t - (size_t)s * nodes_out;
    const xyzz_t* c0 = in + 2 * node * k;  // children hold k points each
    const xyzz_t* c1 = c0 + k;
    xyzz_t* o = out + node * (k + 1);
    xyzz_t a = ld_xyzz(c0 + s);
    xyzz_t b = ld_xyzz(c1 + s);
    if (s == 0) st_xyzz(o + k, b);
    // buckets finished by the affine tree arrive as (x, y, 1, 1), and so do their copies in the lowest plane of level 2:
    // their sum needs a third of the products
    const fp_t one = fp_t::one();
    if ((k == 1 || (k == 2 && s == 1)) && a.ZZ == one && a.ZZZ == one && b.ZZ == one && b.ZZZ == one && a.X != b.X)
        a = xyzz_from_affine_sum(a.X, a.Y, b.X, b.Y);
    else
        xyzz_add(a, b);
    st_xyzz(o + s, a);
}

void msm_launch_plane_tree_level(cudaStream_t stream, const xyzz_t* in, xyzz_t* out, uint32_t k, size_t nodes_out) {
    msm_plane_tree_level_kernel<<<(unsigned)((nodes_out * k + 127) / 128), 128, 0, stream>>>(in, out, k, nodes_out);
}

}  // namespace bpk
