// Register-only throughput probes for the integer instruction forms a big-integer multiplier can be
// built from on sm_100a.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_forms imad_forms.cu
// Prints ops / clk / SM (at the reported SM clock) for each form; results in profiles/r1_imad_forms.md.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) probe(uint64_t* out, uint32_t iters, uint32_t seed) {
    uint32_t a = seed + threadIdx.x * 2654435761u + blockIdx.x;
    uint32_t b = a * 747796405u + 2891336453u;
    uint32_t x[12], y[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { x[i] = a + i * 7; y[i] = b ^ (i * 13); }
    for (uint32_t it = 0; it < iters; it++) {
        if (MODE == 0) {  // IMAD.WIDE.U32 Rd, Ra, Rb, RZ : (x,y) = x*y, 12 independent chains
#pragma unroll
            for (int i = 0; i < 12; i++)
                asm volatile("{ .reg .b64 t; mul.wide.u32 t, %0, %1; mov.b64 {%0, %1}, t; }" : "+r"(x[i]), "+r"(y[i]));
        } else if (MODE == 1) {  // IMAD.WIDE.U32 Rd, Ra, Rb, Rd : 64-bit accumulate, no carry flags
#pragma unroll
            for (int i = 0; i < 12; i++)
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(x[i]), "+r"(y[i]) : "r"(a), "r"(b));
        } else if (MODE == 2) {  // IMAD.WIDE.U32.X chains: 2 chains of 6 (carry in and out)
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x[0]), "+r"(x[1]) : "r"(a), "r"(b));
#pragma unroll
            for (int i = 2; i < 12; i += 2)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x[i]), "+r"(x[i + 1]) : "r"(a), "r"(b));
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(y[0]), "+r"(y[1]) : "r"(b), "r"(a));
#pragma unroll
            for (int i = 2; i < 12; i += 2)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(y[i]), "+r"(y[i + 1]) : "r"(b), "r"(a));
        } else if (MODE == 3) {  // IADD3.X chain: 12-word add with carry
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x[0]) : "r"(y[0]));
#pragma unroll
            for (int i = 1; i < 11; i++) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
            asm volatile("addc.u32 %0, %0, %1;" : "+r"(x[11]) : "r"(y[11]));
        } else if (MODE == 4) {  // IMAD 32-bit with addend
#pragma unroll
            for (int i = 0; i < 12; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(y[i]));
        } else if (MODE == 5) {  // IMAD.HI.U32 with addend
#pragma unroll
            for (int i = 0; i < 12; i++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(y[i]));
        } else if (MODE == 6) {  // IADD3 (three-input add, no carry)
#pragma unroll
            for (int i = 0; i < 12; i++) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(x[i]) : "r"(y[i]), "r"(a));
        } else if (MODE == 7) {  // mix: 6 plain IMAD.WIDE(RZ) + 12 IADD3.X folding them in (the "split" row)
            uint32_t pl[6], ph[6];
#pragma unroll
            for (int i = 0; i < 6; i++)
                asm volatile("{ .reg .b64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(pl[i]), "=r"(ph[i]) : "r"(x[2 * i + 1]), "r"(b));
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x[0]) : "r"(pl[0]));
            asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(x[1]) : "r"(ph[0]));
#pragma unroll
            for (int i = 1; i < 6; i++) {
                asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(x[2 * i]) : "r"(pl[i]));
                asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(x[2 * i + 1]) : "r"(ph[i]));
            }
        } else if (MODE == 8) {  // IMAD.WIDE.U32 with carry-OUT only (first link of a chain), 12 independent
#pragma unroll
            for (int i = 0; i < 12; i += 2) {
                uint32_t c;
                asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, 0, 0;"
                             : "+r"(x[i]), "+r"(x[i + 1]), "=r"(c) : "r"(a), "r"(b));
                y[i] += c;
            }
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) s ^= ((uint64_t)x[i] << 32) | y[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static int run(const char* name, double ops_per_iter, int sms, double clk_hz, uint64_t* d_out) {
    const uint32_t iters = 1u << 13;
    const unsigned blocks = sms * 8;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    probe<MODE><<<blocks, 256>>>(d_out, iters / 8, 1);
    CK(cudaEventRecord(e0));
    probe<MODE><<<blocks, 256>>>(d_out, iters, 2);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double rate = ops_per_iter * iters * 256.0 * blocks / (ms * 1e-3);
    printf("| %-62s | %8.3f ms | %10.3e op/s | %6.1f op/clk/SM |\n", name, ms, rate, rate / clk_hz / sms);
    return 0;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    double clk = clk_khz * 1e3;
    printf("device %s, %d SMs, clock attr %.0f MHz; 16 warps/SMSP resident, 12 independent streams per thread\n", p.name,
           p.multiProcessorCount, clk / 1e6);
    uint64_t* d_out;
    CK(cudaMalloc(&d_out, (size_t)p.multiProcessorCount * 8 * 256 * 8));
    int sms = p.multiProcessorCount;
    printf("| form | time | rate | per clk per SM |\n|---|---|---|---|\n");
    run<0>("IMAD.WIDE.U32 Rd,Ra,Rb,RZ (mul.wide)", 12, sms, clk, d_out);
    run<1>("IMAD.WIDE.U32 Rd,Ra,Rb,Rd (mad.lo.cc+madc.hi, 64-bit addend)", 12, sms, clk, d_out);
    run<2>("IMAD.WIDE.U32.X chains (madc.lo.cc+madc.hi.cc)", 12, sms, clk, d_out);
    run<8>("IMAD.WIDE.U32 carry-out only + IADD3.X capture", 6, sms, clk, d_out);
    run<4>("IMAD 32-bit, 32-bit addend (mad.lo.u32)", 12, sms, clk, d_out);
    run<5>("IMAD.HI.U32, 32-bit addend (mad.hi.u32)", 12, sms, clk, d_out);
    run<3>("IADD3.X carry chain (add.cc/addc.cc), adds counted", 12, sms, clk, d_out);
    run<6>("IADD3 three-input add", 12, sms, clk, d_out);
    run<7>("split row: 6 IMAD.WIDE(RZ) + 12 IADD3.X, products counted", 6, sms, clk, d_out);
    return 0;
}
