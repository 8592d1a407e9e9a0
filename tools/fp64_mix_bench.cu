// Micro-benchmark: issue rate of the FP64 pipe of sm_100a per instruction kind (DFMA / DMUL / DADD / DSETP-select and
// the LM kernel's mix), 16 independent chains per thread, 8 or 12 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_mix_bench.bin tools/fp64_mix_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
// lane-mask variant: only the lanes whose bit is set in `mask` execute the DFMA chains (a divergent branch): does a
// partially active warp instruction cost the FP64 pipe less than a full one?
__global__ void k_masked(double *out, int iters, double b, double c, unsigned mask) {
    double a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i + 1.0;
    if ((mask >> (threadIdx.x & 31)) & 1u) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
        }
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
void run_masked(const char *name, double *out, unsigned mask) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 13, threads = 128, blocks = 148 * 2;
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k_masked<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, mask);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = ms < best ? ms : best;
    }
    printf("DFMA, lanes %-28s (mask %08x, %2d active): %.3f ms\n", name, mask, __builtin_popcount(mask), best);
}

template <int MODE>
__global__ void k(double *out, int iters, double b, double c) {
    double a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i + 1.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) a[i] = fma(a[i], b, c);
            else if (MODE == 1) a[i] = __dmul_rn(a[i], b);
            else if (MODE == 2) a[i] = __dadd_rn(a[i], c);
            else if (MODE == 3) a[i] = (i & 1) ? __dmul_rn(a[i], b) : fma(a[i], b, c);           // 1:1 DFMA : DMUL
            else if (MODE == 4) a[i] = (i % 3 == 2) ? __dmul_rn(a[i], b) : fma(a[i], b, c);      // 2:1 (the LM loop's mix)
            else if (MODE == 5) a[i] = a[i] < c ? b : a[i] + c;                                   // DSETP + select + DADD
        }
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char *name, double *out, int warps_per_sm) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 13, threads = 128, blocks = 148 * warps_per_sm / 4;
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = ms < best ? ms : best;
    }
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double instr = 16.0 * iters * blocks * (threads / 32);   // warp instructions (MODE 5: 16 x (DSETP + DADD + select))
    printf("%-28s %2d warps/SM: %.3f ms, %.3f warp-instr/clk/SM-subpartition (at %d MHz nominal)\n", name, warps_per_sm, best,
           instr / (best * 1e-3) / (clk * 1e3) / (148 * 4), clk / 1000);
}
int main() {
    double *out;
    cudaMalloc(&out, 148 * 16 * 128 * 8);
    run_masked("all", out, 0xffffffffu);
    run_masked("lower half", out, 0x0000ffffu);
    run_masked("every other", out, 0x55555555u);
    run_masked("lower quarter", out, 0x000000ffu);
    run_masked("one per quarter", out, 0x01010101u);
    run_masked("random 22 of 32", out, 0xb7e5d3f6u);
    run_masked("one lane", out, 0x00000001u);
    for (int w : {8, 12, 16}) {
        run<0>("DFMA", out, w); run<1>("DMUL", out, w); run<2>("DADD", out, w);
        run<3>("DFMA:DMUL 1:1", out, w); run<4>("DFMA:DMUL 2:1", out, w); run<5>("DSETP+sel+DADD (x1 counted)", out, w);
    }
    return 0;
}
