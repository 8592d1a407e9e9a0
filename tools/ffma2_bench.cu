// Micro-benchmark: does the packed FP32 FMA of sm_100a (fma.rn.f32x2 -> FFMA2) raise FP32 throughput
// per issue slot?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__global__ void k_scalar(float *out, int iters) {
    float a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
    const float b = 1.0000001f, c = 1e-9f;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
    float s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float *out, int iters) {
    u64 a[8];
    for (int i = 0; i < 8; ++i) a[i] = pack2(threadIdx.x * 1e-3f + i, threadIdx.x * 1e-3f + i + 8);
    const u64 b = pack2(1.0000001f, 1.0000001f), c = pack2(1e-9f, 1e-9f);
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma2(a[i], b, c);
    u64 s = 0;
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(s & 0xffff);
}
int main() {
    float *out;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 14, blocks = 148 * 8, threads = 256;
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e9;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k_scalar<<<blocks, threads>>>(out, iters); else k_packed<<<blocks, threads>>>(out, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) best = ms < best ? ms : best;
        }
        const double flops = 2.0 * 16 * (double)iters * blocks * threads;
        printf("%s: %.3f ms, %.1f TFLOP/s\n", mode == 0 ? "FFMA  (scalar)" : "FFMA2 (f32x2) ", best, flops / (best * 1e-3) / 1e12);
    }
    return 0;
}
