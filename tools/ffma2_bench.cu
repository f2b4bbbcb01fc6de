// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100a) issue and flop rates.
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE> __global__ void k(float *out, int iters) {
    float x[8]; unsigned long long y[8]; int z[4] = {(int)threadIdx.x, 1, 2, 3};
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-9f + i; y[i] = pk(x[i], x[i] + 0.5f); }
    const float a = 0.999999f, c = 1e-7f; const unsigned long long a2 = pk(a, a), c2 = pk(c, c);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, c);
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = fma2(y[i], a2, c2);
            } else if (MODE == 2) {       // 8 FFMA + 8 integer ops
#pragma unroll
                for (int i = 0; i < 8; ++i) { x[i] = fmaf(x[i], a, c); z[i & 3] = (z[i & 3] ^ (z[(i + 1) & 3] + i)) + it; }
            } else {                       // 4 FFMA2 (same flops as 8 FFMA) + 8 integer ops
#pragma unroll
                for (int i = 0; i < 4; ++i) y[i] = fma2(y[i], a2, c2);
#pragma unroll
                for (int i = 0; i < 8; ++i) z[i & 3] = (z[i & 3] ^ (z[(i + 1) & 3] + i)) + it;
            }
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float((unsigned)(y[i] & 0xffffffffu)) + __uint_as_float((unsigned)(y[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + z[0] + z[1] + z[2] + z[3];
}
template <int MODE> double run(float *d, int grid, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 1024>>>(d, 10); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<grid, 1024>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount * 2, iters = 4096; float *d; cudaMalloc(&d, sizeof(float) * grid * 1024);
    const double thr = (double)grid * 1024 * iters * 8;     // per-thread inner bodies
    double m0 = run<0>(d, grid, iters), m1 = run<1>(d, grid, iters), m2 = run<2>(d, grid, iters), m3 = run<3>(d, grid, iters);
    printf("FFMA  x8          : %.3f ms  %.1f TFLOP/s\n", m0, thr * 8 * 2 / m0 / 1e9);
    printf("FFMA2 x8          : %.3f ms  %.1f TFLOP/s\n", m1, thr * 16 * 2 / m1 / 1e9);
    printf("FFMA x8 + int x8  : %.3f ms  %.1f TFLOP/s\n", m2, thr * 8 * 2 / m2 / 1e9);
    printf("FFMA2 x4 + int x8 : %.3f ms  %.1f TFLOP/s (same flops as the line above)\n", m3, thr * 8 * 2 / m3 / 1e9);
    return 0;
}
