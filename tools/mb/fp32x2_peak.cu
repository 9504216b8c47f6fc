#include <cstdio>
#include <cuda_runtime.h>
__constant__ unsigned long long c_negz2 = 0x8000000080000000ull;
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) { unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c_negz2)); return d; }
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) { unsigned long long d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ unsigned long long pk(float x, float y) { unsigned long long d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(x), "f"(y)); return d; }
__global__ void __launch_bounds__(256) k_x2(int iters, float* out)
{
    unsigned long long a[8];
    for (int k = 0; k < 8; k++) a[k] = pk(threadIdx.x * 1e-3f + k, threadIdx.x * 2e-3f + k);
    const unsigned long long m = pk(0.999999f, 0.999998f), c = pk(1e-6f, 2e-6f);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = (k & 1) ? add2(a[k], c) : mul2(a[k], m);
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = (k & 1) ? mul2(a[k], m) : add2(a[k], c);
    }
    unsigned long long s = 0; for (int k = 0; k < 8; k++) s ^= a[k];
    if (s == 123456ull) out[0] = 1.0f;
}
__global__ void __launch_bounds__(256) k_x1(int iters, float* out)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999999f, c = 1e-6f;
    for (int i = 0; i < iters; i++) {
        a0 = a0 * m; a1 = a1 + c; a2 = a2 * m; a3 = a3 + c; a4 = a4 * m; a5 = a5 + c; a6 = a6 * m; a7 = a7 + c;
        a0 = a0 + c; a1 = a1 * m; a2 = a2 + c; a3 = a3 * m; a4 = a4 + c; a5 = a5 * m; a6 = a6 + c; a7 = a7 * m;
    }
    float s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456f) out[0] = s;
}
int main()
{
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, blocks = 148 * 8;
    for (int which = 0; which < 2; which++) for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        if (which) k_x2<<<blocks, 256>>>(iters, d); else k_x1<<<blocks, 256>>>(iters, d);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flop = (double)blocks * 256.0 * iters * 16.0 * (which ? 2 : 1);
        printf("%s: %.3f ms  %.2f TFLOP/s\n", which ? "f32x2" : "scalar", ms, flop / (ms * 1e-3) / 1e12);
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
