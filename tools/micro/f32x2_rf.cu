// Does FFMA2 with three distinct per-thread register pairs run slower than with uniform operands? (RF bank rule)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi){ asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
#define ITER 4096
template <int MODE> __global__ void k(float* out, const float* in, float a, float b, int n) {
    u64 r[8];
    const float t = in[threadIdx.x];            // per-thread values -> vector registers
    const u64 x = pk(t, t + 1.0f), y = pk(t * 0.5f, t * 0.25f), x2 = pk(t + 2.0f, t + 3.0f), y2 = pk(t * 0.125f, t * 0.0625f);
    const u64 ua = pk(a, a), ub = pk(b, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = pk(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(ua), "l"(ub));                 // R, UR, UR
            if (MODE == 1) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(x), "l"(ub));                  // R, R, UR
            if (MODE == 2) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"((i & 1) ? x : x2), "l"((i & 1) ? y : y2));   // R, R, R
            if (MODE == 3) asm("add.rn.f32x2 %0, %0, %1;" : "+l"(r[i]) : "l"((i & 1) ? x : x2));                // R, R
            if (MODE == 4) asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(r[i]) : "l"((i & 1) ? x : x2));                // R, R
        }
    }
    float acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float p, q; upk(r[i], p, q); acc += p + q; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, float* d, float* in) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256;
    k<MODE><<<grid, block>>>(d, in, 1.0001f, 0.0001f, 16);
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(d, in, 1.0001f, 0.0001f, ITER);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double laneops = (double)grid * block * ITER * 16.0;
    printf("%-26s %8.3f ms  %6.1f lane-results per clk per SM\n", name, ms, laneops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    float *d, *in; cudaMalloc(&d, 148 * 8 * 256 * 4); cudaMalloc(&in, 1024 * 4); cudaMemset(in, 0, 4096);
    run<0>("FFMA2 R,UR,UR", d, in); run<1>("FFMA2 R,R,UR", d, in); run<2>("FFMA2 R,R,R", d, in); run<3>("FADD2 R,R", d, in); run<4>("FMUL2 R,R", d, in);
    cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
