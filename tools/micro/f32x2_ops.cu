// Per-op throughput of FMUL2 / FADD2 / FFMA2 vs scalar FMUL / FADD / FFMA on sm_100a (independent chains).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi){ asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
#define ITER 4096
template <int MODE> __global__ void k(float* out, float a, float b, u64 nz, int n) {
    u64 r[8]; float s[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = pk(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = threadIdx.x * 0.001f + i;
    const u64 pa = pk(a, a), pb = pk(b, b);
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(r[i]) : "l"(pa));
            if (MODE == 1) asm("add.rn.f32x2 %0, %0, %1;" : "+l"(r[i]) : "l"(pb));
            if (MODE == 2) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(pa), "l"(nz));
            if (MODE == 3) asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(pa), "l"(pb));
        }
        if (MODE >= 4) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MODE == 4) s[i] = __fmul_rn(s[i], a);
                if (MODE == 5) s[i] = __fadd_rn(s[i], b);
                if (MODE == 6) s[i] = __fmaf_rn(s[i], a, b);
            }
        }
    }
    float acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x, y; upk(r[i], x, y); acc += x + y; }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, float* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256;
    k<MODE><<<grid, block>>>(d, 1.0001f, 0.0001f, 0x8000000080000000ull, 16);
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(d, 1.0001f, 0.0001f, 0x8000000080000000ull, ITER);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double laneops = (double)grid * block * ITER * 16.0;
    printf("%-22s %8.3f ms  %6.1f lane-results per clk per SM (1.965 GHz)\n", name, ms, laneops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("FMUL2", d); run<1>("FADD2", d); run<2>("FFMA2 (c = -0 pair)", d); run<3>("FFMA2", d);
    run<4>("FMUL", d); run<5>("FADD", d); run<6>("FFMA", d);
    cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
