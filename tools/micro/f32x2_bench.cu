// Microbenchmark: throughput of scalar FMUL/FADD vs packed FMUL2/FADD2 (and mixed with ALU work) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2_bench f32x2_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
__device__ __forceinline__ void add2(float& x, float& y, float a, float b) {
    asm("{.reg .b64 p,q,r; mov.b64 p,{%0,%1}; mov.b64 q,{%2,%3}; add.rn.f32x2 r,p,q; mov.b64 {%0,%1},r;}" : "+f"(x), "+f"(y) : "f"(a), "f"(b));
}
__device__ __forceinline__ void mul2(float& x, float& y, float a, float b) {
    asm("{.reg .b64 p,q,r; mov.b64 p,{%0,%1}; mov.b64 q,{%2,%3}; mul.rn.f32x2 r,p,q; mov.b64 {%0,%1},r;}" : "+f"(x), "+f"(y) : "f"(a), "f"(b));
}

template <int MODE>
__global__ void k(float* out, float a, float b, int n) {
    float r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 0.001f + i;
    uint32_t u[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3};
    for (int it = 0; it < n; ++it) {
        if (MODE == 0) {            // 16 scalar ops (8 FMUL + 8 FADD), independent chains
#pragma unroll
            for (int i = 0; i < 16; i += 2) { r[i] = __fmul_rn(r[i], a); r[i + 1] = __fadd_rn(r[i + 1], b); }
        } else if (MODE == 1) {     // 8 packed ops = 16 lane-ops
#pragma unroll
            for (int i = 0; i < 16; i += 4) { mul2(r[i], r[i + 1], a, a); add2(r[i + 2], r[i + 3], b, b); }
        } else if (MODE == 2) {     // 16 scalar FP + 8 ALU (LOP3/min)
#pragma unroll
            for (int i = 0; i < 16; i += 2) { r[i] = __fmul_rn(r[i], a); r[i + 1] = __fadd_rn(r[i + 1], b); }
#pragma unroll
            for (int i = 0; i < 4; ++i) { u[i] = min(u[i] ^ 0x5bd1e995u, u[(i + 1) & 3] + 7u); u[i] = (u[i] >> 3) | (u[i] << 29); }
        } else {                    // 8 packed FP + 8 ALU
#pragma unroll
            for (int i = 0; i < 16; i += 4) { mul2(r[i], r[i + 1], a, a); add2(r[i + 2], r[i + 3], b, b); }
#pragma unroll
            for (int i = 0; i < 4; ++i) { u[i] = min(u[i] ^ 0x5bd1e995u, u[(i + 1) & 3] + 7u); u[i] = (u[i] >> 3) | (u[i] << 29); }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + u[0] + u[1] + u[2] + u[3];
}

template <int MODE> void run(const char* name, float* d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256;
    k<MODE><<<grid, block>>>(d, 1.0001f, 0.0001f, 16);
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(d, 1.0001f, 0.0001f, ITER);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double laneops = (double)grid * block * ITER * 16.0;
    printf("%-28s %8.3f ms  %7.2f T lane-fp-ops/s  (%.1f per clk per SM at 1.965 GHz)\n", name, ms, laneops / ms / 1e9, laneops / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("scalar FMUL+FADD", d);
    run<1>("packed FMUL2+FADD2", d);
    run<2>("scalar FP + 8 ALU", d);
    run<3>("packed FP + 8 ALU", d);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
