// Issue rate of scalar FMUL / FADD / FFMA forms on sm_100a (16 independent chains per thread, 8 x unrolled: loop overhead < 2 %).
// Question: do FMUL / FADD issue at the FFMA rate?  Would fma(a, b, -0) / fma(a, 1, b) be faster spellings of the same roundings?
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
template <int MODE> __global__ void k(float* out, float a, float b, float nz, float one, int n) {
    float s[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = threadIdx.x * 0.001f + i;
    float ra = a + threadIdx.x * 1e-9f, rb = b + threadIdx.x * 1e-9f;      // per-thread register operands
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MODE == 0) s[i] = __fmul_rn(s[i], ra);                   // FMUL r, r
                if (MODE == 1) s[i] = __fadd_rn(s[i], rb);                   // FADD r, r
                if (MODE == 2) s[i] = __fmaf_rn(s[i], ra, rb);               // FFMA r, r, r
                if (MODE == 3) s[i] = __fmaf_rn(s[i], ra, nz);               // FFMA as rounded product (c = -0 in a register)
                if (MODE == 4) s[i] = __fmaf_rn(s[i], one, rb);              // FFMA as rounded sum (b = 1 in a register)
                if (MODE == 5) s[i] = __fmul_rn(s[i], 1.0001f);              // FMUL r, imm
                if (MODE == 6) s[i] = __fadd_rn(s[i], 0.0001f);              // FADD r, imm
                if (MODE == 7) s[i] = (i & 1) ? __fmul_rn(s[i], ra) : __fadd_rn(s[i], rb);   // FMUL / FADD mix
                if (MODE == 8) s[i] = (i & 1) ? __fmul_rn(s[i], ra) : __fmaf_rn(s[i], ra, rb);   // FMUL / FFMA mix
                if (MODE == 9) s[i] = __fmul_rn(s[i], s[(i + 1) & 15]);      // FMUL with two varying registers
            }
        }
    }
    float acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, float* d, int block, int ctas) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * ctas;
    k<MODE><<<grid, block>>>(d, 1.0001f, 0.0001f, -0.0f, 1.0f, 16);
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(d, 1.0001f, 0.0001f, -0.0f, 1.0f, ITER);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)grid * block * ITER * 128.0;
    printf("%-44s %4d thr x %d CTA/SM %8.3f ms  %6.1f lane-results per clk per SM (1.965 GHz)\n", name, block, ctas, ms, ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
    for (int cfg = 0; cfg < 2; ++cfg) {
        const int block = cfg ? 128 : 256, ctas = cfg ? 6 : 8;      // 64 / 24 warps per SM
        run<0>("FMUL r,r", d, block, ctas); run<1>("FADD r,r", d, block, ctas); run<2>("FFMA r,r,r", d, block, ctas);
        run<3>("FFMA r,r,(-0)  == rounded product", d, block, ctas); run<4>("FFMA r,(1),r   == rounded sum", d, block, ctas);
        run<5>("FMUL r,imm", d, block, ctas); run<6>("FADD r,imm", d, block, ctas); run<7>("FMUL/FADD alternating", d, block, ctas);
        run<8>("FMUL/FFMA alternating", d, block, ctas); run<9>("FMUL r,r' (two varying)", d, block, ctas);
    }
    cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
