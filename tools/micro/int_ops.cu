// Issue rates of the integer / min ops the BC4/BC5 encoder is made of (independent chains, per-thread operands).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 4096
template <int MODE> __global__ void k(int* out, const int* in, int n) {
    int r[16]; const int a = in[threadIdx.x], b = in[threadIdx.x + 32], c = in[threadIdx.x + 64];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 3 + i;
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 0.5f + i;
    const float fa = __int_as_float(a) , fb = __int_as_float(b);
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) r[i] = r[i] * a + b;                                   // IMAD
            if (MODE == 1) r[i] = __vimin3_s32(r[i], a + i, b - i);               // VIMNMX3 (operands hoisted)
            if (MODE == 2) r[i] = (r[i] & a) ^ b;                                 // LOP3
            if (MODE == 3) f[i] = __fmaf_rn(f[i], fa, fb);                        // FFMA
            if (MODE == 4) f[i] = fminf(fminf(f[i], fa), fb);                     // FMNMX3
            if (MODE == 5) r[i] = min(r[i], a + i);                               // VIMNMX
            if (MODE == 6) { r[i] = r[i] * a + b; f[i] = __fmaf_rn(f[i], fa, fb); }   // IMAD + FFMA mix
            if (MODE == 7) { r[i] = r[i] * a + b; r[(i + 8) & 15] = min(r[(i + 8) & 15], c); }   // IMAD + VIMNMX mix
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i] + (int)f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int* d, int* in, int per_iter) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256;
    k<MODE><<<grid, block>>>(d, in, 16);
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(d, in, ITER);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)grid * block * ITER * per_iter;
    printf("%-22s %8.3f ms  %6.1f thread-instr per clk per SM (1.965 GHz)\n", name, ms, ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    int *d, *in; cudaMalloc(&d, 148 * 8 * 256 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 1, 4096);
    run<0>("IMAD", d, in, 16); run<1>("VIMNMX3", d, in, 16); run<2>("LOP3", d, in, 16); run<3>("FFMA", d, in, 16);
    run<4>("FMNMX3(2xFMNMX?)", d, in, 16); run<5>("VIMNMX", d, in, 16); run<6>("IMAD+FFMA", d, in, 32); run<7>("IMAD+VIMNMX", d, in, 32);
    cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
