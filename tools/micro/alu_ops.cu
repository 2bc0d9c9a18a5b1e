// Issue rates of the byte-permute / byte-SIMD / packed-fp32 ops the BC4/BC5 lattice kernel is made of, alone and
// mixed (independent chains, per-thread operands).  nvcc -arch=sm_100a -O3 -o alu_ops alu_ops.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 4096
typedef unsigned long long u64;
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { uint32_t d; asm volatile("prmt.b32 %0,%1,%2,%3;" : "=r"(d) : "r"(a), "r"(b), "r"(s)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0,%1,%2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
template <int MODE> __global__ void k(uint32_t* out, const uint32_t* in, int n) {
    uint32_t r[16]; const uint32_t a = in[threadIdx.x], b = in[threadIdx.x + 32], c = in[threadIdx.x + 64];
    u64 p[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 3 + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = ((u64)__float_as_uint(1.0f + i) << 32) | __float_as_uint(2.0f + threadIdx.x);
    const u64 pa = ((u64)a << 32) | a, pb = ((u64)b << 32) | b;
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) r[i] = prmt(r[i], a, b);                                      // PRMT reg selector
            if (MODE == 1) r[i] = __vabsdiffu4(r[i], a);                                 // VABSDIFF4
            if (MODE == 2) r[i] = __dp4a(r[i], a, r[i]);                                 // IDP.4A
            if (MODE == 3) r[i] = (r[i] >> 5) | b;                                       // SHF + LOP (2 instr)
            if (MODE == 4) r[i] = r[i] * 16u + a;                                        // shift-add (IMAD or LEA)
            if (MODE == 5) r[i] = r[i] > a ? b : c;                                      // ISETP + SEL
            if (MODE == 6) r[i] = __vimin3_u32(r[i], a, b);                              // VIMNMX3
            if (MODE == 7) r[i] = min(r[i], a + i);                                      // VIMNMX
            if (MODE == 8 && i < 8) p[i] = fma2(p[i], pa, pb);                           // FFMA2
            if (MODE == 9 && i < 8) p[i] = add2(p[i], pa);                               // FADD2
            if (MODE == 10) { if (i < 8) p[i] = fma2(p[i], pa, pb); else r[i] = prmt(r[i], a, b); }          // FFMA2 + PRMT 1:1
            if (MODE == 11) { if (i < 8) p[i] = fma2(p[i], pa, pb); else r[i] = __vimin3_u32(r[i], a, b); }  // FFMA2 + VIMNMX3 1:1
            if (MODE == 12) { if (i < 8) r[i] = r[i] * 16u + a; else r[i] = prmt(r[i], a, b); }              // IMAD + PRMT 1:1
            if (MODE == 13) { if (i < 8) r[i] = __vimin3_u32(r[i], a, b); else r[i] = prmt(r[i], a, b); }    // VIMNMX3 + PRMT 1:1
            if (MODE == 14) { if (i < 8) r[i] = __dp4a(r[i], a, r[i]); else r[i] = prmt(r[i], a, b); }       // IDP4A + PRMT 1:1
            if (MODE == 15) { if (i < 8) r[i] = __vabsdiffu4(r[i], a); else r[i] = prmt(r[i], a, b); }       // VABSDIFF4 + PRMT 1:1
            if (MODE == 16) r[i] = (r[i] & a) | b;                                       // LOP3
            if (MODE == 17) r[i] = __float_as_uint(fminf(__uint_as_float(r[i]), __uint_as_float(a)));        // FMNMX
            if (MODE == 18) { if (i < 8) r[i] = __dp4a(r[i], a, r[i]); else r[i] = r[i] * 16u + a; }         // IDP4A + IMAD 1:1
            if (MODE == 19) { if (i < 8) p[i] = fma2(p[i], pa, pb); else r[i] = r[i] * 16u + a; }            // FFMA2 + IMAD 1:1
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (uint32_t)p[i] + (uint32_t)(p[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, uint32_t* d, uint32_t* in, int per_iter) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256;
    k<MODE><<<grid, block>>>(d, in, 16);
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(d, in, ITER);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)grid * block * ITER * per_iter;
    printf("%-26s %8.3f ms  %6.1f thread-instr per clk per SM (at 1.965 GHz)\n", name, ms, ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    uint32_t *d, *in; cudaMalloc(&d, 148 * 8 * 256 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 1, 4096);
    run<0>("PRMT", d, in, 16); run<1>("VABSDIFF4", d, in, 16); run<2>("IDP4A", d, in, 16); run<3>("SHF+LOP (2)", d, in, 32);
    run<4>("shift-add", d, in, 16); run<5>("ISETP+SEL (2)", d, in, 32); run<6>("VIMNMX3", d, in, 16); run<7>("VIMNMX", d, in, 16);
    run<8>("FFMA2 (issue)", d, in, 8); run<9>("FADD2 (issue)", d, in, 8); run<10>("FFMA2+PRMT", d, in, 16); run<11>("FFMA2+VIMNMX3", d, in, 16);
    run<12>("IMAD+PRMT", d, in, 16); run<13>("VIMNMX3+PRMT", d, in, 16); run<14>("IDP4A+PRMT", d, in, 16); run<15>("VABSDIFF4+PRMT", d, in, 16);
    run<16>("LOP3", d, in, 16); run<17>("FMNMX", d, in, 16); run<18>("IDP4A+IMAD", d, in, 16); run<19>("FFMA2+IMAD", d, in, 16);
    cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
