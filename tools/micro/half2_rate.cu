// Issue rate of packed-half (HFMA2 / HADD2) against packed-fp32 (FFMA2 / FADD2), PRMT and IDP.4A on sm_100a, and of mixes of them.
// Question (BC4/BC5 lattice path): would the slot computation in half2 (3 ops per pixel pair) beat fp32x2 (2 ops that take 2 dispatch cycles each)?
// 16 independent chains per thread, 8 x unrolled.  Prints warp-instructions per clock per SM (4 = one per scheduler per clock).
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define ITER 2048
__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) { uint32_t d; asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) { unsigned long long d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { uint32_t d; asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s)); return d; }
template <int MODE> __global__ void k(uint32_t* out, uint32_t ua, uint32_t ub, int n) {
    uint32_t s[16]; unsigned long long w[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = 0x3C003C00u + threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = 0x3F8000003F800000ull + threadIdx.x + i;
    const uint32_t ra = ua + (threadIdx.x & 1), rb = ub + (threadIdx.x & 2);
    const unsigned long long wa = 0x3F8000013F800001ull + (threadIdx.x & 1), wb = 0x3A0000003A000000ull + (threadIdx.x & 2);
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (MODE == 0) s[i] = hfma2(s[i], ra, rb);
                if (MODE == 1) s[i] = hadd2(s[i], rb);
                if (MODE == 2) s[i] = prmt(s[i], ra, rb);
                if (MODE == 3) s[i] = __dp4a(s[i], ra, rb);
                if (MODE == 4) { if (i < 8) w[i] = ffma2(w[i], wa, wb); }                     // 8 FFMA2 per pass
                if (MODE == 5) { if (i < 8) w[i] = fadd2(w[i], wb); }
                if (MODE == 6) s[i] = (i & 1) ? hfma2(s[i], ra, rb) : prmt(s[i], ra, rb);     // HFMA2 / PRMT alternating
                if (MODE == 7) { if (i & 1) s[i] = prmt(s[i], ra, rb); else w[i >> 1] = ffma2(w[i >> 1], wa, wb); }   // FFMA2 / PRMT alternating
                if (MODE == 8) s[i] = (i & 1) ? hfma2(s[i], ra, rb) : hadd2(s[i], rb);
                if (MODE == 9) s[i] = (i % 3 == 0) ? prmt(s[i], ra, rb) : hfma2(s[i], ra, rb);   // 2 HFMA2 : 1 PRMT
                if (MODE == 10) s[i] = s[i] * 17u + ra;                                        // IMAD r, imm, r
                if (MODE == 11) s[i] = (i & 1) ? (s[i] * 17u + ra) : prmt(s[i], ra, rb);
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc ^= s[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, uint32_t* d, int block, int ctas, double per_pass) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * ctas;
    k<MODE><<<grid, block>>>(d, 0x3C013C01u, 0x10001000u, 16);
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(d, 0x3C013C01u, 0x10001000u, ITER);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double winstr = (double)grid * (block / 32) * ITER * 8.0 * per_pass;
    printf("%-34s %4d thr x %d CTA/SM %8.3f ms  %5.2f warp-instr per clk per SM\n", name, block, ctas, ms, winstr / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    uint32_t* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
    for (int cfg = 0; cfg < 2; ++cfg) {
        const int block = cfg ? 512 : 256, ctas = cfg ? 1 : 8;      // 64 / 16 warps per SM
        run<0>("HFMA2 r,r,r", d, block, ctas, 16); run<1>("HADD2 r,r", d, block, ctas, 16); run<2>("PRMT r,r,r", d, block, ctas, 16);
        run<3>("IDP.4A r,r,r", d, block, ctas, 16); run<4>("FFMA2", d, block, ctas, 8); run<5>("FADD2", d, block, ctas, 8);
        run<6>("HFMA2 / PRMT 1:1", d, block, ctas, 16); run<7>("FFMA2 / PRMT 1:1", d, block, ctas, 16); run<8>("HFMA2 / HADD2 1:1", d, block, ctas, 16);
        run<9>("HFMA2 / PRMT 2:1", d, block, ctas, 16); run<10>("IMAD r,imm,r", d, block, ctas, 16); run<11>("IMAD / PRMT 1:1", d, block, ctas, 16);
    }
    cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
