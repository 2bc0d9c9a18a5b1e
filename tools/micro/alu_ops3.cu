// Per-opcode issue rates, third batch: the exact operand forms the lattice kernel uses.  nvcc -arch=sm_100a -O3 -o alu_ops3 alu_ops3.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 4096
#define OP1(name, asmstr) __device__ __forceinline__ uint32_t name(uint32_t x, uint32_t a, uint32_t b) { uint32_t d; asm volatile(asmstr : "=r"(d) : "r"(x), "r"(a), "r"(b)); return d; }
OP1(prmt_rr_imm, "prmt.b32 %0,%1,%2,0x5140;")                 // PRMT R, R, imm, R
OP1(prmt_r_imm_imm, "prmt.b32 %0,%1,0x47400000,0x3240;")      // one register source (magic must be a reg -> check SASS)
OP1(prmt_rrr, "prmt.b32 %0,%1,%2,%3;")
OP1(lop3_rr_imm, "lop3.b32 %0,%1,0xFF00,%2,0xEA;")            // (x & 0xFF00) | a : 2 regs + imm
OP1(lop3_r_imm_imm, "{.reg .u32 t; and.b32 t,%1,0xFF00; or.b32 %0,t,0x47400000;}")   // 1 reg + 2 imm -> one LOP3?
OP1(lop3_rrr, "lop3.b32 %0,%1,%2,%3,0x96;")
OP1(imad_r_imm_r, "mad.lo.u32 %0,%1,16,%2;")
OP1(imad_rrr, "mad.lo.u32 %0,%1,%2,%3;")
OP1(shf_1, "shr.u32 %0,%1,1;")
OP1(add_rr, "add.u32 %0,%1,%2;")
OP1(isetp_sel, "{.reg .pred p; setp.gt.u32 p,%1,%2; selp.u32 %0,%3,%1,p;}")
OP1(vimnmx3, "{.reg .u32 t; min.u32 t,%1,%2; min.u32 %0,t,%3;}")
OP1(vabsdiff4, "vabsdiff4.u32.u32.u32 %0,%1,%2,%3;")
OP1(dp4a_rrr, "dp4a.u32.u32 %0,%1,%2,%3;")
OP1(dp4a_rr0, "dp4a.u32.u32 %0,%1,%2,0;")
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
template <int MODE> __global__ void k(uint32_t* out, const uint32_t* in, int n) {
    uint32_t r[16]; const uint32_t a = in[threadIdx.x], b = in[threadIdx.x + 32];
    u64 p[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 3 + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = ((u64)__float_as_uint(1.0f + i) << 32) | __float_as_uint(2.0f + threadIdx.x);
    const u64 pa = ((u64)a << 32) | a, pb = ((u64)b << 32) | b;
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) r[i] = prmt_rr_imm(r[i], a, b);
            if (MODE == 1) r[i] = prmt_r_imm_imm(r[i], a, b);
            if (MODE == 2) r[i] = prmt_rrr(r[i], a, b);
            if (MODE == 3) r[i] = lop3_rr_imm(r[i], a, b);
            if (MODE == 4) r[i] = lop3_r_imm_imm(r[i], a, b);
            if (MODE == 5) r[i] = lop3_rrr(r[i], a, b);
            if (MODE == 6) r[i] = imad_r_imm_r(r[i], a, b);
            if (MODE == 7) r[i] = imad_rrr(r[i], a, b);
            if (MODE == 8) r[i] = add_rr(shf_1(r[i], a, b), a, b);      // SHF(1 reg) + IADD
            if (MODE == 9) r[i] = isetp_sel(r[i], a, b);
            if (MODE == 10) r[i] = vimnmx3(r[i], a, b);
            if (MODE == 11) r[i] = vabsdiff4(r[i], a, b);
            if (MODE == 12) r[i] = dp4a_rrr(r[i], a, b);
            if (MODE == 13) r[i] = dp4a_rr0(r[i], a, b);
            // mixes with packed FFMA2 (does a 2-cycle ALU op hide under a 2-cycle FFMA2?)
            if (MODE == 14) { if (i < 8) p[i] = fma2(p[i], pa, pb); else r[i] = prmt_rr_imm(r[i], a, b); }
            if (MODE == 15) { if (i < 8) p[i] = fma2(p[i], pa, pb); else r[i] = add_rr(r[i], a, b); }
            if (MODE == 16) { if (i < 8) r[i] = dp4a_rrr(r[i], a, b); else r[i] = prmt_rr_imm(r[i], a, b); }
            if (MODE == 17) { if (i < 8) r[i] = imad_r_imm_r(r[i], a, b); else r[i] = prmt_rr_imm(r[i], a, b); }
            if (MODE == 18) { if (i < 8) r[i] = imad_r_imm_r(r[i], a, b); else r[i] = add_rr(r[i], a, b); }
            if (MODE == 19) { if (i < 5) p[i] = fma2(p[i], pa, pb); else if (i < 10) r[i] = prmt_rr_imm(r[i], a, b); else r[i] = add_rr(r[i], a, b); }  // 5 FFMA2 + 5 PRMT + 6 IADD
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (uint32_t)p[i] + (uint32_t)(p[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, uint32_t* d, uint32_t* in, int per_iter = 16) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256;
    k<MODE><<<grid, block>>>(d, in, 16);
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(d, in, ITER);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)grid * block * ITER * per_iter;
    printf("%-22s %8.3f ms  %6.1f source-ops per clk per SM   (%.2f clk per warp-op-group per scheduler)\n", name, ms, ops / (ms * 1e-3) / 148 / 1.965e9,
           128.0 / (ops / (ms * 1e-3) / 148 / 1.965e9));
}
int main() {
    uint32_t *d, *in; cudaMalloc(&d, 148 * 8 * 256 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 1, 4096);
    run<0>("PRMT r,r,imm", d, in); run<1>("PRMT r,imm,imm", d, in); run<2>("PRMT r,r,r", d, in); run<3>("LOP3 r,imm,r", d, in); run<4>("LOP3 r,imm,imm", d, in);
    run<5>("LOP3 r,r,r", d, in); run<6>("IMAD r,imm,r", d, in); run<7>("IMAD r,r,r", d, in); run<8>("SHF+IADD (pair)", d, in); run<9>("ISETP+SEL (pair)", d, in);
    run<10>("VIMNMX3", d, in); run<11>("VABSDIFF4", d, in); run<12>("IDP4A r,r,r", d, in); run<13>("IDP4A r,r,0", d, in);
    run<14>("8 FFMA2 + 8 PRMT", d, in); run<15>("8 FFMA2 + 8 IADD", d, in); run<16>("8 IDP4A + 8 PRMT", d, in); run<17>("8 IMAD + 8 PRMT", d, in);
    run<18>("8 IMAD + 8 IADD", d, in); run<19>("5 FFMA2+5 PRMT+6 IADD", d, in);
    cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
