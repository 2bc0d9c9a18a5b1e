// Per-opcode issue rates, second batch (asm volatile so nothing is folded).  nvcc -arch=sm_100a -O3 -o alu_ops2 alu_ops2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 4096
#define OP1(name, asmstr) __device__ __forceinline__ uint32_t name(uint32_t x, uint32_t a, uint32_t b) { uint32_t d; asm volatile(asmstr : "=r"(d) : "r"(x), "r"(a), "r"(b)); return d; }
OP1(op_lop3, "lop3.b32 %0,%1,%2,%3,0x96;")
OP1(op_shf, "shf.r.wrap.b32 %0,%1,%2,5;")
OP1(op_shr, "shr.u32 %0,%1,5;")
OP1(op_iadd3, "{.reg .u32 t; add.u32 t,%1,%2; add.u32 %0,t,%3;}")
OP1(op_add, "add.u32 %0,%1,%2;")
OP1(op_lea, "{.reg .u32 t; shl.b32 t,%1,4; add.u32 %0,t,%2;}")
OP1(op_imad, "mad.lo.u32 %0,%1,%2,%3;")
OP1(op_min, "min.u32 %0,%1,%2;")
OP1(op_fmin, "min.f32 %0,%1,%2;")  // operands are b32 regs holding floats
OP1(op_sel, "{.reg .pred p; setp.gt.u32 p,%2,%3; selp.u32 %0,%1,%2,p;}")

OP1(op_and, "and.b32 %0,%1,%2;")
OP1(op_popc, "popc.b32 %0,%1;")
OP1(op_fadd, "add.rn.f32 %0,%1,%2;")
OP1(op_ffma, "fma.rn.f32 %0,%1,%2,%3;")
OP1(op_i2f, "cvt.rn.f32.u32 %0,%1;")
OP1(op_dp2a, "dp2a.lo.u32.u32 %0,%1,%2,%3;")
OP1(op_vmin2, "vmin2.u32.u32.u32 %0,%1,%2,%3;")
OP1(op_vabsdiff, "vabsdiff.u32.u32.u32 %0,%1,%2;")
template <int MODE> __global__ void k(uint32_t* out, const uint32_t* in, int n) {
    uint32_t r[16]; const uint32_t a = in[threadIdx.x], b = in[threadIdx.x + 32];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 3 + i;
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) r[i] = op_lop3(r[i], a, b);
            if (MODE == 1) r[i] = op_shf(r[i], a, b);
            if (MODE == 2) r[i] = op_shr(r[i], a, b);
            if (MODE == 3) r[i] = op_iadd3(r[i], a, b);
            if (MODE == 4) r[i] = op_add(r[i], a, b);
            if (MODE == 5) r[i] = op_lea(r[i], a, b);
            if (MODE == 6) r[i] = op_imad(r[i], a, b);
            if (MODE == 7) r[i] = op_min(r[i], a, b);
            if (MODE == 8) r[i] = op_fmin(r[i], a, b);
            if (MODE == 9) r[i] = op_sel(r[i], a, b);
            if (MODE == 10) r[i] = op_and(r[i], a, b);
            if (MODE == 11) r[i] = op_popc(r[i], a, b);
            if (MODE == 12) r[i] = op_fadd(r[i], a, b);
            if (MODE == 13) r[i] = op_ffma(r[i], a, b);
            if (MODE == 14) r[i] = op_i2f(r[i], a, b);
            if (MODE == 15) r[i] = op_dp2a(r[i], a, b);
            if (MODE == 16) r[i] = op_vmin2(r[i], a, b);
            if (MODE == 17) r[i] = op_vabsdiff(r[i], a, b);
            if (MODE == 18) { if (i & 1) r[i] = op_lop3(r[i], a, b); else r[i] = op_fadd(r[i], a, b); }     // LOP3 + FADD
            if (MODE == 19) { if (i & 1) r[i] = op_lop3(r[i], a, b); else r[i] = op_imad(r[i], a, b); }     // LOP3 + IMAD
            if (MODE == 20) { if (i & 1) r[i] = op_lop3(r[i], a, b); else r[i] = op_min(r[i], a, b); }      // LOP3 + VIMNMX
            if (MODE == 21) { if (i & 1) r[i] = op_shf(r[i], a, b); else r[i] = op_lea(r[i], a, b); }       // SHF + LEA
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, uint32_t* d, uint32_t* in) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256;
    k<MODE><<<grid, block>>>(d, in, 16);
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(d, in, ITER);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)grid * block * ITER * 16;
    printf("%-14s %8.3f ms  %6.1f source-ops per clk per SM (at 1.965 GHz)\n", name, ms, ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    uint32_t *d, *in; cudaMalloc(&d, 148 * 8 * 256 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 1, 4096);
    run<0>("LOP3", d, in); run<1>("SHF", d, in); run<2>("SHR", d, in); run<3>("IADD3", d, in); run<4>("IADD", d, in); run<5>("LEA(shl+add)", d, in);
    run<6>("IMAD", d, in); run<7>("VIMNMX", d, in); run<8>("FMNMX", d, in); run<9>("ISETP+SEL", d, in); run<10>("AND", d, in); run<11>("POPC", d, in);
    run<12>("FADD", d, in); run<13>("FFMA", d, in); run<14>("I2F.U32", d, in); run<15>("DP2A", d, in); run<16>("vmin2", d, in); run<17>("vabsdiff", d, in);
    run<18>("LOP3+FADD", d, in); run<19>("LOP3+IMAD", d, in); run<20>("LOP3+VIMNMX", d, in); run<21>("SHF+LEA", d, in);
    cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
