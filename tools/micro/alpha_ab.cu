// Stand-alone A/B harness for the BC4/BC5 lattice kernel (texpresso_b200/csrc/txp_alpha_lattice.cuh): times several
// launch shapes of ONE build (variant macros are passed with -D) on a device-generated 16384^2 noise image and checks
// every output bit for bit against the literal kernel (alpha_encode_kernel).  No Python, compiles in seconds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -lineinfo [-DTXP_LAT_...] -o alpha_ab alpha_ab.cu
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cstdlib>
#include <vector>
#include "../../texpresso_b200/csrc/txp_common.cuh"
#include "../../texpresso_b200/csrc/txp_alpha.cuh"
#include "../../texpresso_b200/csrc/alpha_lattice_data.h"
#include "../../texpresso_b200/csrc/txp_alpha_lattice.cuh"
using namespace txp;

__global__ void gen(uint32_t* img, size_t n, int kind) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long z = i + 0x9E3779B97F4A7C15ull * 4;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    uint32_t v = (uint32_t)z;
    if (kind == 1) {                                        // smooth-ish: narrow ranges, no 0/255
        const uint32_t x = (uint32_t)(i & 16383), y = (uint32_t)(i >> 14);
        const uint32_t r = 40 + ((x * 3 + y) >> 7) % 150 + (v & 31), g = 30 + ((x + 2 * y) >> 6) % 180 + ((v >> 8) & 15);
        v = r | (g << 8);
    }
    img[i] = (v & 0xFFFFu) | 0xFF000000u;
}
__global__ void flushk(uint4* p, size_t n) { size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = make_uint4(1, 2, 3, 4); }
__global__ void cmp(const uint2* a, const uint2* b, size_t n, unsigned long long* bad) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (a[i].x != b[i].x || a[i].y != b[i].y)) atomicAdd(bad, 1ull);
}

static uint4* g_flush; static size_t g_flush_n = (256u << 20) / 16;
template <class F> float timeit(F f, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f, sum = 0;
    for (int i = 0; i < reps + 2; ++i) {
        flushk<<<(unsigned)((g_flush_n + 255) / 256), 256>>>(g_flush, g_flush_n);
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (i >= 2) { sum += ms; best = ms < best ? ms : best; }
    }
    return sum / reps;
}

template <int FMT, int T, int M, int S>
void run_shape(const BlockSource& src, uint8_t* out, const uint8_t* ref, int sms, const char* tag) {
    cudaFuncSetAttribute(alpha_lattice_image_kernel<FMT, T, M, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lattice_image_smem<T, S>());
    const uint32_t ntiles = (uint32_t)((src.nblocks + 31) / 32);
    const uint32_t grid = (uint32_t)sms * M;
    const uint64_t step = (uint64_t)grid * (T / 32) * 32;
    const size_t bs = FMT == BC4 ? 8 : 16;
    cudaMemset(out, 0xEE, src.nblocks * bs);
    auto f = [&] { alpha_lattice_image_kernel<FMT, T, M, S><<<grid, T, lattice_image_smem<T, S>()>>>(src, out, ntiles, (uint32_t)(step / src.bw), (uint32_t)(step % src.bw)); };
    const float ms = timeit(f, 5);
    unsigned long long* bad; cudaMallocManaged(&bad, 8); *bad = 0;
    const size_t n2 = src.nblocks * bs / 8;
    cmp<<<(unsigned)((n2 + 255) / 256), 256>>>((const uint2*)out, (const uint2*)ref, n2, bad);
    cudaDeviceSynchronize();
    const double bytes = (double)src.nblocks * (64 + bs);
    printf("%-6s %s T=%3d M=%d S=%d  %7.4f ms  %5.1f%% of 6455.6 GB/s  mismatches=%llu  %s\n", tag, FMT == BC4 ? "BC4" : "BC5", T, M, S, ms,
           100.0 * bytes / (ms * 1e-3) / 6455.6e9, *bad, cudaGetErrorString(cudaGetLastError()));
    cudaFree(bad);
}

// the TMA-staged kernel (one cp.async.bulk.tensor per 32-block strip) in the same launch shape
static bool make_map(const BlockSource& src, TmaDesc* out) {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) return false;
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap m; const cuuint64_t dims[2] = {src.w, src.h}, strides[1] = {(cuuint64_t)src.w * 4}; const cuuint32_t box[2] = {128, 4}, es[2] = {1, 1};
    if (((Fn)fp)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)src.rgba, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    memcpy(out, &m, sizeof m); return true;
}
template <int FMT, int T, int M, int S>
void run_shape_tma(const BlockSource& src, uint8_t* out, const uint8_t* ref, int sms, const char* tag) {
    cudaFuncSetAttribute(alpha_lattice_tma_kernel<FMT, T, M, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lattice_tma_smem<T, S>());
    TmaDesc tm; if (!make_map(src, &tm)) { printf("tensor map failed\n"); return; }
    const uint32_t ntiles = (uint32_t)((src.nblocks + 31) / 32), grid = (uint32_t)sms * M;
    const size_t bs = FMT == BC4 ? 8 : 16;
    cudaMemset(out, 0xEE, src.nblocks * bs);
    auto f = [&] { alpha_lattice_tma_kernel<FMT, T, M, S><<<grid, T, lattice_tma_smem<T, S>()>>>(tm, src, out, ntiles); };
    const float ms = timeit(f, 5);
    unsigned long long* bad; cudaMallocManaged(&bad, 8); *bad = 0;
    const size_t n2 = src.nblocks * bs / 8;
    cmp<<<(unsigned)((n2 + 255) / 256), 256>>>((const uint2*)out, (const uint2*)ref, n2, bad);
    cudaDeviceSynchronize();
    const double bytes = (double)src.nblocks * (64 + bs);
    printf("%-6s %s TMA T=%3d M=%d S=%d  %7.4f ms  %5.1f%% of 6455.6 GB/s  mismatches=%llu  %s\n", tag, FMT == BC4 ? "BC4" : "BC5", T, M, S, ms,
           100.0 * bytes / (ms * 1e-3) / 6455.6e9, *bad, cudaGetErrorString(cudaGetLastError()));
    cudaFree(bad);
}

template <int FMT> void run_fmt(const BlockSource& src, int sms, const char* tag) {
    const size_t bs = FMT == BC4 ? 8 : 16;
    uint8_t *out, *ref; cudaMalloc(&out, src.nblocks * bs); cudaMalloc(&ref, src.nblocks * bs);
    alpha_encode_kernel<FMT, 128, 6><<<(unsigned)((src.nblocks + 127) / 128), 128>>>(src, ref);
    cudaDeviceSynchronize();
    run_shape<FMT, 512, 1, 3>(src, out, ref, sms, tag);
    run_shape_tma<FMT, 512, 1, 3>(src, out, ref, sms, tag);
    run_shape<FMT, 512, 1, 3>(src, out, ref, sms, tag);
    run_shape_tma<FMT, 512, 1, 3>(src, out, ref, sms, tag);
#ifndef AB_ONE_SHAPE
    run_shape_tma<FMT, 512, 1, 2>(src, out, ref, sms, tag);
    run_shape_tma<FMT, 512, 1, 4>(src, out, ref, sms, tag);
    run_shape_tma<FMT, 256, 2, 3>(src, out, ref, sms, tag);
    run_shape_tma<FMT, 640, 1, 2>(src, out, ref, sms, tag);
#endif
    cudaFree(out); cudaFree(ref);
}

int main(int argc, char** argv) {
    const char* tag = argc > 1 ? argv[1] : "base";
    const int kind = argc > 2 ? atoi(argv[2]) : 0;
    const uint32_t w = 16384, h = 16384;
    uint32_t* img; cudaMalloc(&img, (size_t)w * h * 4);
    gen<<<(unsigned)(((size_t)w * h + 255) / 256), 256>>>(img, (size_t)w * h, kind);
    cudaMalloc(&g_flush, g_flush_n * 16);
    cudaMemcpyToSymbol(g_alpha_lattice, TXP_ALPHA_LATTICE, sizeof(TXP_ALPHA_LATTICE));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    BlockSource src = {};
    src.rgba = (const uint8_t*)img; src.masks = nullptr; src.w = w; src.h = h; src.bw = w / 4; src.nblocks = (uint64_t)(w / 4) * (h / 4); src.vec_ok = 1; src.nlevels = 1;
    run_fmt<BC4>(src, sms, tag);
    run_fmt<BC5>(src, sms, tag);
    return 0;
}
