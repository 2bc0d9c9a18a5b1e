// txp_api.cu -- C ABI (include/texpresso_b200.h) over the sm_100a kernels.
//
// Host side of the drop-in boundary: argument checks that mirror the reference's panics
// (lib.rs:295, :138, :324), per-device contexts (streams, pinned staging, device scratch), a chunked
// H2D -> kernel -> D2H pipeline for host buffers, block-row sharding across devices.  No CPU fallback:
// every data path ends in a kernel launch or an error code.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#endif
#include <cuda.h>                 // CUtensorMap + cuTensorMapEncodeTiled prototype (resolved at run time through cudaGetDriverEntryPoint; libcuda is not linked)

#include "../../include/texpresso_b200.h"
#include "single_lut_data.h"
#include "txp_common.cuh"
#include "txp_alpha.cuh"
#include "alpha_lattice_data.h"
#include "txp_alpha_lattice.cuh"
#include "txp_colour.cuh"
#include "txp_range.cuh"
#include "txp_cluster_setup.cuh"
#include "txp_cluster_lane.cuh"
#include "txp_decode.cuh"

namespace txp {

#ifndef TXP_LATTICE_STAGES
#define TXP_LATTICE_STAGES 3
#endif
constexpr int LATTICE_STAGES = TXP_LATTICE_STAGES;   // cp.async ring depth of alpha_lattice_image_kernel

// ---------------------------------------------------------------------------------------------------
// BC1/BC2/BC3 encoder kernel: one warp per block (see txp_colour.cuh)
// ---------------------------------------------------------------------------------------------------
#ifndef TXP_COLOUR_MIN_CTAS
#define TXP_COLOUR_MIN_CTAS 8        // 8 CTAs x 4 warps x 64 registers = 32 warps per SM
#endif
template <int FMT>
__global__ void __launch_bounds__(COLOUR_WARPS * 32, TXP_COLOUR_MIN_CTAS) colour_encode_kernel(const BlockSource src, const EncodeParams prm,
                                                                          uint8_t* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem[];
    WarpScratch* scratch = reinterpret_cast<WarpScratch*>(smem);
    const uint32_t* tab4 = g_tab4;                          // candidate lists are read through L1 (coalesced 128 B per step)
    const uint32_t* tab3 = g_tab3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t b = (uint64_t)blockIdx.x * COLOUR_WARPS + warp;
    if (b >= src.nblocks) return;

    // gather the 4x4 block: lanes 0..15 <-> pixels (lib.rs:311-330)
    uint32_t pix = 0;
    bool valid = false;
    if (lane < 16) {
        if (src.masks) {
            pix = __ldg(reinterpret_cast<const uint32_t*>(src.rgba) + b * 16 + lane);
            valid = (__ldg(src.masks + b) >> lane) & 1u;
        } else {
            const BlockPos bp = locate_block(src, (uint32_t)b);     // nblocks < 2^31 (checked by the host)
            const uint32_t sx = bp.x0 + (lane & 3), sy = bp.y0 + (lane >> 2);
            valid = sx < bp.w && sy < bp.h;
            if (valid) pix = __ldg(reinterpret_cast<const uint32_t*>(bp.base) + (size_t)sy * bp.w + sx);
        }
    }
    uint2 alpha_half = make_uint2(0u, 0u);
    if (FMT == BC2) alpha_half = warp_alpha_bc2(pix >> 24, valid, lane);          // lib.rs:198
    if (FMT == BC3) alpha_half = warp_alpha_bc3(pix >> 24, valid, lane);          // lib.rs:199
    const uint2 colour = colour_block<FMT == BC1, false>(pix, valid, prm, scratch + warp, tab3, tab4, lane);
    if (lane == 0) {
        if (FMT == BC1) reinterpret_cast<uint2*>(out)[b] = colour;
        else reinterpret_cast<uint4*>(out)[b] = make_uint4(alpha_half.x, alpha_half.y, colour.x, colour.y);   // lib.rs:213
    }
}

// ---------------------------------------------------------------------------------------------------
// ClusterFit search kernel ("K2"): one warp per block that cluster_setup_kernel flagged for the partition search.
// Writes only the colour half; the alpha half of BC2/BC3 was written by the setup kernel.
// ---------------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(COLOUR_WARPS * 32, TXP_COLOUR_MIN_CTAS) colour_search_kernel(const BlockSource src, const EncodeParams prm,
                                                                                               const uint4* __restrict__ setup,
                                                                                               uint8_t* __restrict__ out, const uint64_t first, const uint32_t n) {
    extern __shared__ __align__(16) unsigned char smem[];
    WarpScratch* scratch = reinterpret_cast<WarpScratch*>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t lb = (uint64_t)blockIdx.x * COLOUR_WARPS + warp;      // launch-local: blocks [first, first + n) of src
    if (lb >= n) return;
    const uint64_t b = first + lb;
    const uint4 su = __ldg(setup + lb);
    if (!(su.z & SETUP_SEARCH)) return;                   // finished by the setup kernel (0 or 1 points)
    uint32_t pix = 0;
    bool valid = false;
    if (lane < 16) {
        if (src.masks) {
            pix = __ldg(reinterpret_cast<const uint32_t*>(src.rgba) + b * 16 + lane);
            valid = (__ldg(src.masks + b) >> lane) & 1u;
        } else {
            const BlockPos bp = locate_block(src, (uint32_t)b);
            const uint32_t sx = bp.x0 + (lane & 3), sy = bp.y0 + (lane >> 2);
            valid = sx < bp.w && sy < bp.h;
            if (valid) pix = __ldg(reinterpret_cast<const uint32_t*>(bp.base) + (size_t)sy * bp.w + sx);
        }
    }
    const unsigned long long ow0 = (unsigned long long)su.x | ((unsigned long long)su.y << 32);
    const uint2 colour = colour_block<FMT == BC1, true>(pix, valid, prm, scratch + warp, g_tab3, g_tab4, lane, ow0,
                                                         (su.z & SETUP_DEGENERATE) != 0);
    if (lane == 0) {
        uint2* out2 = reinterpret_cast<uint2*>(out);
        if (FMT == BC1) out2[b] = colour; else out2[2 * b + 1] = colour;                // lib.rs:213
    }
}

// ---------------------------------------------------------------------------------------------------
// 2x2 box-filter mip level (harness-defined: the reference has no mip generation):
// dst(x,y) = (a + b + c + d + 2) >> 2 per channel over src(2x..2x+1, 2y..2y+1), coordinates clamped to the edge.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mip_downsample_kernel(const uint32_t* __restrict__ src, const uint32_t sw, const uint32_t sh,
                                                             uint32_t* __restrict__ dst, const uint32_t dw, const uint32_t dh) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dw * dh) return;
    const uint32_t y = i / dw, x = i - y * dw;
    const uint32_t x0 = min(2 * x, sw - 1), x1 = min(2 * x + 1, sw - 1), y0 = min(2 * y, sh - 1), y1 = min(2 * y + 1, sh - 1);
    const uint32_t a = __ldg(src + (size_t)y0 * sw + x0), b = __ldg(src + (size_t)y0 * sw + x1);
    const uint32_t c = __ldg(src + (size_t)y1 * sw + x0), d = __ldg(src + (size_t)y1 * sw + x1);
    const uint32_t m = 0x00FF00FFu;
    const uint32_t even = (((a & m) + (b & m) + (c & m) + (d & m) + 0x00020002u) >> 2) & m;
    const uint32_t odd = ((((a >> 8) & m) + ((b >> 8) & m) + ((c >> 8) & m) + ((d >> 8) & m) + 0x00020002u) >> 2) & m;
    dst[i] = even | (odd << 8);
}

// ---------------------------------------------------------------------------------------------------
// The whole mip chain of a group of textures in ONE launch (same filter as mip_downsample_kernel, level by level).
// A CTA owns a 64x64 tile of level 0 of one texture and produces the tile's share of levels 1..6 out of shared memory (a level-l
// pixel (x, y) only needs level l-1 pixels (2x..2x+1, 2y..2y+1): tiles stay self-contained); the CTA that finishes last for
// its texture (ticket counter, self-resetting) produces the remaining levels from level 6, which is at most 64x64 pixels for
// textures up to 4096^2.  One launch instead of ten per texture, for any number of textures of one shape.
// ---------------------------------------------------------------------------------------------------
struct MipTable { uint32_t lw[16], lh[16], loff[16]; int nlevels; uint32_t tex_px, tiles_x, tiles; };

__device__ __forceinline__ uint32_t mip_avg4(const uint32_t a, const uint32_t b, const uint32_t c, const uint32_t d) {
    const uint32_t m = 0x00FF00FFu;
    const uint32_t even = (((a & m) + (b & m) + (c & m) + (d & m) + 0x00020002u) >> 2) & m;
    const uint32_t odd = ((((a >> 8) & m) + ((b >> 8) & m) + ((c >> 8) & m) + ((d >> 8) & m) + 0x00020002u) >> 2) & m;
    return even | (odd << 8);
}

__global__ void __launch_bounds__(256) mip_chain_kernel(uint32_t* __restrict__ base, const __grid_constant__ MipTable mt, uint32_t* __restrict__ tickets) {
    __shared__ uint32_t sa[64 * 64], sb[32 * 32];              // level l / level l+1 of the tile; tail: level 6 of the texture (<= 64x64) / level 7
    __shared__ uint32_t s_last;
    uint32_t* tex = base + (size_t)blockIdx.y * mt.tex_px;
    const uint32_t tx = blockIdx.x % mt.tiles_x, ty = blockIdx.x / mt.tiles_x;
    const int tid = threadIdx.x;
    // ---- levels 1 .. min(6, nlevels - 1): from global (level 0) resp. shared memory
    const int in_tile = mt.nlevels - 1 < 6 ? mt.nlevels - 1 : 6;
    uint32_t* src_s = sa; uint32_t* dst_s = sb;
    for (int l = 1; l <= in_tile; ++l) {
        const uint32_t sw = mt.lw[l - 1], sh = mt.lh[l - 1], dw = mt.lw[l], dh = mt.lh[l];
        const uint32_t ox = (tx * 64) >> l, oy = (ty * 64) >> l, side = 64u >> l;          // this tile's region of level l
        const uint32_t sox = (tx * 64) >> (l - 1), soy = (ty * 64) >> (l - 1), sside = 64u >> (l - 1);
        uint32_t* dst_g = tex + mt.loff[l];
        const uint32_t* src_g = tex + mt.loff[l - 1];
        for (uint32_t i = tid; i < side * side; i += 256) {
            const uint32_t lx = i % side, ly = i / side, x = ox + lx, y = oy + ly;
            if (x < dw && y < dh) {
                const uint32_t x0 = min(2 * x, sw - 1), x1 = min(2 * x + 1, sw - 1), y0 = min(2 * y, sh - 1), y1 = min(2 * y + 1, sh - 1);
                uint32_t a, b, c, d;
                if (l == 1) {
                    a = __ldg(src_g + (size_t)y0 * sw + x0); b = __ldg(src_g + (size_t)y0 * sw + x1);
                    c = __ldg(src_g + (size_t)y1 * sw + x0); d = __ldg(src_g + (size_t)y1 * sw + x1);
                } else {
                    a = src_s[(y0 - soy) * sside + (x0 - sox)]; b = src_s[(y0 - soy) * sside + (x1 - sox)];
                    c = src_s[(y1 - soy) * sside + (x0 - sox)]; d = src_s[(y1 - soy) * sside + (x1 - sox)];
                }
                const uint32_t v = mip_avg4(a, b, c, d);
                dst_s[ly * side + lx] = v;
                dst_g[(size_t)y * dw + x] = v;
            }
        }
        __syncthreads();
        uint32_t* t = src_s; src_s = dst_s; dst_s = t;
    }
    if (mt.nlevels - 1 <= 6) return;
    // ---- the remaining levels, by the CTA that finishes last for this texture
    __threadfence();
    if (tid == 0) {
        const uint32_t ticket = atomicAdd(tickets + blockIdx.y, 1u);
        s_last = ticket == mt.tiles - 1;
        if (s_last) tickets[blockIdx.y] = 0;                   // ready for the next launch on this slot
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {
        const uint32_t w6 = mt.lw[6], h6 = mt.lh[6];
        const uint32_t* g6 = tex + mt.loff[6];
        for (uint32_t i = tid; i < w6 * h6; i += 256) sa[i] = __ldcg(g6 + i);      // written by other CTAs: read through L2
        __syncthreads();
    }
    src_s = sa; dst_s = sb;
    for (int l = 7; l < mt.nlevels; ++l) {
        const uint32_t sw = mt.lw[l - 1], sh = mt.lh[l - 1], dw = mt.lw[l], dh = mt.lh[l];
        uint32_t* dst_g = tex + mt.loff[l];
        for (uint32_t i = tid; i < dw * dh; i += 256) {
            const uint32_t x = i % dw, y = i / dw;
            const uint32_t x0 = min(2 * x, sw - 1), x1 = min(2 * x + 1, sw - 1), y0 = min(2 * y, sh - 1), y1 = min(2 * y + 1, sh - 1);
            const uint32_t v = mip_avg4(src_s[y0 * sw + x0], src_s[y0 * sw + x1], src_s[y1 * sw + x0], src_s[y1 * sw + x1]);
            dst_s[i] = v;
            dst_g[i] = v;
        }
        __syncthreads();
        uint32_t* t = src_s; src_s = dst_s; dst_s = t;
    }
}

// ---------------------------------------------------------------------------------------------------
// Channel expansion to RGBA8 on the device (the reference's CLI does this on the host before Format::compress,
// cli/src/image/png.rs:47-62, jpeg.rs:42-52): L8 -> (l, l, l, 255), LA8 -> (l, l, l, a), RGB8 -> (r, g, b, 255).
// One thread per pixel; a warp reads 32..96 contiguous bytes and writes 128.
// ---------------------------------------------------------------------------------------------------
// RG8 -> (r, g, 0, 255) has no counterpart in the reference (PNG has no two-channel colour type): it is the layout of
// two-channel normal maps, what BC5 encodes (lib.rs:200-203 reads R and G only).
template <int LAYOUT>
__global__ void __launch_bounds__(256) expand_pixels_kernel(const uint8_t* __restrict__ src, uint32_t* __restrict__ dst, const size_t npix) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    constexpr int BPP = LAYOUT == TXP_PIXELS_RG8 ? 2 : LAYOUT;
    const uint8_t* p = src + i * BPP;
    uint32_t v;
    if (LAYOUT == TXP_PIXELS_L8) { const uint32_t l = __ldg(p); v = l * 0x00010101u | 0xFF000000u; }
    else if (LAYOUT == TXP_PIXELS_LA8) { const uint32_t l = __ldg(p), a = __ldg(p + 1); v = l * 0x00010101u | (a << 24); }
    else if (LAYOUT == TXP_PIXELS_RG8) { v = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | 0xFF000000u; }
    else { v = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16) | 0xFF000000u; }
    dst[i] = v;
}

// ---------------------------------------------------------------------------------------------------
// Measurement helper (txp_measure_fp32_issue): 16 independent chains of rounded fp32 products per thread, 8 x unrolled.
// With FMA contraction forbidden every fp32 operation of the ClusterFit search is one lane-instruction, so the rate of
// this loop is the roof the search kernels are quoted against (bench.py roofline.peak_measured).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 6) fp32_issue_kernel(float* __restrict__ out, const float a, const int iters) {
    float s[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = threadIdx.x * 0.001f + i;
    const float ra = a + threadIdx.x * 1e-9f;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 16; ++i) s[i] = __fmul_rn(s[i], ra);
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc = __fadd_rn(acc, s[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---------------------------------------------------------------------------------------------------
// host runtime
// ---------------------------------------------------------------------------------------------------
static thread_local std::string t_last_error;
static std::atomic<uint64_t> g_launches{0};
static std::atomic<uint64_t> g_path_lane{0}, g_path_lane_iter{0}, g_path_warp{0}, g_path_hybrid{0};   // ClusterFit launches by kernel structure (txp_debug_get)
// ClusterFit kernel structure (tuning knob, txp_debug_set(0, v) or TXP_COLOUR_VARIANT=auto|fused|warp|lane):
//  0 auto  : setup kernel + lane-per-block search (txp_cluster_lane.cuh) for launches of at least g_lane_min_blocks
//            blocks, setup kernel + warp-per-block search otherwise (few blocks: the warp kernel has 32x the parallelism)
//  1 fused : the original single warp-per-block kernel          2 warp : always setup + warp-per-block search
//  3 lane  : setup + lane-per-block search for every launch
//  4 hybrid: full lane rounds + warp-per-block tail from one round upwards (A/B)
static std::atomic<int> g_colour_variant{[] {
    const char* v = getenv("TXP_COLOUR_VARIANT");
    const std::string s = v ? v : "";
    return s == "fused" ? 1 : s == "warp" ? 2 : s == "lane" ? 3 : 0;
}()};
static std::atomic<long long> g_lane_min_blocks{[] { const char* v = getenv("TXP_LANE_MIN_BLOCKS"); return v ? atoll(v) : 262144ll; }()};
// hybrid launches (launch_encode): a last lane round less than this many percent full takes the warp-per-block search; 0 = off
// Measured (profiles/size_sweep_r02.jsonl): NOT a win.  A partly filled last round is not a whole round of time -- with fewer
// warps per SM each lane gets a larger share of the FMA pipe, the lane kernel's time is linear in the block count above one
// round -- so the default is 0 (off); variant 4 / key 2 keep the structure measurable.
#ifndef TXP_TAIL_FRAC
#define TXP_TAIL_FRAC 0
#endif
static std::atomic<int> g_chunk_mib{[] { const char* v = getenv("TXP_CHUNK_MIB"); return v ? atoi(v) : 0; }()};   // > 0: pipeline chunk size override (A/B)
static std::atomic<int> g_plan_growth{[] { const char* v = getenv("TXP_PLAN_GROWTH"); return v ? atoi(v) : 3; }()};   // geometric chunk plan of small ClusterFit shards; <= 1: off
// round-aligned chunk plan of small ClusterFit shards (pipeline_plan): smallest shard, in rounds of the lane-per-block search, that takes it; 0 = off
static std::atomic<int> g_wave_plan{[] { const char* v = getenv("TXP_WAVE_PLAN"); return v ? atoi(v) : 4; }()};
static std::atomic<int> g_wave_plan_max{[] { const char* v = getenv("TXP_WAVE_PLAN_MAX"); return v ? atoi(v) : 40; }()};   // ... and the largest (<= 120 rounds: 64 chunks)
static std::atomic<int> g_wave_chunk{[] { const char* v = getenv("TXP_WAVE_CHUNK"); return v ? atoi(v) : 2; }()};   // rounds per lane chunk of that plan
static std::atomic<int> g_hybrid_tail{[] { const char* v = getenv("TXP_HYBRID_TAIL"); return v ? atoi(v) : TXP_TAIL_FRAC; }()};
constexpr uint64_t LANE_CHUNK_BLOCKS = 4u << 20;    // blocks per setup/search launch pair of the lane path (292 B of scratch per block)

static int fail(int code, const std::string& msg) { t_last_error = msg; return code; }

// ---- staging copies of pageable caller buffers ---------------------------------------------------------------------------------------
// A caller who hands over ordinary (pageable) memory -- the reference's &[u8] slices (lib.rs:287-294) -- gets every pipeline chunk copied through
// pinned staging buffers.  One thread moves ~11 GB/s, a fifth of what the PCIe link takes, which made such calls 2-9x slower than calls on pinned
// buffers (profiles/pageable_r02.jsonl: BC4 8192^2 23.3 ms against 5.0 ms).  A few helper threads split every large copy.  TXP_COPY_THREADS = threads
// per copy including the caller (default 8, capped at half the hardware threads; <= 1: plain memcpy).  The pool is created on first use and never
// joined (helpers sleep on a condition variable; a static destructor racing with the CUDA runtime's own teardown would be worse than the leak).
// One part of a staging copy.  Large parts use non-temporal stores (SSE2 MOVNTDQ): the destination is either a pinned buffer the DMA engine reads
// next or a caller buffer nobody reads soon, so write-allocating it in the cache only adds a read of every destination line (TXP_COPY_NT=0: memcpy).
static void copy_part(uint8_t* d, const uint8_t* s, size_t n) {
#if defined(__x86_64__) || defined(_M_X64)
    static const bool nt = [] { const char* v = getenv("TXP_COPY_NT"); return !v || atoi(v) != 0; }();
    if (nt && n >= (256u << 10)) {
        const size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
        std::memcpy(d, s, head);
        d += head; s += head; n -= head;
        const size_t lines = n / 64;
        for (size_t i = 0; i < lines; ++i, d += 64, s += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s)), b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 32)), e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 48));
            _mm_stream_si128(reinterpret_cast<__m128i*>(d), a); _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32), c); _mm_stream_si128(reinterpret_cast<__m128i*>(d + 48), e);
        }
        _mm_sfence();
        std::memcpy(d, s, n - lines * 64);
        return;
    }
#endif
    std::memcpy(d, s, n);
}

class CopyPool {
public:
    static CopyPool& get() { static CopyPool* p = new CopyPool(); return *p; }
    void copy(void* dst, const void* src, size_t n) {
        constexpr size_t MIN_PART = 512u << 10;
        const size_t parts = std::min<size_t>((size_t)threads_, n / MIN_PART);
        if (parts <= 1) { copy_part(static_cast<uint8_t*>(dst), static_cast<const uint8_t*>(src), n); return; }
        const size_t step = ((n + parts - 1) / parts + 4095) & ~size_t(4095);
        std::atomic<int> left{0};
        uint8_t* d = static_cast<uint8_t*>(dst);
        const uint8_t* s = static_cast<const uint8_t*>(src);
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (size_t off = step; off < n; off += step) { queue_.push_back({d + off, s + off, std::min(step, n - off), &left}); left.fetch_add(1, std::memory_order_relaxed); }
        }
        cv_.notify_all();
        copy_part(d, s, std::min(step, n));
        // help with the queue (possibly another caller's parts) until this copy's parts are done, then wait for the ones in flight
        for (;;) {
            Part p;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (left.load(std::memory_order_acquire) == 0 || queue_.empty()) break;
                p = queue_.front(); queue_.pop_front();
            }
            run(p);
        }
        std::unique_lock<std::mutex> lk(done_mu_);
        done_cv_.wait(lk, [&] { return left.load(std::memory_order_acquire) == 0; });
    }
private:
    struct Part { uint8_t* dst; const uint8_t* src; size_t n; std::atomic<int>* left; };
    CopyPool() {
        const char* v = getenv("TXP_COPY_THREADS");
        int want = v ? atoi(v) : 8;
        const int hw = (int)std::thread::hardware_concurrency();
        if (hw > 0 && want > hw / 2) want = hw / 2;
        threads_ = want < 1 ? 1 : want;
        for (int i = 1; i < threads_; ++i) std::thread([this] { worker(); }).detach();
    }
    void run(const Part& p) {
        copy_part(p.dst, p.src, p.n);
        if (p.left->fetch_sub(1, std::memory_order_acq_rel) == 1) { std::lock_guard<std::mutex> lk(done_mu_); done_cv_.notify_all(); }
    }
    void worker() {
        for (;;) {
            Part p;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return !queue_.empty(); });
                p = queue_.front(); queue_.pop_front();
            }
            run(p);
        }
    }
    int threads_ = 1;
    std::mutex mu_, done_mu_;
    std::condition_variable cv_, done_cv_;
    std::deque<Part> queue_;
};
static inline void host_copy(void* dst, const void* src, size_t n) { CopyPool::get().copy(dst, src, n); }

#define TXP_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            return fail(TXP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
        }                                                                                       \
    } while (0)

constexpr int MAX_DEVICES = 64;
#ifndef TXP_HOST_CONCURRENT
#define TXP_HOST_CONCURRENT 0      // 1: chunks of one image count as concurrent launches (lane kernels from 32768 blocks)
#endif
#ifndef TXP_SMALL_FIRST_CHUNK
#define TXP_SMALL_FIRST_CHUNK 1
#endif
#ifndef TXP_NSLOTS
#define TXP_NSLOTS 6                // batches of 1024^2 textures: 3 -> 6 slots + lane kernels = +41 % textures/s (profiles/README.md)
#endif
constexpr int NSLOTS = TXP_NSLOTS;   // pipeline slots (stream + staging) per device
constexpr size_t CHUNK_BYTES = 32u << 20;       // largest input chunk per pipeline stage
constexpr size_t MIN_CHUNK_BYTES = 2u << 20;    // smallest chunk worth a separate launch + copy

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    uint8_t *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr, *d_raw = nullptr;
    size_t h_in_cap = 0, h_out_cap = 0, d_in_cap = 0, d_out_cap = 0, d_raw_cap = 0;   // d_raw: 1-3 byte pixels before expansion
    // deferred copy of a staged result into a pageable caller buffer
    uint8_t* user_out = nullptr;
    size_t user_out_bytes = 0;
    // the same for texture groups (one entry per texture: destination, offset inside h_out, bytes)
    struct Deferred { uint8_t* dst; size_t off, n; };
    std::vector<Deferred> deferred;
    uint32_t* d_tickets = nullptr;      // mip_chain_kernel: one self-resetting counter per texture of a group
    bool busy = false;
};
constexpr int GROUP_MAX = 16;           // textures per group launch

struct DeviceCtx {
    std::mutex mu;
    bool ready = false;
    Slot slots[NSLOTS];
    uint32_t* d_masks = nullptr;
    size_t masks_cap = 0;
    // stream-ordered scratch (setup records between the ClusterFit kernels).  A private pool with an unlimited
    // release threshold: the default pool hands memory back to the OS at every synchronisation, which would put a
    // fresh device allocation inside every timed call.
    cudaMemPool_t pool = nullptr;
    int sm_count = 0;
};

static DeviceCtx g_ctx[MAX_DEVICES];

static void build_tables(std::vector<uint32_t>& t4, std::vector<uint32_t>& t3) {
    // 4-colour: reference loop nest cluster.rs:309-318 flattened k-major so that the candidates valid for
    // `count` points are a prefix; entry = i<<18 | (17i+j)<<9 | (17j+k) (also the loop-order tie key).
    t4.clear(); t3.clear();
    for (uint32_t k = 1; k <= 16; ++k)
        for (uint32_t j = 0; j <= k; ++j)
            for (uint32_t i = 0; i <= j && i <= 15; ++i)
                t4.push_back((i << 18) | ((17 * i + j) << 9) | (17 * j + k));
    // 3-colour: cluster.rs:182-187, j-major
    for (uint32_t j = 1; j <= 16; ++j)
        for (uint32_t i = 0; i <= j && i <= 15; ++i)
            t3.push_back((i << 9) | (17 * i + j));
}

static int ensure_ctx(int dev, DeviceCtx** out) {
    if (dev < 0 || dev >= MAX_DEVICES) return fail(TXP_ERR_CUDA, "device index out of range");
    DeviceCtx& c = g_ctx[dev];
    std::lock_guard<std::mutex> lk(c.mu);
    if (!c.ready) {
        std::vector<uint32_t> t4, t3;
        build_tables(t4, t3);
        if (t4.size() != TAB4_N || t3.size() != TAB3_N) return fail(TXP_ERR_CUDA, "internal: candidate table size");
        t4.resize(TAB4_PAD, 0xFFFFFFFFu); t3.resize(TAB3_PAD, 0xFFFFFFFFu);
        static const uint8_t lut[TXP_SINGLE_LUT_BYTES] = TXP_SINGLE_LUT_INIT;
        TXP_CUDA(cudaMemcpyToSymbol(g_tab4, t4.data(), TAB4_PAD * 4));
        TXP_CUDA(cudaMemcpyToSymbol(g_tab3, t3.data(), TAB3_PAD * 4));
        TXP_CUDA(cudaMemcpyToSymbol(c_single_lut, lut, sizeof lut));
        TXP_CUDA(cudaMemcpyToSymbol(g_single_lut, lut, sizeof lut));
        static_assert(sizeof(TXP_ALPHA_LATTICE) == 512 * sizeof(uint4), "alpha lattice table size");
        TXP_CUDA(cudaMemcpyToSymbol(g_alpha_lattice, TXP_ALPHA_LATTICE, sizeof(TXP_ALPHA_LATTICE)));
        TXP_CUDA(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, dev));
#define TXP_LATTICE_ATTR(F, T, M)                                                                                                                                      \
        TXP_CUDA(cudaFuncSetAttribute(alpha_lattice_image_kernel<F, T, M, LATTICE_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lattice_image_smem<T, LATTICE_STAGES>())); \
        TXP_CUDA(cudaFuncSetAttribute(alpha_lattice_tma_kernel<F, T, M, LATTICE_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lattice_tma_smem<T, LATTICE_STAGES>())); \
        TXP_CUDA(cudaFuncSetAttribute(alpha_lattice_kernel<F, T, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lattice_smem<T>()))
        TXP_LATTICE_ATTR(BC4, 256, 2); TXP_LATTICE_ATTR(BC4, 512, 1);
        TXP_LATTICE_ATTR(BC5, 256, 2); TXP_LATTICE_ATTR(BC5, 512, 1);
#undef TXP_LATTICE_ATTR
        TXP_CUDA(cudaFuncSetAttribute(colour_encode_kernel<BC1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COLOUR_SMEM));
        TXP_CUDA(cudaFuncSetAttribute(colour_encode_kernel<BC2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COLOUR_SMEM));
        TXP_CUDA(cudaFuncSetAttribute(colour_encode_kernel<BC3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COLOUR_SMEM));
        TXP_CUDA(cudaFuncSetAttribute(colour_search_kernel<BC1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COLOUR_SMEM));
        TXP_CUDA(cudaFuncSetAttribute(colour_search_kernel<BC2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COLOUR_SMEM));
        TXP_CUDA(cudaFuncSetAttribute(colour_search_kernel<BC3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)COLOUR_SMEM));
        {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            TXP_CUDA(cudaMemPoolCreate(&c.pool, &props));
            uint64_t keep = ~0ull;
            TXP_CUDA(cudaMemPoolSetAttribute(c.pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        for (Slot& s : c.slots) {
            TXP_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
            TXP_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        }
        c.ready = true;
    }
    *out = &c;
    return TXP_OK;
}

static int current_ctx(DeviceCtx** out, int* dev_out = nullptr) {
    int dev = 0;
    TXP_CUDA(cudaGetDevice(&dev));
    if (dev_out) *dev_out = dev;
    return ensure_ctx(dev, out);
}

static int grow_dev(uint8_t** p, size_t* cap, size_t need) {
    if (*cap >= need) return TXP_OK;
    if (*p) TXP_CUDA(cudaFree(*p));
    *p = nullptr; *cap = 0;
    need = (need + 255) & ~size_t(255);
    TXP_CUDA(cudaMalloc(reinterpret_cast<void**>(p), need));
    *cap = need;
    return TXP_OK;
}

static int grow_pinned(uint8_t** p, size_t* cap, size_t need) {
    if (*cap >= need) return TXP_OK;
    if (*p) TXP_CUDA(cudaFreeHost(*p));
    *p = nullptr; *cap = 0;
    TXP_CUDA(cudaHostAlloc(reinterpret_cast<void**>(p), need, cudaHostAllocDefault));
    *cap = need;
    return TXP_OK;
}

// true if the driver can DMA from/to this pointer directly (pinned / registered host, device, managed)
static bool dma_direct(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type != cudaMemoryTypeUnregistered;
}

static bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice;
}

static int check_params(int format, const txp_params* p) {
    if (format < 0 || format > 4) return fail(TXP_ERR_FORMAT, "format must be 0..4 (Bc1..Bc5)");
    if (!p) return fail(TXP_ERR_ARGUMENT, "params is null");
    if (p->algorithm > 2) return fail(TXP_ERR_FORMAT, "algorithm must be 0..2");
    return TXP_OK;
}

static EncodeParams to_device_params(const txp_params* p) {
    EncodeParams e;
    e.algorithm = (int)p->algorithm;
    e.wx = p->weights[0]; e.wy = p->weights[1]; e.wz = p->weights[2];
    e.alpha_weighted = p->weigh_colour_by_alpha ? 1 : 0;
    e.negzero2 = 0x8000000080000000ull;
    const float g[2] = {31.0f, 63.0f}, gr[2] = {1.0f / 31.0f, 1.0f / 63.0f};
    memcpy(&e.grid_xy, g, 8);
    memcpy(&e.gridrcp_xy, gr, 8);
    return e;
}

// blocks one round of the lane-per-block search holds: every SM runs TXP_LANE_MIN_CTAS CTAs of LANE_THREADS lanes, one block per lane
static uint64_t lane_wave_blocks(const DeviceCtx& ctx) { return (uint64_t)ctx.sm_count * TXP_LANE_MIN_CTAS * LANE_THREADS; }

// ClusterFit / IterativeClusterFit on blocks [first, first + n) of src, lane-per-block search:
// K1 (thread per block; also emits the points of every colour set and a window-sorted permutation) -> K2L (lane per block)
static int launch_cluster_lane(DeviceCtx& ctx, int format, const BlockSource& src, const EncodeParams& e, uint8_t* d_out, cudaStream_t st,
                               const uint64_t first_block, const uint64_t nblocks) {
    // at most LANE_CHUNK_BLOCKS blocks per launch pair (100-104 B of scratch per block), split evenly so that no launch is small
    const uint64_t n_chunks = (nblocks + LANE_CHUNK_BLOCKS - 1) / LANE_CHUNK_BLOCKS;
    const uint64_t chunk_blocks = ((nblocks + n_chunks - 1) / n_chunks + SETUP_WINDOW - 1) / SETUP_WINDOW * SETUP_WINDOW;
    for (uint64_t off = 0; off < nblocks; off += chunk_blocks) {
        const uint64_t first = first_block + off;
        const uint32_t n = (uint32_t)std::min<uint64_t>(chunk_blocks, nblocks - off);
        const size_t n32 = ((size_t)n + 31) & ~size_t(31);
        const size_t rec_bytes = n32 * (size_t)rec_quads(true, e.alpha_weighted != 0) * sizeof(uint4), word_bytes = n32 * sizeof(uint32_t);
        uint8_t* scratch = nullptr;
        TXP_CUDA(cudaMallocFromPoolAsync(reinterpret_cast<void**>(&scratch), rec_bytes + 2 * word_bytes + 256, ctx.pool, st));
        uint4* rec = reinterpret_cast<uint4*>(scratch);
        uint32_t* perm = reinterpret_cast<uint32_t*>(scratch + rec_bytes);
        uint32_t* carry = perm + n32;
        uint32_t* counters = carry + n32;
        const unsigned g1 = (unsigned)((n + SETUP_WINDOW - 1) / SETUP_WINDOW), g2 = (unsigned)((n + LANE_THREADS - 1) / LANE_THREADS);
        if (format == BC1) cluster_setup_sorted_kernel<BC1><<<g1, ROLL_THREADS, 0, st>>>(src, e, d_out, rec, perm, first, n);
        else if (format == BC2) cluster_setup_sorted_kernel<BC2><<<g1, ROLL_THREADS, 0, st>>>(src, e, d_out, rec, perm, first, n);
        else cluster_setup_sorted_kernel<BC3><<<g1, ROLL_THREADS, 0, st>>>(src, e, d_out, rec, perm, first, n);
        cudaError_t aux_err = cudaSuccess;
        if (e.algorithm == CLUSTER_FIT) {
            if (format == BC1) cluster_lane_kernel<BC1><<<g2, LANE_THREADS, 0, st>>>(e, rec, perm, d_out, first, n);
            else if (format == BC2) cluster_lane_kernel<BC2><<<g2, LANE_THREADS, 0, st>>>(e, rec, perm, d_out, first, n);
            else cluster_lane_kernel<BC3><<<g2, LANE_THREADS, 0, st>>>(e, rec, perm, d_out, first, n);
            g_launches.fetch_add(2, std::memory_order_relaxed);
            g_path_lane.fetch_add(1, std::memory_order_relaxed);
        } else {
            g_path_lane_iter.fetch_add(1, std::memory_order_relaxed);
            // IterativeClusterFit: persistent warps draw blocks from a counter; BC1 = compress3 launch + compress4 launch
            aux_err = cudaMemsetAsync(counters, 0, 256, st);
            const unsigned cap = (unsigned)ctx.sm_count * TXP_LANE_ITER_MIN_CTAS, g3 = g2 < cap ? g2 : cap;
            if (format == BC1) {
                cluster_lane_iter_kernel<BC1, true><<<g3, LANE_THREADS, 0, st>>>(e, rec, perm, carry, d_out, counters, first, n);
                cluster_lane_iter_kernel<BC1, false><<<g3, LANE_THREADS, 0, st>>>(e, rec, perm, carry, d_out, counters + 1, first, n);
                g_launches.fetch_add(1, std::memory_order_relaxed);
            } else if (format == BC2) {
                cluster_lane_iter_kernel<BC2, false><<<g3, LANE_THREADS, 0, st>>>(e, rec, perm, carry, d_out, counters, first, n);
            } else {
                cluster_lane_iter_kernel<BC3, false><<<g3, LANE_THREADS, 0, st>>>(e, rec, perm, carry, d_out, counters, first, n);
            }
            g_launches.fetch_add(2, std::memory_order_relaxed);
        }
        const cudaError_t launch_err = cudaGetLastError();
        const cudaError_t free_err = cudaFreeAsync(scratch, st);    // stream-ordered: released after the search kernel
        if (aux_err != cudaSuccess) return fail(TXP_ERR_CUDA, std::string("ClusterFit memset: ") + cudaGetErrorString(aux_err));
        if (launch_err != cudaSuccess) return fail(TXP_ERR_CUDA, std::string("ClusterFit launch: ") + cudaGetErrorString(launch_err));
        TXP_CUDA(free_err);
    }
    return TXP_OK;
}

// ... warp-per-block search: K1 (thread per block: alpha half, colour set, principal axis, first ordering) -> K2 (warp per block)
static int launch_cluster_warp(DeviceCtx& ctx, int format, const BlockSource& src, const EncodeParams& e, uint8_t* d_out, cudaStream_t st,
                               const uint64_t first, const uint64_t nblocks) {
    const uint32_t n = (uint32_t)nblocks;
    const unsigned grid = (unsigned)((nblocks + COLOUR_WARPS - 1) / COLOUR_WARPS), threads = COLOUR_WARPS * 32;
    uint4* setup = nullptr;
    TXP_CUDA(cudaMallocFromPoolAsync(reinterpret_cast<void**>(&setup), (size_t)nblocks * sizeof(uint4), ctx.pool, st));
    const unsigned g1 = (unsigned)((nblocks + 127) / 128);
    if (format == BC1) cluster_setup_kernel<BC1><<<g1, ROLL_THREADS, 0, st>>>(src, e, d_out, setup, first, n);
    else if (format == BC2) cluster_setup_kernel<BC2><<<g1, ROLL_THREADS, 0, st>>>(src, e, d_out, setup, first, n);
    else cluster_setup_kernel<BC3><<<g1, ROLL_THREADS, 0, st>>>(src, e, d_out, setup, first, n);
    if (format == BC1) colour_search_kernel<BC1><<<grid, threads, COLOUR_SMEM, st>>>(src, e, setup, d_out, first, n);
    else if (format == BC2) colour_search_kernel<BC2><<<grid, threads, COLOUR_SMEM, st>>>(src, e, setup, d_out, first, n);
    else colour_search_kernel<BC3><<<grid, threads, COLOUR_SMEM, st>>>(src, e, setup, d_out, first, n);
    g_launches.fetch_add(2, std::memory_order_relaxed);
    g_path_warp.fetch_add(1, std::memory_order_relaxed);
    const cudaError_t launch_err = cudaGetLastError();
    const cudaError_t free_err = cudaFreeAsync(setup, st);      // stream-ordered: released after the search kernel
    if (launch_err != cudaSuccess) return fail(TXP_ERR_CUDA, std::string("ClusterFit launch: ") + cudaGetErrorString(launch_err));
    TXP_CUDA(free_err);
    return TXP_OK;
}

// 2-D tensor map over a tightly packed RGBA8 image (uint32 elements, w x h), box = one strip of 32 blocks: 128 pixels x 4 rows
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); p = nullptr; }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
static bool make_strip_tensor_map(const BlockSource& src, TmaDesc* out) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    static_assert(sizeof(CUtensorMap) == sizeof(TmaDesc), "CUtensorMap size");
    CUtensorMap m;
    const cuuint64_t dims[2] = {src.w, src.h};
    const cuuint64_t strides[1] = {(cuuint64_t)src.w * 4};
    const cuuint32_t box[2] = {128, 4}, estr[2] = {1, 1};
    const CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t*>(src.rgba), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    std::memcpy(out, &m, sizeof m);
    return true;
}

// ---- kernel launchers -------------------------------------------------------------------------------
// concurrent: the caller keeps several launches in flight on different streams (texture batches), so a launch does not
// have to fill the GPU on its own for the lane-per-block kernels to pay off
static int launch_encode(DeviceCtx& ctx, int format, const BlockSource& src, const txp_params* p, uint8_t* d_out, cudaStream_t st,
                         const bool concurrent = false) {
    if (src.nblocks == 0) return TXP_OK;
    if (src.nblocks > 0x3FFFFFFFull) return fail(TXP_ERR_DIMENSIONS, "more than 2^30-1 blocks in one launch");
    const EncodeParams e = to_device_params(p);
    if (format == BC4 || format == BC5) {
        // TXP_ALPHA_VARIANT (tuning knob): unset / 0 = lattice fast path + compacted literal path (txp_alpha_lattice.cuh);
        // 1..3 = the literal one-thread-per-block kernel (txp_alpha.cuh) in different launch shapes, kept for A/B runs.
        static const int variant = [] { const char* e = getenv("TXP_ALPHA_VARIANT"); return e ? atoi(e) : 0; }();
#define TXP_ALPHA_LAUNCH(F, T, M) alpha_encode_kernel<F, T, M><<<(unsigned)((src.nblocks + (T) - 1) / (T)), T, 0, st>>>(src, d_out)
#define TXP_LATTICE_LAUNCH(F, T, M)                                                                                   \
    do {                                                                                                              \
        const uint32_t ntiles = (uint32_t)((src.nblocks + 31) / 32);                                                  \
        const uint32_t need = (ntiles + (T) / 32 - 1) / ((T) / 32), cap = (uint32_t)ctx.sm_count * (M);               \
        const uint32_t grid = need < cap ? need : cap;                                                                \
        TmaDesc tmap;                                                                                                 \
        if (!src.masks && src.nlevels <= 1 && src.ntex <= 1 && src.vec_ok && alpha_staged && alpha_tma && (src.bw % 32) == 0 && src.h >= 4 && make_strip_tensor_map(src, &tmap)) { \
            /* strips of 32 blocks staged with one cp.async.bulk.tensor each (TMA) */                                 \
            alpha_lattice_tma_kernel<F, T, M, LATTICE_STAGES><<<grid, T, lattice_tma_smem<T, LATTICE_STAGES>(), st>>>(tmap, src, d_out, ntiles); \
        } else if (!src.masks && src.nlevels <= 1 && src.ntex <= 1 && src.vec_ok && alpha_staged && (uint64_t)src.w * src.h * 4 < 0xF0000000ull) {                                           \
            /* plain aligned image: cp.async-staged kernel; per-iteration block stride as (quotient, remainder) of bw */ \
            const uint64_t step = (uint64_t)grid * ((T) / 32) * 32;                                                   \
            alpha_lattice_image_kernel<F, T, M, LATTICE_STAGES><<<grid, T, lattice_image_smem<T, LATTICE_STAGES>(), st>>>(             \
                src, d_out, ntiles, (uint32_t)(step / src.bw), (uint32_t)(step % src.bw));                            \
        } else {                                                                                                      \
            alpha_lattice_kernel<F, T, M><<<grid, T, lattice_smem<T>(), st>>>(src, d_out, ntiles);                                    \
        }                                                                                                             \
    } while (0)
        static const bool alpha_staged = [] { const char* e = getenv("TXP_ALPHA_STAGED"); return !e || atoi(e) != 0; }();
        static const bool alpha_tma = [] { const char* e = getenv("TXP_ALPHA_TMA"); return !e || atoi(e) != 0; }();
        // launch shapes measured with tools/micro/alpha_ab.cu (profiles/README.md): one 512-thread CTA per SM (16 warps,
        // 4 per scheduler, <= 128 registers) is the fastest; warp counts that are not a multiple of 4 per SM lose 10-15 %.
        if (format == BC4) {
            switch (variant) {
            case 1: TXP_ALPHA_LAUNCH(BC4, 128, 8); break;
            case 2: TXP_ALPHA_LAUNCH(BC4, 128, 6); break;
            case 3: TXP_ALPHA_LAUNCH(BC4, 256, 3); break;
            case 6: TXP_LATTICE_LAUNCH(BC4, 256, 2); break;
            default: TXP_LATTICE_LAUNCH(BC4, 512, 1); break;
            }
        } else {
            switch (variant) {
            case 1: TXP_ALPHA_LAUNCH(BC5, 128, 8); break;
            case 2: TXP_ALPHA_LAUNCH(BC5, 128, 6); break;
            case 3: TXP_ALPHA_LAUNCH(BC5, 256, 3); break;
            case 6: TXP_LATTICE_LAUNCH(BC5, 256, 2); break;
            default: TXP_LATTICE_LAUNCH(BC5, 512, 1); break;
            }
        }
#undef TXP_LATTICE_LAUNCH
#undef TXP_ALPHA_LAUNCH
    } else if (e.algorithm == RANGE_FIT) {
        // RangeFit: one thread per block (txp_range.cuh)
        const unsigned grid = (unsigned)((src.nblocks + ROLL_THREADS - 1) / ROLL_THREADS);
        if (format == BC1) range_encode_kernel<BC1><<<grid, ROLL_THREADS, 0, st>>>(src, e, d_out);
        else if (format == BC2) range_encode_kernel<BC2><<<grid, ROLL_THREADS, 0, st>>>(src, e, d_out);
        else range_encode_kernel<BC3><<<grid, ROLL_THREADS, 0, st>>>(src, e, d_out);
    } else {
        const int variant = g_colour_variant.load(std::memory_order_relaxed);
        if (variant == 1) {                                                                        // single-kernel variant kept for A/B measurements
            const unsigned grid = (unsigned)((src.nblocks + COLOUR_WARPS - 1) / COLOUR_WARPS), threads = COLOUR_WARPS * 32;
            if (format == BC1) colour_encode_kernel<BC1><<<grid, threads, COLOUR_SMEM, st>>>(src, e, d_out);
            else if (format == BC2) colour_encode_kernel<BC2><<<grid, threads, COLOUR_SMEM, st>>>(src, e, d_out);
            else colour_encode_kernel<BC3><<<grid, threads, COLOUR_SMEM, st>>>(src, e, d_out);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            TXP_CUDA(cudaGetLastError());
            return TXP_OK;
        }
        // IterativeClusterFit: a block is 3-6 searches back to back on one lane, so a launch has to be larger still before the
        // lane structure wins (8192^2 over 8 GPUs = 524 288 blocks per rank: 9.14 ms lane, 9.07 ms warp; 54.1 vs 66.7 ms at 4 Mi blocks)
        const bool iterate = e.algorithm == ITERATIVE_CLUSTER_FIT;
        const uint64_t lane_min = (uint64_t)g_lane_min_blocks.load(std::memory_order_relaxed) * (iterate ? 3 : 1) / (concurrent ? 8 : 1);
        const uint64_t wave = lane_wave_blocks(ctx);
        uint64_t n_lane = 0;                                     // blocks [0, n_lane) -> lane-per-block search, the rest -> warp-per-block search
        if (variant == 3 || (variant == 0 && src.nblocks >= lane_min)) n_lane = src.nblocks;
        if (variant == 4 && src.nblocks >= wave) n_lane = src.nblocks;
        // Hybrid launch: one lane evaluates its block's 967 candidates back to back (0.2 ms), so a lane launch costs whole rounds of
        // `wave` blocks.  A last round that is less than TXP_TAIL_FRAC full goes to the warp-per-block search instead, which spreads
        // those blocks over the whole GPU at ~1.2-1.4x the work per block.
        if (n_lane && !iterate && !concurrent && (variant == 0 || variant == 4) && g_hybrid_tail.load(std::memory_order_relaxed)) {
            const uint64_t full = (src.nblocks / wave) * wave / SETUP_WINDOW * SETUP_WINDOW, tail = src.nblocks - full;
            if (full > 0 && tail > 0 && tail * 100 < wave * (uint64_t)g_hybrid_tail.load(std::memory_order_relaxed)) n_lane = full;
        }
        int rc;
        if (n_lane && (rc = launch_cluster_lane(ctx, format, src, e, d_out, st, 0, n_lane)) != TXP_OK) return rc;
        if (n_lane < src.nblocks && (rc = launch_cluster_warp(ctx, format, src, e, d_out, st, n_lane, src.nblocks - n_lane)) != TXP_OK) return rc;
        if (n_lane && n_lane < src.nblocks) g_path_hybrid.fetch_add(1, std::memory_order_relaxed);
        return TXP_OK;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    TXP_CUDA(cudaGetLastError());
    return TXP_OK;
}

static int launch_decode(int format, const uint8_t* d_data, uint64_t nblocks, uint32_t w, uint32_t h, uint32_t bw,
                         uint8_t* d_out, cudaStream_t st) {
    if (nblocks == 0) return TXP_OK;
    if (nblocks > 0x7FFFFFFFull) return fail(TXP_ERR_DIMENSIONS, "more than 2^31-1 blocks in one launch");
    const unsigned grid = (unsigned)((nblocks + 255) / 256);
    const int vec_ok = (w != 0 && (w % 4) == 0 && (reinterpret_cast<uintptr_t>(d_out) % 16) == 0) ? 1 : 0;
    switch (format) {
    case BC1: decode_kernel<BC1><<<grid, 256, 0, st>>>(d_data, nblocks, w, h, bw, vec_ok, d_out); break;
    case BC2: decode_kernel<BC2><<<grid, 256, 0, st>>>(d_data, nblocks, w, h, bw, vec_ok, d_out); break;
    case BC3: decode_kernel<BC3><<<grid, 256, 0, st>>>(d_data, nblocks, w, h, bw, vec_ok, d_out); break;
    case BC4: decode_kernel<BC4><<<grid, 256, 0, st>>>(d_data, nblocks, w, h, bw, vec_ok, d_out); break;
    default:  decode_kernel<BC5><<<grid, 256, 0, st>>>(d_data, nblocks, w, h, bw, vec_ok, d_out); break;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    TXP_CUDA(cudaGetLastError());
    return TXP_OK;
}

static BlockSource image_source(const uint8_t* d_rgba, size_t w, size_t h, uint64_t nblocks) {
    BlockSource s;
    s.rgba = d_rgba; s.masks = nullptr;
    s.w = (uint32_t)w; s.h = (uint32_t)h; s.bw = (uint32_t)((w + 3) / 4);
    s.nblocks = nblocks;
    s.nlevels = 1;
    s.ntex = 1; s.tex_px = 0; s.tex_blocks = 0;
    s.vec_ok = ((w % 4) == 0 && (reinterpret_cast<uintptr_t>(d_rgba) % 16) == 0) ? 1 : 0;
    return s;
}

static int check_dims(size_t w, size_t h) {
    if (w == 0) return fail(TXP_ERR_DIMENSIONS, "width must be non-zero");
    if (w > 0x3FFFFFFFull || h > 0x3FFFFFFFull) return fail(TXP_ERR_DIMENSIONS, "dimension too large");
    return TXP_OK;
}

// ---- slot helpers -------------------------------------------------------------------------------------
static int slot_wait(Slot& s) {
    if (!s.busy) return TXP_OK;
    TXP_CUDA(cudaEventSynchronize(s.done));
    if (s.user_out) { host_copy(s.user_out, s.h_out, s.user_out_bytes); s.user_out = nullptr; }
    for (const Slot::Deferred& d : s.deferred) host_copy(d.dst, s.h_out + d.off, d.n);
    s.deferred.clear();
    s.busy = false;
    return TXP_OK;
}

// CUDA call inside a pipeline loop: record the failure and leave the loop, so that the slots are always drained
#define TXP_CUDA_BREAK(expr)                                                                    \
    {                                                                                           \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) { rc = fail(TXP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); break; } \
    }

// after a failed pipeline: wait for everything in flight and forget the deferred copies into the caller's buffer
static void slots_abandon(DeviceCtx& c) {
    for (Slot& s : c.slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        s.user_out = nullptr; s.user_out_bytes = 0; s.deferred.clear(); s.busy = false;
    }
    cudaGetLastError();
}

// Rows per pipeline chunk for a shard of `rows` block rows of a w-wide image.
//  * at least ~6 chunks per call so that H2D, kernels and D2H of neighbouring chunks overlap (small shards at 8 ranks)
//  * ClusterFit searches are compute-bound (>= 1 ms per 16 MiB against 0.3 ms of H2D) and their lane-per-block kernels want
//    launches of at least lane_min blocks: a shard that holds two or more such launches is cut into chunks of exactly that many
//    blocks, rounded UP to whole block rows (measured, profiles/iter_e2e_r02.txt: 16 MiB chunks 18.56 ms, 32 MiB 18.83 ms,
//    64 MiB 19.50 ms end to end for BC3 ClusterFit 8192^2 against 18.42 ms device-resident -- the copies of the first and the
//    last chunk are the exposed ones)
//  * IterativeClusterFit: every launch ends in a drain tail of up to 8 orderings per block, twice for BC1, so few large launches:
//    chunks of ~ITER_CHUNK_BYTES (same file: 48 MiB chunks 60.6 ms, 96 MiB 57.6 ms, 192 MiB 59.2 ms against 53.7 ms; the persistent
//    search kernels of different chunks do not overlap -- a chunk's setup kernel finds no room on an SM full of search CTAs -- so every
//    further chunk adds two drain tails, about 1 ms: profiles/plan_iter_r02.txt, 8192^2 BC1: 64 + 2 x 992 block rows 57.1 ms,
//    64 + 3 x 662 rows 57.9 ms, 64 + 4 x 496 rows 59.4 ms, one chunk 59.7 ms, device-resident 54.3 ms)
constexpr size_t ITER_CHUNK_BYTES = 128u << 20;
static size_t pipeline_rows_per_chunk(int format, const txp_params* p, size_t w, size_t rows) {
    const size_t bw = (w + 3) / 4, row_bytes = 16 * w;
    size_t chunk_bytes = rows * row_bytes / 6;
    if (chunk_bytes > CHUNK_BYTES) chunk_bytes = CHUNK_BYTES;
    if (chunk_bytes < MIN_CHUNK_BYTES) chunk_bytes = MIN_CHUNK_BYTES;
    size_t rows_per_chunk = chunk_bytes / row_bytes;
    if (format <= BC3 && p->algorithm != RANGE_FIT) {
        const long long lm = g_lane_min_blocks.load(std::memory_order_relaxed);
        const size_t lane_rows = ((size_t)(lm > 0 ? lm : 1) + bw - 1) / bw;                        // ceil: the launch must not fall below the threshold
        if (p->algorithm == ITERATIVE_CLUSTER_FIT) {
            const size_t iter_rows = ITER_CHUNK_BYTES / row_bytes;
            if (iter_rows >= 3 * lane_rows && rows >= 3 * lane_rows + lane_rows / 2) rows_per_chunk = iter_rows;
        } else if (rows >= 2 * lane_rows) {
            rows_per_chunk = lane_rows;
        }
    }
    const int force_mib = g_chunk_mib.load(std::memory_order_relaxed);
    if (force_mib > 0) rows_per_chunk = ((size_t)force_mib << 20) / row_bytes;
    return rows_per_chunk ? rows_per_chunk : 1;
}

// Chunk sizes (block rows) of the H2D -> kernels -> D2H pipeline for a shard of `rows` block rows.
//  * default: uniform chunks, the first one quarter-sized (its H2D copy is the only one nothing overlaps); for ClusterFit the rows
//    after it are spread evenly, so that no short last chunk falls below the lane-per-block threshold
//  * IterativeClusterFit: the first chunk is half a lane launch (131 072 blocks, 8 MiB: its 2 ms of kernels cover the copy of a
//    96 MiB chunk), the rest is spread evenly over chunks of ~ITER_CHUNK_BYTES
//  * ClusterFit shards of fewer than three lane-sized chunks (8192^2 over 8 GPUs: 256 block rows per rank): geometric growth
//    c, G c, rest -- every chunk's kernels cover the next chunk's copy (compute : PCIe time is ~3.6 : 1) and only a small copy
//    is exposed at either end (profiles/iter_e2e_r02.txt: 2.83 ms instead of 2.86 ms for BC3, 3.12 instead of 3.25 ms for BC1)
//  * ClusterFit shards of g_wave_plan..g_wave_plan_max rounds of the lane-per-block search (8192^2 over 8 GPUs: 4.6 rounds per rank):
//    round-aligned plan.  One lane evaluates its block's 967 candidates back to back, so a lane launch costs whole rounds of
//    `wave` blocks; the shard is cut into lane chunks of exactly two rounds / one round (rounded DOWN to whole block rows; the
//    single rounds last: the copy of the last chunk's output is the one nothing overlaps) and the remainder -- the part of a round
//    that would otherwise be a partly filled last round -- goes FIRST, in two small chunks on the warp-per-block search, which is
//    the right structure for a launch that is alone on the GPU while the larger copies are still in flight.  *lane_chunks: bit i
//    set = chunk i takes the lane-per-block search although it is below the lone-launch threshold (its neighbours overlap it).
static std::vector<size_t> pipeline_plan(const DeviceCtx& ctx, int format, const txp_params* p, size_t w, size_t rows, uint64_t* lane_chunks) {
    const size_t bw = (w + 3) / 4;
    size_t rows_per_chunk = pipeline_rows_per_chunk(format, p, w, rows);
    std::vector<size_t> plan;
    *lane_chunks = 0;
    const bool cluster = format <= BC3 && p->algorithm != RANGE_FIT, forced = g_chunk_mib.load(std::memory_order_relaxed) != 0;
    const int growth = g_plan_growth.load(std::memory_order_relaxed);
    const int wave_min = g_wave_plan.load(std::memory_order_relaxed);
    if (const char* ex = getenv("TXP_PLAN")) {              // experiments only: "rows[xN],rows,...;L=<index of the first lane chunk>"
        size_t sum = 0, first_lane = 64;
        for (const char* q = ex; *q && *q != ';';) {
            char* end;
            const size_t r = strtoul(q, &end, 10);
            size_t rep = 1;
            if (*end == 'x') rep = strtoul(end + 1, &end, 10);
            for (size_t i = 0; i < rep && r > 0 && plan.size() < 4096; ++i) { plan.push_back(r); sum += r; }
            q = (*end == ',') ? end + 1 : end;
            if (end == q && *q && *q != ';') break;
        }
        if (const char* l = strstr(ex, "L=")) first_lane = strtoul(l + 2, nullptr, 10);
        if (!plan.empty() && sum >= rows) {
            for (size_t i = first_lane; i < plan.size() && i < 64; ++i) *lane_chunks |= 1ull << i;
            return plan;
        }
        plan.clear();
    }
    if (cluster && !forced && wave_min > 0 && p->algorithm == CLUSTER_FIT && g_colour_variant.load(std::memory_order_relaxed) == 0) {
        const uint64_t wave = lane_wave_blocks(ctx);
        uint64_t k = (uint64_t)rows * bw / wave;
        const size_t r1 = (size_t)(wave / bw), r2 = (size_t)(2 * wave / bw);        // block rows of a one-round / two-round chunk
        const size_t min_rows = (16384 + bw - 1) / bw;                             // a launch of at least 16 Ki blocks
        if (k >= (uint64_t)wave_min && k <= (uint64_t)g_wave_plan_max.load(std::memory_order_relaxed) && r1 >= min_rows) {
            std::vector<size_t> lane;
            size_t lane_rows = 0;
            const uint64_t cw = (uint64_t)std::max(2, g_wave_chunk.load(std::memory_order_relaxed));
            const size_t rc = (size_t)(cw * wave / bw);
            for (; k > cw && cw > 2; k -= cw) { lane.push_back(rc); lane_rows += rc; }
            for (; k > 2; k -= 2) { lane.push_back(r2); lane_rows += r2; }
            for (; k > 0; --k) { lane.push_back(r1); lane_rows += r1; }
            size_t head = rows - lane_rows;
            if (head < min_rows) { head += lane.back(); lane.pop_back(); }           // (nearly) whole rounds: the last single round opens the pipeline instead
            size_t first = head / 4 > min_rows ? head / 4 : min_rows;
            if (head < first + min_rows) first = head;
            plan.push_back(first);
            if (head > first) plan.push_back(head - first);
            for (const size_t r : lane) { *lane_chunks |= 1ull << plan.size(); plan.push_back(r); }
            return plan;
        }
    }
    if (cluster && !forced && p->algorithm == ITERATIVE_CLUSTER_FIT && rows > rows_per_chunk && rows_per_chunk * 16 * w >= ITER_CHUNK_BYTES / 2) {
        const long long lm = g_lane_min_blocks.load(std::memory_order_relaxed);
        size_t first = ((size_t)(lm > 0 ? lm : 1) / 2 + bw - 1) / bw;
        if (first < 1) first = 1;
        const size_t rest = rows - first, n = (rest + rows_per_chunk - 1) / rows_per_chunk;
        plan.push_back(first);
        plan.push_back((rest + n - 1) / n);
        return plan;
    }
    if (cluster && !forced && growth > 1 && p->algorithm == CLUSTER_FIT && rows < 3 * rows_per_chunk && rows * bw >= 131072) {
        const size_t G = (size_t)growth;
        const size_t min_rows = (16384 + bw - 1) / bw;                       // a launch of at least 16 Ki blocks
        size_t c1 = (rows + G * G + G) / (1 + G + G * G);
        if (c1 < min_rows) c1 = min_rows;
        size_t c2 = G * c1;
        if (c1 + c2 + c2 / 2 > rows) { plan.push_back(c1); plan.push_back(rows - c1); return plan; }
        plan.push_back(c1); plan.push_back(c2); plan.push_back(rows - c1 - c2);
        return plan;
    }
    const size_t first_rows = (TXP_SMALL_FIRST_CHUNK && rows >= 3 * rows_per_chunk && rows_per_chunk >= 4) ? rows_per_chunk / 4 : rows_per_chunk;
    if (cluster && rows > first_rows) {
        const size_t rest = rows - first_rows, n = rest / rows_per_chunk;
        if (n >= 1) rows_per_chunk = (rest + n - 1) / n;
    }
    plan.push_back(first_rows);
    plan.push_back(rows_per_chunk);
    return plan;
}

// Encode block rows [row0,row1) (clipped to nblocks_total) of an image held in HOST memory on the
// current device.  `out` points at the first byte of block row `row0`.
// layout: TXP_PIXELS_RGBA8 = as the reference takes it; the other layouts are expanded on the device.
static size_t layout_bpp(int layout) { return layout == TXP_PIXELS_RG8 ? 2 : (size_t)layout; }
static int compress_host_rows(DeviceCtx& c, int format, const uint8_t* rgba, size_t w, size_t h, const txp_params* p,
                              uint8_t* out, size_t row0, size_t row1, uint64_t blocks_in_range, const int layout = TXP_PIXELS_RGBA8) {
    const size_t bpp = layout_bpp(layout);
    const size_t bs = (size_t)block_bytes(format), bw = (w + 3) / 4;
    const bool in_direct = dma_direct(rgba), out_direct = dma_direct(out);
    uint64_t lane_chunks = 0;
    const std::vector<size_t> plan = pipeline_plan(c, format, p, w, row1 - row0, &lane_chunks);
    const size_t rows_per_chunk = plan.size() > 1 ? plan[1] : plan[0];
    int rc = TXP_OK;
    size_t chunk = 0;
    uint64_t blocks_left = blocks_in_range;
    size_t step = plan[0];
    for (size_t r = row0; r < row1 && blocks_left > 0 && rc == TXP_OK; r += step, ++chunk, step = std::max<size_t>(1, plan[chunk < plan.size() ? chunk : plan.size() - 1])) {
        const size_t r_end = (r + step < row1) ? r + step : row1;
        uint64_t nblk = (uint64_t)(r_end - r) * bw;
        if (nblk > blocks_left) nblk = blocks_left;
        blocks_left -= nblk;
        Slot& s = c.slots[chunk % NSLOTS];
        if ((rc = slot_wait(s)) != TXP_OK) break;
        // pixel rows of this chunk that exist in the image (rows past h are fully masked)
        const size_t y0 = 4 * r, y1 = (4 * r_end < h) ? 4 * r_end : h;
        const size_t h_sub = y1 > y0 ? y1 - y0 : 0;
        const size_t in_bytes = h_sub * w * bpp, out_bytes = (size_t)nblk * bs;
        if ((rc = grow_dev(&s.d_in, &s.d_in_cap, h_sub * w * 4 ? h_sub * w * 4 : 16)) != TXP_OK) break;
        if ((rc = grow_dev(&s.d_out, &s.d_out_cap, out_bytes)) != TXP_OK) break;
        if (bpp != 4 && (rc = grow_dev(&s.d_raw, &s.d_raw_cap, in_bytes ? in_bytes : 16)) != TXP_OK) break;
        if (in_bytes) {
            const uint8_t* src_ptr = rgba + y0 * w * bpp;
            if (!in_direct) {
                if ((rc = grow_pinned(&s.h_in, &s.h_in_cap, in_bytes)) != TXP_OK) break;
                host_copy(s.h_in, src_ptr, in_bytes);
                src_ptr = s.h_in;
            }
            TXP_CUDA_BREAK(cudaMemcpyAsync(bpp == 4 ? s.d_in : s.d_raw, src_ptr, in_bytes, cudaMemcpyDefault, s.stream));
            if (bpp != 4) {
                const size_t npix = h_sub * w;
                const unsigned g = (unsigned)((npix + 255) / 256);
                uint32_t* d32 = reinterpret_cast<uint32_t*>(s.d_in);
                if (layout == TXP_PIXELS_L8) expand_pixels_kernel<TXP_PIXELS_L8><<<g, 256, 0, s.stream>>>(s.d_raw, d32, npix);
                else if (layout == TXP_PIXELS_LA8) expand_pixels_kernel<TXP_PIXELS_LA8><<<g, 256, 0, s.stream>>>(s.d_raw, d32, npix);
                else if (layout == TXP_PIXELS_RG8) expand_pixels_kernel<TXP_PIXELS_RG8><<<g, 256, 0, s.stream>>>(s.d_raw, d32, npix);
                else expand_pixels_kernel<TXP_PIXELS_RGB8><<<g, 256, 0, s.stream>>>(s.d_raw, d32, npix);
                g_launches.fetch_add(1, std::memory_order_relaxed);
                TXP_CUDA_BREAK(cudaGetLastError());
            }
        }
        const BlockSource bsrc = image_source(s.d_in, w, h_sub, nblk);
        if ((rc = launch_encode(c, format, bsrc, p, s.d_out, s.stream,
                                (chunk < 64 && ((lane_chunks >> chunk) & 1u)) || (TXP_HOST_CONCURRENT != 0 && (row1 - row0) > 2 * rows_per_chunk))) != TXP_OK) break;
        uint8_t* dst = out + (r - row0) * bw * bs;
        if (out_direct) {
            TXP_CUDA_BREAK(cudaMemcpyAsync(dst, s.d_out, out_bytes, cudaMemcpyDefault, s.stream));
        } else {
            if ((rc = grow_pinned(&s.h_out, &s.h_out_cap, out_bytes)) != TXP_OK) break;
            TXP_CUDA_BREAK(cudaMemcpyAsync(s.h_out, s.d_out, out_bytes, cudaMemcpyDeviceToHost, s.stream));
            s.user_out = dst; s.user_out_bytes = out_bytes;
        }
        TXP_CUDA_BREAK(cudaEventRecord(s.done, s.stream));
        s.busy = true;
    }
    if (rc != TXP_OK) { const std::string keep = t_last_error; slots_abandon(c); t_last_error = keep; return rc; }
    for (Slot& s : c.slots) { const int r2 = slot_wait(s); if (rc == TXP_OK) rc = r2; }
    return rc;
}

// Decode block rows [row0,row1) of a w x h image on the current device.  `data` points at the first block of block row
// row0, `out` at pixel row 4*row0 (reference grain: one block row per rayon task, lib.rs:128-134).  Chunks of block rows are
// pipelined through the slots: the small H2D copy and the kernel of one chunk overlap the 8x larger D2H copy of the previous one.
static int decompress_host_rows(DeviceCtx& c, int format, const uint8_t* data, size_t w, size_t h, uint8_t* out, size_t row0, size_t row1) {
    const size_t bs = (size_t)block_bytes(format), bw = (w + 3) / 4;
    const bool in_direct = dma_direct(data), out_direct = dma_direct(out);
    size_t chunk_bytes = (row1 - row0) * 16 * w / 6;                 // bytes of decoded pixels per chunk
    if (chunk_bytes > CHUNK_BYTES) chunk_bytes = CHUNK_BYTES;
    if (chunk_bytes < MIN_CHUNK_BYTES) chunk_bytes = MIN_CHUNK_BYTES;
    size_t rows_per_chunk = chunk_bytes / (16 * w);
    if (rows_per_chunk == 0) rows_per_chunk = 1;
    int rc = TXP_OK;
    size_t chunk = 0;
    for (size_t r = row0; r < row1 && rc == TXP_OK; r += rows_per_chunk, ++chunk) {
        const size_t r_end = (r + rows_per_chunk < row1) ? r + rows_per_chunk : row1;
        const size_t y0 = 4 * r, y1 = (4 * r_end < h) ? 4 * r_end : h;
        if (y1 <= y0) break;
        const size_t h_sub = y1 - y0;
        const uint64_t nblk = (uint64_t)(r_end - r) * bw;
        const size_t in_bytes = (size_t)nblk * bs, out_bytes = h_sub * w * 4;
        Slot& s = c.slots[chunk % NSLOTS];
        if ((rc = slot_wait(s)) != TXP_OK) break;
        if ((rc = grow_dev(&s.d_in, &s.d_in_cap, in_bytes)) != TXP_OK) break;
        if ((rc = grow_dev(&s.d_out, &s.d_out_cap, out_bytes)) != TXP_OK) break;
        const uint8_t* src_ptr = data + (r - row0) * bw * bs;
        if (!in_direct) {
            if ((rc = grow_pinned(&s.h_in, &s.h_in_cap, in_bytes)) != TXP_OK) break;
            host_copy(s.h_in, src_ptr, in_bytes);
            src_ptr = s.h_in;
        }
        TXP_CUDA_BREAK(cudaMemcpyAsync(s.d_in, src_ptr, in_bytes, cudaMemcpyDefault, s.stream));
        if ((rc = launch_decode(format, s.d_in, nblk, (uint32_t)w, (uint32_t)h_sub, (uint32_t)bw, s.d_out, s.stream)) != TXP_OK) break;
        uint8_t* dst = out + (y0 - 4 * row0) * w * 4;
        if (out_direct) {
            TXP_CUDA_BREAK(cudaMemcpyAsync(dst, s.d_out, out_bytes, cudaMemcpyDefault, s.stream));
        } else {
            if ((rc = grow_pinned(&s.h_out, &s.h_out_cap, out_bytes)) != TXP_OK) break;
            TXP_CUDA_BREAK(cudaMemcpyAsync(s.h_out, s.d_out, out_bytes, cudaMemcpyDeviceToHost, s.stream));
            s.user_out = dst; s.user_out_bytes = out_bytes;
        }
        TXP_CUDA_BREAK(cudaEventRecord(s.done, s.stream));
        s.busy = true;
    }
    if (rc != TXP_OK) { const std::string keep = t_last_error; slots_abandon(c); t_last_error = keep; return rc; }
    for (Slot& s : c.slots) { const int r2 = slot_wait(s); if (rc == TXP_OK) rc = r2; }
    return rc;
}

static int decompress_checked(int format, const uint8_t* data, size_t data_len, size_t width, size_t height, const uint8_t* output, size_t output_len) {
    int rc;
    if (format < 0 || format > 4) return fail(TXP_ERR_FORMAT, "format must be 0..4 (Bc1..Bc5)");
    if ((rc = check_dims(width, height)) != TXP_OK) return rc;
    if (!data || !output) return fail(TXP_ERR_ARGUMENT, "null pointer");
    if (data_len < txp_compressed_size(format, width, height)) return fail(TXP_ERR_BUFFER_TOO_SMALL, "data shorter than compressed_size (reference: slice panic, lib.rs:138)");
    if (output_len < width * height * 4) return fail(TXP_ERR_BUFFER_TOO_SMALL, "output shorter than 4*width*height");
    return TXP_OK;
}

static int compress_checked(int format, const uint8_t* rgba, size_t rgba_len, size_t w, size_t h, const txp_params* p,
                            const uint8_t* output, size_t output_len) {
    int rc;
    if ((rc = check_params(format, p)) != TXP_OK) return rc;
    if ((rc = check_dims(w, h)) != TXP_OK) return rc;
    if (!rgba && w * h) return fail(TXP_ERR_ARGUMENT, "rgba is null");
    if (!output) return fail(TXP_ERR_ARGUMENT, "output is null");
    if (rgba_len < w * h * 4) return fail(TXP_ERR_BUFFER_TOO_SMALL, "rgba shorter than 4*width*height");
    const size_t need = txp_compressed_size(format, w, h);
    if (output_len < need) return fail(TXP_ERR_BUFFER_TOO_SMALL, "output shorter than compressed_size (reference asserts, lib.rs:295)");
    if (output_len % (size_t)block_bytes(format)) return fail(TXP_ERR_ARGUMENT, "output_len is not a whole number of blocks");
    return TXP_OK;
}


// Device driven by worker g of a multi-device call over n_gpus devices: devices 0 .. n_gpus - 1, except that a call with
// n_gpus == 1 stays on the calling thread's CURRENT device (one process per GPU under torchrun: every rank passes n_gpus = 1
// after txp_set_device(local_rank)).
static int worker_device(const int g, const int n_gpus, const int caller_device) { return n_gpus == 1 ? caller_device : g; }

// ---- mip chains (extension; SURVEY 8(f) row 4 / BASELINE config 5) -------------------------------------------------
static int mip_layout(int format, size_t w, size_t h, BlockSource* src, size_t* total_px, size_t* total_out) {
    int n = 0; size_t px = 0; uint64_t blocks = 0;
    size_t lw = w, lh = h;
    while (true) {
        if (n >= MAX_LEVELS) return fail(TXP_ERR_DIMENSIONS, "more than 16 mip levels");
        if (src) { src->lw[n] = (uint32_t)lw; src->lh[n] = (uint32_t)lh; src->loff[n] = (uint32_t)px; src->lfirst[n] = (uint32_t)blocks; }
        px += lw * lh; blocks += (uint64_t)txp_num_blocks(lw) * txp_num_blocks(lh);
        ++n;
        if (lw == 1 && lh == 1) break;
        lw = lw > 1 ? lw / 2 : 1; lh = lh > 1 ? lh / 2 : 1;
    }
    if (px > 0xFFFFFFFFull || blocks > 0x7FFFFFFFull) return fail(TXP_ERR_DIMENSIONS, "mip chain too large");
    if (src) { src->lfirst[n] = (uint32_t)blocks; src->nlevels = n; src->nblocks = blocks; }
    if (total_px) *total_px = px;
    if (total_out) *total_out = (size_t)blocks * (size_t)block_bytes(format);
    return n;
}

// Enqueue a GROUP of k textures of one shape on slot s: k H2D copies (level 0) -> one mip-chain launch -> ONE encode launch over
// all levels of all textures -> k D2H copies; the caller waits on the slot.  with_mips = false: level 0 only (txp_compress_batch).
// concurrent: other slots are in flight too (lets a small launch take the lane-per-block search, see launch_encode).
static int group_enqueue_impl(DeviceCtx& ctx, Slot& s, int format, const uint8_t* const* rgba, const int k, size_t w, size_t h, const txp_params* p,
                              uint8_t* const* outputs, const bool concurrent, const bool with_mips) {
    BlockSource src;
    size_t total_px = 0, total_out = 0;
    int n = 1;
    if (with_mips) {
        n = mip_layout(format, w, h, &src, &total_px, &total_out);
        if (n < 0) return n;
    } else {
        total_px = w * h;
        total_out = txp_compressed_size(format, w, h);
    }
    const size_t tex_px = (total_px + 3) & ~size_t(3);            // textures 16-byte aligned
    int rc;
    if ((rc = grow_dev(&s.d_in, &s.d_in_cap, (size_t)k * tex_px * 4)) != TXP_OK) return rc;
    if ((rc = grow_dev(&s.d_out, &s.d_out_cap, (size_t)k * total_out)) != TXP_OK) return rc;
    const size_t in_bytes = w * h * 4;
    bool staged_in = false;
    for (int t = 0; t < k; ++t) {
        const uint8_t* src_ptr = rgba[t];
        if (!dma_direct(src_ptr)) {
            if (!staged_in && (rc = grow_pinned(&s.h_in, &s.h_in_cap, (size_t)k * in_bytes)) != TXP_OK) return rc;
            staged_in = true;
            host_copy(s.h_in + (size_t)t * in_bytes, src_ptr, in_bytes);
            src_ptr = s.h_in + (size_t)t * in_bytes;
        }
        TXP_CUDA(cudaMemcpyAsync(s.d_in + (size_t)t * tex_px * 4, src_ptr, in_bytes, cudaMemcpyDefault, s.stream));
    }
    uint32_t* base = reinterpret_cast<uint32_t*>(s.d_in);
    if (n > 1) {
        if (src.lw[0] <= 4096 && src.lh[0] <= 4096) {
            if (!s.d_tickets) {
                TXP_CUDA(cudaMalloc(reinterpret_cast<void**>(&s.d_tickets), GROUP_MAX * sizeof(uint32_t)));
                TXP_CUDA(cudaMemsetAsync(s.d_tickets, 0, GROUP_MAX * sizeof(uint32_t), s.stream));
            }
            MipTable mt;
            for (int l = 0; l < 16; ++l) { mt.lw[l] = l < n ? src.lw[l] : 1; mt.lh[l] = l < n ? src.lh[l] : 1; mt.loff[l] = l < n ? src.loff[l] : 0; }
            mt.nlevels = n; mt.tex_px = (uint32_t)tex_px;
            mt.tiles_x = (uint32_t)((w + 63) / 64); mt.tiles = mt.tiles_x * (uint32_t)((h + 63) / 64);
            mip_chain_kernel<<<dim3(mt.tiles, (unsigned)k), 256, 0, s.stream>>>(base, mt, s.d_tickets);
            g_launches.fetch_add(1, std::memory_order_relaxed);
        } else {                                                      // very large textures: one launch per level
            for (int t = 0; t < k; ++t)
                for (int l = 1; l < n; ++l) {
                    const uint32_t dw = src.lw[l], dh = src.lh[l];
                    uint32_t* tb = base + (size_t)t * tex_px;
                    mip_downsample_kernel<<<(dw * dh + 255) / 256, 256, 0, s.stream>>>(tb + src.loff[l - 1], src.lw[l - 1], src.lh[l - 1], tb + src.loff[l], dw, dh);
                    g_launches.fetch_add(1, std::memory_order_relaxed);
                }
        }
        TXP_CUDA(cudaGetLastError());
    }
    uint64_t tex_blocks;
    if (with_mips) {
        src.rgba = s.d_in; src.masks = nullptr; src.w = (uint32_t)w; src.h = (uint32_t)h; src.bw = (uint32_t)txp_num_blocks(w);
        src.vec_ok = 1;                                   // cudaMalloc base; per-level width checked in locate_block
        tex_blocks = src.nblocks;
    } else {
        tex_blocks = (uint64_t)txp_num_blocks(w) * txp_num_blocks(h);
        src = image_source(s.d_in, w, h, tex_blocks);
    }
    src.ntex = (uint32_t)k; src.tex_px = (uint32_t)tex_px; src.tex_blocks = (uint32_t)tex_blocks;
    src.nblocks = tex_blocks * (uint64_t)k;
    if ((rc = launch_encode(ctx, format, src, p, s.d_out, s.stream, concurrent)) != TXP_OK) return rc;
    bool staged_out = false;
    for (int t = 0; t < k; ++t) {
        if (dma_direct(outputs[t])) {
            TXP_CUDA(cudaMemcpyAsync(outputs[t], s.d_out + (size_t)t * total_out, total_out, cudaMemcpyDefault, s.stream));
        } else {
            if (!staged_out && (rc = grow_pinned(&s.h_out, &s.h_out_cap, (size_t)k * total_out)) != TXP_OK) return rc;
            staged_out = true;
            TXP_CUDA(cudaMemcpyAsync(s.h_out + (size_t)t * total_out, s.d_out + (size_t)t * total_out, total_out, cudaMemcpyDeviceToHost, s.stream));
            s.deferred.push_back({outputs[t], (size_t)t * total_out, total_out});
        }
    }
    TXP_CUDA(cudaEventRecord(s.done, s.stream));
    s.busy = true;
    return TXP_OK;
}

// a failure after the first asynchronous operation must not leave work in flight on a slot that is not marked busy
static int group_enqueue(DeviceCtx& ctx, Slot& s, int format, const uint8_t* const* rgba, const int k, size_t w, size_t h, const txp_params* p,
                         uint8_t* const* outputs, const bool concurrent, const bool with_mips) {
    const int rc = group_enqueue_impl(ctx, s, format, rgba, k, w, h, p, outputs, concurrent, with_mips);
    if (rc != TXP_OK) {
        const std::string keep = t_last_error;
        cudaStreamSynchronize(s.stream); cudaGetLastError();
        s.user_out = nullptr; s.user_out_bytes = 0; s.deferred.clear(); s.busy = false;
        t_last_error = keep;
    }
    return rc;
}

static int mipchain_enqueue(DeviceCtx& ctx, Slot& s, int format, const uint8_t* rgba, size_t w, size_t h, const txp_params* p, uint8_t* output,
                            const bool concurrent = false, const bool with_mips = true) {
    return group_enqueue(ctx, s, format, &rgba, 1, w, h, p, &output, concurrent, with_mips);
}

// Textures of this worker (t = first, first + stride, ...) from index `from`: how many consecutive ones share the shape of
// texture `from` and fit one group (<= GROUP_MAX textures, <= GROUP_BYTES of pixels)
constexpr size_t GROUP_BYTES = 48u << 20;
static int group_extent(const size_t* widths, const size_t* heights, size_t n_textures, size_t from, size_t stride, size_t bytes_per_texture) {
    int k = 1;
    size_t bytes = bytes_per_texture;
    for (size_t t = from + stride; t < n_textures && k < GROUP_MAX; t += stride) {
        if (widths[t] != widths[from] || heights[t] != heights[from] || bytes + bytes_per_texture > GROUP_BYTES) break;
        ++k; bytes += bytes_per_texture;
    }
    return k;
}

}  // namespace txp

using namespace txp;

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

size_t txp_num_blocks(size_t size) { return (size + 3) / 4; }

size_t txp_block_size(int format) { return (format < 0 || format > 4) ? 0 : (size_t)block_bytes(format); }

size_t txp_compressed_size(int format, size_t width, size_t height) {
    return txp_num_blocks(width) * txp_num_blocks(height) * txp_block_size(format);
}

const char* txp_last_error(void) { return t_last_error.c_str(); }
uint64_t txp_kernel_launches(void) { return g_launches.load(); }
const char* txp_version(void) { return "texpresso_b200 0.1 (sm_100a)"; }

int txp_debug_set(int key, int value) {
    if (key == 0) { if (value < 0 || value > 4) return fail(TXP_ERR_ARGUMENT, "colour variant must be 0..4"); g_colour_variant.store(value); return TXP_OK; }
    if (key == 7) { if (value < 2 || value > 16) return fail(TXP_ERR_ARGUMENT, "round-aligned plan: rounds per lane chunk, 2..16"); g_wave_chunk.store(value); return TXP_OK; }
    if (key == 6) { if (value < 0 || value > 120) return fail(TXP_ERR_ARGUMENT, "round-aligned plan: largest shard in rounds, 0..120"); g_wave_plan_max.store(value); return TXP_OK; }
    if (key == 5) { if (value < 0 || value > 64) return fail(TXP_ERR_ARGUMENT, "round-aligned plan: smallest shard in rounds, 0..64 (0 = off)"); g_wave_plan.store(value); return TXP_OK; }
    if (key == 4) { if (value < 0 || value > 16) return fail(TXP_ERR_ARGUMENT, "plan growth must be 0..16"); g_plan_growth.store(value); return TXP_OK; }
    if (key == 3) { if (value < 0 || value > 4096) return fail(TXP_ERR_ARGUMENT, "chunk override must be 0..4096 MiB"); g_chunk_mib.store(value); return TXP_OK; }
    if (key == 2) { if (value < 0 || value > 100) return fail(TXP_ERR_ARGUMENT, "hybrid tail threshold must be 0..100 (percent of a lane round)"); g_hybrid_tail.store(value); return TXP_OK; }
    if (key == 1) { if (value < 1) return fail(TXP_ERR_ARGUMENT, "lane threshold must be >= 1 block"); g_lane_min_blocks.store(value); return TXP_OK; }
    return fail(TXP_ERR_ARGUMENT, "unknown debug key");
}

int txp_debug_get(int key, uint64_t* value) {
    if (!value) return fail(TXP_ERR_ARGUMENT, "null pointer");
    switch (key) {
    case 0: *value = (uint64_t)g_colour_variant.load(); return TXP_OK;
    case 1: *value = (uint64_t)g_lane_min_blocks.load(); return TXP_OK;
    case 2: *value = g_path_lane.load(); return TXP_OK;          // ClusterFit launches that took the lane-per-block search
    case 3: *value = g_path_lane_iter.load(); return TXP_OK;     // IterativeClusterFit launches that took the lane-per-block search
    case 4: *value = g_path_warp.load(); return TXP_OK;          // (Iterative)ClusterFit launches that took the warp-per-block search
    case 5: *value = g_path_hybrid.load(); return TXP_OK;        // ClusterFit launches split into full lane rounds + a warp-per-block tail
    case 6: *value = (uint64_t)g_hybrid_tail.load(); return TXP_OK;
    case 7: *value = (uint64_t)g_wave_plan.load(); return TXP_OK;
    case 8: *value = (uint64_t)g_wave_plan_max.load(); return TXP_OK;
    default: return fail(TXP_ERR_ARGUMENT, "unknown debug key");
    }
}

int txp_debug_plan(int format, const txp_params* params, size_t width, size_t rows, int sm_count, size_t* chunk_rows, size_t max_chunks, size_t* n_chunks,
                   uint64_t* lane_chunks) {
    int rc;
    if ((rc = check_params(format, params)) != TXP_OK) return rc;
    if (!chunk_rows || !n_chunks || !lane_chunks || width == 0 || sm_count < 1) return fail(TXP_ERR_ARGUMENT, "bad argument");
    static DeviceCtx fake;                                // only sm_count is read by pipeline_plan
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    fake.sm_count = sm_count;
    const std::vector<size_t> plan = pipeline_plan(fake, format, params, width, rows, lane_chunks);
    size_t n = 0, chunk = 0;
    for (size_t r = 0; r < rows; ++chunk) {               // the walk of compress_host_rows: the last entry repeats
        const size_t step = std::max<size_t>(1, plan[chunk < plan.size() ? chunk : plan.size() - 1]);
        if (n >= max_chunks) return fail(TXP_ERR_BUFFER_TOO_SMALL, "more chunks than max_chunks");
        chunk_rows[n++] = std::min(step, rows - r);
        r += step;
    }
    *n_chunks = n;
    return TXP_OK;
}

int txp_debug_host_copy(void* dst, const void* src, size_t n) {
    if (n && (!dst || !src)) return fail(TXP_ERR_ARGUMENT, "null pointer");
    host_copy(dst, src, n);
    return TXP_OK;
}

int txp_measure_fp32_issue(double* lane_ops_per_second) {
    if (!lane_ops_per_second) return fail(TXP_ERR_ARGUMENT, "null pointer");
    DeviceCtx* c;
    int rc;
    if ((rc = current_ctx(&c)) != TXP_OK) return rc;
    const int grid = c->sm_count * 6, iters = 2048;
    float* d = nullptr;
    TXP_CUDA(cudaMalloc(reinterpret_cast<void**>(&d), (size_t)grid * 128 * sizeof(float)));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t err = cudaEventCreate(&e0);
    if (err == cudaSuccess) err = cudaEventCreate(&e1);
    float best = 0.f;
    for (int rep = 0; rep < 4 && err == cudaSuccess; ++rep) {              // rep 0 = warm-up
        cudaEventRecord(e0, nullptr);
        fp32_issue_kernel<<<grid, 128>>>(d, 1.0001f, iters);
        cudaEventRecord(e1, nullptr);
        err = cudaEventSynchronize(e1);
        float ms = 0.f;
        if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms > 0.f && (best == 0.f || ms < best)) best = ms;
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(d);
    if (err != cudaSuccess || best <= 0.f) return fail(TXP_ERR_CUDA, std::string("fp32 issue measurement: ") + cudaGetErrorString(err));
    *lane_ops_per_second = (double)grid * 128.0 * iters * 128.0 / (best * 1e-3);
    return TXP_OK;
}

int txp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return fail(TXP_ERR_CUDA, "no CUDA device"); }
    return n;
}

int txp_set_device(int device) {
    TXP_CUDA(cudaSetDevice(device));
    return TXP_OK;
}

void txp_shard_rows(size_t height, int rank, int world, size_t* row_begin, size_t* row_end) {
    const size_t rows = (height + 3) / 4;
    if (world < 1) world = 1;
    if (rank < 0) rank = 0;
    if (rank >= world) rank = world - 1;
    if (row_begin) *row_begin = rows * (size_t)rank / (size_t)world;
    if (row_end) *row_end = rows * (size_t)(rank + 1) / (size_t)world;
}

int txp_compress_device(int format, const void* d_rgba, size_t width, size_t height, const txp_params* params,
                        void* d_output, size_t output_len, void* cuda_stream) {
    int rc;
    if ((rc = check_params(format, params)) != TXP_OK) return rc;
    if ((rc = check_dims(width, height)) != TXP_OK) return rc;
    if (!d_output || (!d_rgba && width * height)) return fail(TXP_ERR_ARGUMENT, "null device pointer");
    if (output_len < txp_compressed_size(format, width, height)) return fail(TXP_ERR_BUFFER_TOO_SMALL, "output shorter than compressed_size");
    if (output_len % (size_t)block_bytes(format)) return fail(TXP_ERR_ARGUMENT, "output_len is not a whole number of blocks");
    DeviceCtx* c;
    if ((rc = current_ctx(&c)) != TXP_OK) return rc;
    const BlockSource src = image_source(static_cast<const uint8_t*>(d_rgba), width, height, output_len / (size_t)block_bytes(format));
    return launch_encode(*c, format, src, params, static_cast<uint8_t*>(d_output), static_cast<cudaStream_t>(cuda_stream));
}

int txp_decompress_device(int format, const void* d_data, size_t width, size_t height, void* d_output, size_t output_len,
                          void* cuda_stream) {
    int rc;
    if (format < 0 || format > 4) return fail(TXP_ERR_FORMAT, "format must be 0..4 (Bc1..Bc5)");
    if ((rc = check_dims(width, height)) != TXP_OK) return rc;
    if (!d_data || !d_output) return fail(TXP_ERR_ARGUMENT, "null device pointer");
    if (output_len < width * height * 4) return fail(TXP_ERR_BUFFER_TOO_SMALL, "output shorter than 4*width*height");
    DeviceCtx* c;
    if ((rc = current_ctx(&c)) != TXP_OK) return rc;
    const uint64_t nblocks = (uint64_t)txp_num_blocks(width) * txp_num_blocks(height);
    return launch_decode(format, static_cast<const uint8_t*>(d_data), nblocks, (uint32_t)width, (uint32_t)height,
                         (uint32_t)txp_num_blocks(width), static_cast<uint8_t*>(d_output), static_cast<cudaStream_t>(cuda_stream));
}

int txp_compress_pixels(int format, const uint8_t* pixels, size_t pixels_len, int layout, size_t width, size_t height,
                        const txp_params* params, uint8_t* output, size_t output_len) {
    if (layout == TXP_PIXELS_RGBA8) return txp_compress(format, pixels, pixels_len, width, height, params, output, output_len);
    if (layout < TXP_PIXELS_L8 || layout > TXP_PIXELS_RG8) return fail(TXP_ERR_ARGUMENT, "layout must be 1 (L8), 2 (LA8), 3 (RGB8), 4 (RGBA8) or 5 (RG8)");
    int rc;
    if (pixels_len < width * height * layout_bpp(layout)) return fail(TXP_ERR_BUFFER_TOO_SMALL, "pixels shorter than bytes_per_pixel*width*height");
    // the remaining checks of Format::compress, on the expanded image
    if ((rc = compress_checked(format, pixels, width * height * 4, width, height, params, output, output_len)) != TXP_OK) return rc;
    if (is_device_ptr(pixels) || is_device_ptr(output)) return fail(TXP_ERR_ARGUMENT, "txp_compress_pixels takes host pointers");
    DeviceCtx* c;
    if ((rc = current_ctx(&c)) != TXP_OK) return rc;
    const size_t bs = (size_t)block_bytes(format), bw = txp_num_blocks(width);
    const uint64_t nblocks = output_len / bs;
    const size_t rows = (size_t)((nblocks + bw - 1) / bw);
    std::lock_guard<std::mutex> lk(c->mu);
    return compress_host_rows(*c, format, pixels, width, height, params, output, 0, rows, nblocks, layout);
}

int txp_compress(int format, const uint8_t* rgba, size_t rgba_len, size_t width, size_t height, const txp_params* params,
                 uint8_t* output, size_t output_len) {
    int rc;
    if ((rc = compress_checked(format, rgba, rgba_len, width, height, params, output, output_len)) != TXP_OK) return rc;
    DeviceCtx* c;
    if ((rc = current_ctx(&c)) != TXP_OK) return rc;
    if (is_device_ptr(rgba) && is_device_ptr(output)) {
        if ((rc = txp_compress_device(format, rgba, width, height, params, output, output_len, nullptr)) != TXP_OK) return rc;
        TXP_CUDA(cudaStreamSynchronize(nullptr));
        return TXP_OK;
    }
    const size_t bs = (size_t)block_bytes(format), bw = txp_num_blocks(width);
    const uint64_t nblocks = output_len / bs;
    const size_t rows = (size_t)((nblocks + bw - 1) / bw);
    std::lock_guard<std::mutex> lk(c->mu);
    return compress_host_rows(*c, format, rgba, width, height, params, output, 0, rows, nblocks);
}

int txp_decompress(int format, const uint8_t* data, size_t data_len, size_t width, size_t height, uint8_t* output,
                   size_t output_len) {
    int rc;
    if ((rc = decompress_checked(format, data, data_len, width, height, output, output_len)) != TXP_OK) return rc;
    if (width * height == 0) return TXP_OK;
    DeviceCtx* c;
    if ((rc = current_ctx(&c)) != TXP_OK) return rc;
    if (is_device_ptr(data) && is_device_ptr(output)) {
        if ((rc = txp_decompress_device(format, data, width, height, output, output_len, nullptr)) != TXP_OK) return rc;
        TXP_CUDA(cudaStreamSynchronize(nullptr));
        return TXP_OK;
    }
    std::lock_guard<std::mutex> lk(c->mu);
    return decompress_host_rows(*c, format, data, width, height, output, 0, txp_num_blocks(height));
}

int txp_compress_blocks(int format, const uint8_t* rgba_blocks, const uint32_t* masks, size_t n, const txp_params* params,
                        uint8_t* output) {
    int rc;
    if ((rc = check_params(format, params)) != TXP_OK) return rc;
    if (n == 0) return TXP_OK;
    if (!rgba_blocks || !masks || !output) return fail(TXP_ERR_ARGUMENT, "null pointer");
    DeviceCtx* c;
    if ((rc = current_ctx(&c)) != TXP_OK) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    Slot& s = c->slots[0];
    const size_t bs = (size_t)block_bytes(format);
    if ((rc = grow_dev(&s.d_in, &s.d_in_cap, n * 64)) != TXP_OK) return rc;
    if ((rc = grow_dev(&s.d_out, &s.d_out_cap, n * bs)) != TXP_OK) return rc;
    if (c->masks_cap < n) {
        if (c->d_masks) TXP_CUDA(cudaFree(c->d_masks));
        c->d_masks = nullptr; c->masks_cap = 0;
        TXP_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_masks), n * 4));
        c->masks_cap = n;
    }
    TXP_CUDA(cudaMemcpyAsync(s.d_in, rgba_blocks, n * 64, cudaMemcpyDefault, s.stream));
    TXP_CUDA(cudaMemcpyAsync(c->d_masks, masks, n * 4, cudaMemcpyDefault, s.stream));
    BlockSource src;
    src.rgba = s.d_in; src.masks = c->d_masks; src.w = 0; src.h = 0; src.bw = 1; src.nblocks = n; src.vec_ok = 1; src.nlevels = 1; src.ntex = 1; src.tex_px = 0; src.tex_blocks = 0;
    if ((rc = launch_encode(*c, format, src, params, s.d_out, s.stream)) != TXP_OK) return rc;
    TXP_CUDA(cudaMemcpyAsync(output, s.d_out, n * bs, cudaMemcpyDefault, s.stream));
    TXP_CUDA(cudaStreamSynchronize(s.stream));
    return TXP_OK;
}

int txp_decompress_blocks(int format, const uint8_t* blocks, size_t n, uint8_t* rgba_blocks) {
    int rc;
    if (format < 0 || format > 4) return fail(TXP_ERR_FORMAT, "format must be 0..4 (Bc1..Bc5)");
    if (n == 0) return TXP_OK;
    if (!blocks || !rgba_blocks) return fail(TXP_ERR_ARGUMENT, "null pointer");
    DeviceCtx* c;
    if ((rc = current_ctx(&c)) != TXP_OK) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    Slot& s = c->slots[0];
    const size_t bs = (size_t)block_bytes(format);
    if ((rc = grow_dev(&s.d_in, &s.d_in_cap, n * bs)) != TXP_OK) return rc;
    if ((rc = grow_dev(&s.d_out, &s.d_out_cap, n * 64)) != TXP_OK) return rc;
    TXP_CUDA(cudaMemcpyAsync(s.d_in, blocks, n * bs, cudaMemcpyDefault, s.stream));
    if ((rc = launch_decode(format, s.d_in, n, 0, 0, 1, s.d_out, s.stream)) != TXP_OK) return rc;
    TXP_CUDA(cudaMemcpyAsync(rgba_blocks, s.d_out, n * 64, cudaMemcpyDefault, s.stream));
    TXP_CUDA(cudaStreamSynchronize(s.stream));
    return TXP_OK;
}

int txp_compress_block_masked(int format, const uint8_t rgba[64], uint32_t mask, const txp_params* params, uint8_t* output,
                              size_t output_len) {
    if (format < 0 || format > 4) return fail(TXP_ERR_FORMAT, "format must be 0..4 (Bc1..Bc5)");
    if (output_len < (size_t)block_bytes(format)) return fail(TXP_ERR_BUFFER_TOO_SMALL, "output shorter than block_size");
    return txp_compress_blocks(format, rgba, &mask, 1, params, output);
}

int txp_decompress_block(int format, const uint8_t* block, size_t block_len, uint8_t output[64]) {
    if (format < 0 || format > 4) return fail(TXP_ERR_FORMAT, "format must be 0..4 (Bc1..Bc5)");
    if (block_len < (size_t)block_bytes(format)) return fail(TXP_ERR_BUFFER_TOO_SMALL, "block shorter than block_size");
    return txp_decompress_blocks(format, block, 1, output);
}


int txp_mip_levels(size_t width, size_t height) {
    if (width == 0 || height == 0) return fail(TXP_ERR_DIMENSIONS, "zero dimension");
    return mip_layout(0, width, height, nullptr, nullptr, nullptr);
}

size_t txp_mipchain_compressed_size(int format, size_t width, size_t height) {
    size_t total = 0;
    if (format < 0 || format > 4 || width == 0 || height == 0) return 0;
    if (mip_layout(format, width, height, nullptr, nullptr, &total) < 0) return 0;
    return total;
}

int txp_compress_mipchain(int format, const uint8_t* rgba, size_t rgba_len, size_t width, size_t height, const txp_params* params,
                          uint8_t* output, size_t output_len) {
    int rc;
    if ((rc = check_params(format, params)) != TXP_OK) return rc;
    if ((rc = check_dims(width, height)) != TXP_OK) return rc;
    if (height == 0) return fail(TXP_ERR_DIMENSIONS, "height must be non-zero");
    if (!rgba || !output) return fail(TXP_ERR_ARGUMENT, "null pointer");
    if (rgba_len < width * height * 4) return fail(TXP_ERR_BUFFER_TOO_SMALL, "rgba shorter than 4*width*height");
    if (output_len < txp_mipchain_compressed_size(format, width, height)) return fail(TXP_ERR_BUFFER_TOO_SMALL, "output shorter than the mip chain");
    DeviceCtx* c;
    if ((rc = current_ctx(&c)) != TXP_OK) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    Slot& s = c->slots[0];
    if ((rc = slot_wait(s)) != TXP_OK) return rc;
    if ((rc = mipchain_enqueue(*c, s, format, rgba, width, height, params, output)) != TXP_OK) return rc;
    return slot_wait(s);
}

int txp_compress_batch_mips(int format, const uint8_t* const* rgba, const size_t* widths, const size_t* heights, size_t n_textures,
                            const txp_params* params, uint8_t* const* outputs, int n_gpus) {
    int rc;
    if ((rc = check_params(format, params)) != TXP_OK) return rc;
    if (n_textures == 0) return TXP_OK;
    if (!rgba || !widths || !heights || !outputs) return fail(TXP_ERR_ARGUMENT, "null pointer");
    const int ndev = txp_device_count();
    if (ndev < 0) return ndev;
    int caller_device = 0;
    TXP_CUDA(cudaGetDevice(&caller_device));
    if (n_gpus < 1 || n_gpus > ndev) return fail(TXP_ERR_ARGUMENT, "n_gpus must be between 1 and the device count");
    for (size_t t = 0; t < n_textures; ++t)
        if (widths[t] == 0 || heights[t] == 0 || widths[t] > 0x3FFFFFFFull || heights[t] > 0x3FFFFFFFull)
            return fail(TXP_ERR_DIMENSIONS, "bad texture dimension");
    std::vector<int> rcs((size_t)n_gpus, TXP_OK);
    std::vector<std::string> errs((size_t)n_gpus);
    std::vector<std::thread> workers;
    for (int g = 0; g < n_gpus; ++g) {
        workers.emplace_back([&, g]() {
            int r = TXP_OK;
            DeviceCtx* c = nullptr;
            const int dev = worker_device(g, n_gpus, caller_device);
            if (cudaSetDevice(dev) != cudaSuccess) r = fail(TXP_ERR_CUDA, "cudaSetDevice failed");
            if (r == TXP_OK) r = ensure_ctx(dev, &c);
            if (r == TXP_OK) {
                std::lock_guard<std::mutex> lk(c->mu);
                size_t k = 0;                              // texture t -> device t % n_gpus, slots round-robin so that
                for (size_t t = (size_t)g; t < n_textures && r == TXP_OK; ++k) {   // copies overlap kernels
                    // consecutive textures of one shape share one mip-chain launch and one encode launch
                    const int ng = group_extent(widths, heights, n_textures, t, (size_t)n_gpus, widths[t] * heights[t] * 16 / 3 + 64);
                    const uint8_t* ins[GROUP_MAX]; uint8_t* outs[GROUP_MAX];
                    for (int i = 0; i < ng; ++i) {
                        ins[i] = rgba[t + (size_t)i * n_gpus]; outs[i] = outputs[t + (size_t)i * n_gpus];
                        if (!ins[i] || !outs[i]) r = fail(TXP_ERR_ARGUMENT, "null texture pointer");
                    }
                    if (r != TXP_OK) break;
                    Slot& s = c->slots[k % NSLOTS];
                    if ((r = slot_wait(s)) != TXP_OK) break;
                    r = group_enqueue(*c, s, format, ins, ng, widths[t], heights[t], params, outs, true, true);
                    t += (size_t)ng * n_gpus;
                }
                for (Slot& s : c->slots) { const int r2 = slot_wait(s); if (r == TXP_OK) r = r2; }
            }
            rcs[(size_t)g] = r;
            if (r != TXP_OK) errs[(size_t)g] = t_last_error;
        });
    }
    for (auto& t : workers) t.join();
    for (int g = 0; g < n_gpus; ++g)
        if (rcs[(size_t)g] != TXP_OK) return fail(rcs[(size_t)g], "gpu " + std::to_string(g) + ": " + errs[(size_t)g]);
    return TXP_OK;
}

int txp_compress_multi(int format, const uint8_t* rgba, size_t rgba_len, size_t width, size_t height, const txp_params* params,
                       uint8_t* output, size_t output_len, int n_gpus) {
    int rc;
    if ((rc = compress_checked(format, rgba, rgba_len, width, height, params, output, output_len)) != TXP_OK) return rc;
    const int ndev = txp_device_count();
    if (ndev < 0) return ndev;
    int caller_device = 0;
    TXP_CUDA(cudaGetDevice(&caller_device));
    if (n_gpus < 1 || n_gpus > ndev) return fail(TXP_ERR_ARGUMENT, "n_gpus must be between 1 and the device count");
    const size_t bs = (size_t)block_bytes(format), bw = txp_num_blocks(width);
    const uint64_t nblocks = output_len / bs;
    const size_t rows = (size_t)((nblocks + bw - 1) / bw);
    std::vector<int> rcs((size_t)n_gpus, TXP_OK);
    std::vector<std::string> errs((size_t)n_gpus);
    std::vector<std::thread> workers;
    for (int g = 0; g < n_gpus; ++g) {
        workers.emplace_back([&, g]() {
            const size_t r0 = rows * (size_t)g / (size_t)n_gpus, r1 = rows * (size_t)(g + 1) / (size_t)n_gpus;
            if (r0 >= r1) return;
            int r = TXP_OK;
            DeviceCtx* c = nullptr;
            const int dev = worker_device(g, n_gpus, caller_device);
            if (cudaSetDevice(dev) != cudaSuccess) r = fail(TXP_ERR_CUDA, "cudaSetDevice failed");
            if (r == TXP_OK) r = ensure_ctx(dev, &c);
            if (r == TXP_OK) {
                uint64_t nb = (uint64_t)(r1 - r0) * bw;
                const uint64_t first = (uint64_t)r0 * bw;
                if (first + nb > nblocks) nb = nblocks - first;
                std::lock_guard<std::mutex> lk(c->mu);
                r = compress_host_rows(*c, format, rgba, width, height, params, output + first * bs, r0, r1, nb);
            }
            rcs[(size_t)g] = r;
            if (r != TXP_OK) errs[(size_t)g] = t_last_error;
        });
    }
    for (auto& t : workers) t.join();
    for (int g = 0; g < n_gpus; ++g)
        if (rcs[(size_t)g] != TXP_OK) return fail(rcs[(size_t)g], "gpu " + std::to_string(g) + ": " + errs[(size_t)g]);
    return TXP_OK;
}

int txp_compress_batch(int format, const uint8_t* const* rgba, const size_t* widths, const size_t* heights, size_t n_textures,
                       const txp_params* params, uint8_t* const* outputs, int n_gpus) {
    int rc;
    if ((rc = check_params(format, params)) != TXP_OK) return rc;
    if (n_textures == 0) return TXP_OK;
    if (!rgba || !widths || !heights || !outputs) return fail(TXP_ERR_ARGUMENT, "null pointer");
    const int ndev = txp_device_count();
    if (ndev < 0) return ndev;
    int caller_device = 0;
    TXP_CUDA(cudaGetDevice(&caller_device));
    if (n_gpus < 1 || n_gpus > ndev) return fail(TXP_ERR_ARGUMENT, "n_gpus must be between 1 and the device count");
    std::vector<int> rcs((size_t)n_gpus, TXP_OK);
    std::vector<std::string> errs((size_t)n_gpus);
    std::vector<std::thread> workers;
    for (int g = 0; g < n_gpus; ++g) {
        workers.emplace_back([&, g]() {
            int r = TXP_OK;
            DeviceCtx* c = nullptr;
            const int dev = worker_device(g, n_gpus, caller_device);
            if (cudaSetDevice(dev) != cudaSuccess) r = fail(TXP_ERR_CUDA, "cudaSetDevice failed");
            if (r == TXP_OK) r = ensure_ctx(dev, &c);
            if (r == TXP_OK) {
                std::lock_guard<std::mutex> lk(c->mu);
                size_t k = 0;
                for (size_t t = (size_t)g; t < n_textures && r == TXP_OK;) {
                    const size_t w = widths[t], h = heights[t];
                    if ((r = check_dims(w, h)) != TXP_OK) break;
                    if (!rgba[t] || !outputs[t]) { r = fail(TXP_ERR_ARGUMENT, "null texture pointer"); break; }
                    if (h != 0 && w * h * 4 <= CHUNK_BYTES) {
                        // whole textures per pipeline slot (copies and kernels of neighbouring slots overlap); consecutive textures
                        // of one shape share one encode launch
                        const int ng = group_extent(widths, heights, n_textures, t, (size_t)n_gpus, w * h * 4);
                        const uint8_t* ins[GROUP_MAX]; uint8_t* outs[GROUP_MAX];
                        for (int i = 0; i < ng; ++i) {
                            ins[i] = rgba[t + (size_t)i * n_gpus]; outs[i] = outputs[t + (size_t)i * n_gpus];
                            if (!ins[i] || !outs[i]) r = fail(TXP_ERR_ARGUMENT, "null texture pointer");
                        }
                        if (r != TXP_OK) break;
                        Slot& s = c->slots[k++ % NSLOTS];
                        if ((r = slot_wait(s)) != TXP_OK) break;
                        r = group_enqueue(*c, s, format, ins, ng, w, h, params, outs, true, false);
                        t += (size_t)ng * n_gpus;
                    } else {
                        // large texture: the chunked pipeline of txp_compress (it drains the slots itself at the end)
                        for (Slot& s : c->slots) { const int r2 = slot_wait(s); if (r == TXP_OK) r = r2; }
                        if (r != TXP_OK) break;
                        const size_t bw = txp_num_blocks(w), rows = txp_num_blocks(h);
                        r = compress_host_rows(*c, format, rgba[t], w, h, params, outputs[t], 0, rows, (uint64_t)bw * rows);
                        t += (size_t)n_gpus;
                    }
                }
                for (Slot& s : c->slots) { const int r2 = slot_wait(s); if (r == TXP_OK) r = r2; }
            }
            rcs[(size_t)g] = r;
            if (r != TXP_OK) errs[(size_t)g] = t_last_error;
        });
    }
    for (auto& t : workers) t.join();
    for (int g = 0; g < n_gpus; ++g)
        if (rcs[(size_t)g] != TXP_OK) return fail(rcs[(size_t)g], "gpu " + std::to_string(g) + ": " + errs[(size_t)g]);
    return TXP_OK;
}

int txp_decompress_multi(int format, const uint8_t* data, size_t data_len, size_t width, size_t height, uint8_t* output,
                         size_t output_len, int n_gpus) {
    int rc;
    if ((rc = decompress_checked(format, data, data_len, width, height, output, output_len)) != TXP_OK) return rc;
    const int ndev = txp_device_count();
    if (ndev < 0) return ndev;
    int caller_device = 0;
    TXP_CUDA(cudaGetDevice(&caller_device));
    if (n_gpus < 1 || n_gpus > ndev) return fail(TXP_ERR_ARGUMENT, "n_gpus must be between 1 and the device count");
    if (width * height == 0) return TXP_OK;
    const size_t bs = (size_t)block_bytes(format), bw = txp_num_blocks(width), rows = txp_num_blocks(height);
    std::vector<int> rcs((size_t)n_gpus, TXP_OK);
    std::vector<std::string> errs((size_t)n_gpus);
    std::vector<std::thread> workers;
    for (int g = 0; g < n_gpus; ++g) {
        workers.emplace_back([&, g]() {
            const size_t r0 = rows * (size_t)g / (size_t)n_gpus, r1 = rows * (size_t)(g + 1) / (size_t)n_gpus;
            if (r0 >= r1) return;
            int r = TXP_OK;
            DeviceCtx* c = nullptr;
            const int dev = worker_device(g, n_gpus, caller_device);
            if (cudaSetDevice(dev) != cudaSuccess) r = fail(TXP_ERR_CUDA, "cudaSetDevice failed");
            if (r == TXP_OK) r = ensure_ctx(dev, &c);
            if (r == TXP_OK) {
                std::lock_guard<std::mutex> lk(c->mu);
                r = decompress_host_rows(*c, format, data + r0 * bw * bs, width, height, output + 4 * r0 * width * 4, r0, r1);
            }
            rcs[(size_t)g] = r;
            if (r != TXP_OK) errs[(size_t)g] = t_last_error;
        });
    }
    for (auto& t : workers) t.join();
    for (int g = 0; g < n_gpus; ++g)
        if (rcs[(size_t)g] != TXP_OK) return fail(rcs[(size_t)g], "gpu " + std::to_string(g) + ": " + errs[(size_t)g]);
    return TXP_OK;
}

int txp_decompress_batch(int format, const uint8_t* const* data, const size_t* widths, const size_t* heights, size_t n_textures,
                         uint8_t* const* outputs, int n_gpus) {
    if (format < 0 || format > 4) return fail(TXP_ERR_FORMAT, "format must be 0..4 (Bc1..Bc5)");
    if (n_textures == 0) return TXP_OK;
    if (!data || !widths || !heights || !outputs) return fail(TXP_ERR_ARGUMENT, "null pointer");
    const int ndev = txp_device_count();
    if (ndev < 0) return ndev;
    int caller_device = 0;
    TXP_CUDA(cudaGetDevice(&caller_device));
    if (n_gpus < 1 || n_gpus > ndev) return fail(TXP_ERR_ARGUMENT, "n_gpus must be between 1 and the device count");
    std::vector<int> rcs((size_t)n_gpus, TXP_OK);
    std::vector<std::string> errs((size_t)n_gpus);
    std::vector<std::thread> workers;
    for (int g = 0; g < n_gpus; ++g) {
        workers.emplace_back([&, g]() {
            int r = TXP_OK;
            DeviceCtx* c = nullptr;
            const int dev = worker_device(g, n_gpus, caller_device);
            if (cudaSetDevice(dev) != cudaSuccess) r = fail(TXP_ERR_CUDA, "cudaSetDevice failed");
            if (r == TXP_OK) r = ensure_ctx(dev, &c);
            if (r == TXP_OK) {
                std::lock_guard<std::mutex> lk(c->mu);
                size_t k = 0;
                for (size_t t = (size_t)g; t < n_textures && r == TXP_OK; t += (size_t)n_gpus) {
                    const size_t w = widths[t], h = heights[t];
                    if ((r = check_dims(w, h)) != TXP_OK) break;
                    if (!data[t] || !outputs[t]) { r = fail(TXP_ERR_ARGUMENT, "null texture pointer"); break; }
                    if (w * h == 0) continue;
                    const size_t in_bytes = txp_compressed_size(format, w, h), out_bytes = w * h * 4;
                    if (out_bytes <= CHUNK_BYTES) {
                        // a whole texture per pipeline slot: copies and kernels of neighbouring textures overlap
                        Slot& s = c->slots[k++ % NSLOTS];
                        if ((r = slot_wait(s)) != TXP_OK) break;
                        if ((r = grow_dev(&s.d_in, &s.d_in_cap, in_bytes)) != TXP_OK) break;
                        if ((r = grow_dev(&s.d_out, &s.d_out_cap, out_bytes)) != TXP_OK) break;
                        const uint8_t* src_ptr = data[t];
                        if (!dma_direct(src_ptr)) {
                            if ((r = grow_pinned(&s.h_in, &s.h_in_cap, in_bytes)) != TXP_OK) break;
                            host_copy(s.h_in, src_ptr, in_bytes);
                            src_ptr = s.h_in;
                        }
                        int rc = TXP_OK;
                        do {
                            TXP_CUDA_BREAK(cudaMemcpyAsync(s.d_in, src_ptr, in_bytes, cudaMemcpyDefault, s.stream));
                            if ((rc = launch_decode(format, s.d_in, (uint64_t)txp_num_blocks(w) * txp_num_blocks(h), (uint32_t)w, (uint32_t)h,
                                                    (uint32_t)txp_num_blocks(w), s.d_out, s.stream)) != TXP_OK) break;
                            if (dma_direct(outputs[t])) {
                                TXP_CUDA_BREAK(cudaMemcpyAsync(outputs[t], s.d_out, out_bytes, cudaMemcpyDefault, s.stream));
                            } else {
                                if ((rc = grow_pinned(&s.h_out, &s.h_out_cap, out_bytes)) != TXP_OK) break;
                                TXP_CUDA_BREAK(cudaMemcpyAsync(s.h_out, s.d_out, out_bytes, cudaMemcpyDeviceToHost, s.stream));
                                s.user_out = outputs[t]; s.user_out_bytes = out_bytes;
                            }
                            TXP_CUDA_BREAK(cudaEventRecord(s.done, s.stream));
                            s.busy = true;
                        } while (0);
                        r = rc;
                    } else {
                        for (Slot& s : c->slots) { const int r2 = slot_wait(s); if (r == TXP_OK) r = r2; }
                        if (r != TXP_OK) break;
                        r = decompress_host_rows(*c, format, data[t], w, h, outputs[t], 0, txp_num_blocks(h));
                    }
                }
                if (r != TXP_OK) { const std::string keep = t_last_error; slots_abandon(*c); t_last_error = keep; }
                for (Slot& s : c->slots) { const int r2 = slot_wait(s); if (r == TXP_OK) r = r2; }
            }
            rcs[(size_t)g] = r;
            if (r != TXP_OK) errs[(size_t)g] = t_last_error;
        });
    }
    for (auto& t : workers) t.join();
    for (int g = 0; g < n_gpus; ++g)
        if (rcs[(size_t)g] != TXP_OK) return fail(rcs[(size_t)g], "gpu " + std::to_string(g) + ": " + errs[(size_t)g]);
    return TXP_OK;
}

}  // extern "C"
