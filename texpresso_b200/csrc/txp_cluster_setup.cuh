// txp_cluster_setup.cuh -- ClusterFit setup kernel ("K1"), one THREAD per 4x4 block.
//
// Everything of ClusterFit that happens once per block before the partition search and is serial in nature is done
// here 32 blocks per warp instead of one block per warp:
//   * BC2 / BC3 alpha half (alpha.rs:27-51 / :187-256) -> written straight to the output
//   * ColourSet (colourset.rs:35-112); blocks with no point or one point are finished here (lib.rs:217-225,
//     single.rs) and flagged as done
//   * Sym3x3::weighted_covariance + principle_component (math.rs:44-97)
//   * construct_ordering for the principal axis (cluster.rs:78-105), which is iteration 0 of compress3 AND compress4
// The warp-per-block search kernel (txp_colour.cuh) receives 16 bytes per block: the ordering word and flags.
// With EMIT (lane-per-block search kernels, txp_cluster_lane.cuh) the thread also leaves the points of the set with
// their weights (256 B) and the pixel -> point remap (colourset.rs:130-141), so that the search kernels never touch the
// pixels.
#pragma once
#include "txp_range.cuh"

namespace txp {

// setup record: .x/.y = ordering word (4 bits per sorted position: point index, 0 for padding), .z = flags
constexpr uint32_t SETUP_SEARCH = 0x100u;       // block needs the partition search (>= 2 points)
constexpr uint32_t SETUP_DEGENERATE = 0x200u;   // some projection is NaN/inf: ordering has repeated entries
constexpr uint32_t SETUP_TRANSPARENT = 0x400u;  // BC1 punch-through pixels present: no 4-colour pass (colourfit.rs:51)
// .z bits 0..4 = number of points, bits 16..31 = pixels that belong to a point (valid and not punched through)

// point p of chunk-local block lb: 256 contiguous bytes per block (whole sectors per writer)
__host__ __device__ __forceinline__ size_t pt_index(const uint32_t lb, const int p) { return (size_t)lb * 16 + p; }

constexpr int SETUP_BINS = 35;                   // sort key: points (2..16) + 17 if punch-through; 34 = no search needed

// One block, chunk-local number lb.  Returns the block's sort key.
template <int FMT, bool EMIT>
__device__ __forceinline__ int cluster_setup_block(const BlockSource& src, const EncodeParams& prm, uint8_t* __restrict__ out,
                                                    uint4* __restrict__ setup, uint2* __restrict__ remap, float4* __restrict__ ptbuf,
                                                    const float* lut, const uint64_t first, const uint32_t lb) {
    const uint64_t b = first + lb;
    uint32_t px[16];
    uint32_t mask;
    load_block_thread(src, b, px, mask);
    uint2* out2 = reinterpret_cast<uint2*>(out);
    if (FMT == BC2) out2[2 * b] = alpha_bc2_thread(px, mask);
    if (FMT == BC3) {
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = px[i] >> 24;
        out2[2 * b] = mask == 0xFFFFu ? alpha_fit_full(v) : alpha_fit_thread(v, mask);
    }
    uint2* colour_out = FMT == BC1 ? out2 + b : out2 + 2 * b + 1;

    uint32_t gw[16];
    const ThreadSet ts = thread_colourset<FMT == BC1>(px, mask, prm.alpha_weighted != 0, gw);
    if (ts.active16 == 0) {                              // lib.rs:223 -> RangeFit on an empty set
        *colour_out = FMT == BC1 ? make_uint2(0u, 0xFFFFFFFFu) : make_uint2(0u, 0u);
        setup[lb] = make_uint4(0u, 0u, 0u, 0u);
        return SETUP_BINS - 1;
    }
    if ((ts.new16 & (ts.new16 - 1u)) == 0u) {            // one point: SingleColourFit (lib.rs:217-222)
        *colour_out = single_fit_thread<FMT == BC1>(thread_single_rgb(px, ts.active16), ts.active16, ts.transparent);
        setup[lb] = make_uint4(0u, 0u, 1u, 0u);
        return SETUP_BINS - 1;
    }
    float w[16];
    thread_weights(gw, ts.new16, prm.alpha_weighted != 0, w);
    const float3 axis = thread_principal_axis(px, w, lut);

    // ---- construct_ordering (cluster.rs:78-105) on the principal axis --------------------------------------------
    // keys: finite projections in float order < padding (the reference's (0, f32::MAX) entries) < non-finite
    // projections; ties keep index order (stable insertion sort)
    int sk[16];
    bool degenerate = false;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        int k = 0x7FFFFFFE;                              // padding
        if ((ts.new16 >> i) & 1u) {
            const float x = lut[px[i] & 255u], y = lut[(px[i] >> 8) & 255u], z = lut[(px[i] >> 16) & 255u];
            const uint32_t bits = __float_as_uint(add(add(mul(x, axis.x), mul(y, axis.y)), mul(z, axis.z)));
            if ((bits & 0x7F800000u) == 0x7F800000u) { k = 0x7FFFFFFF; degenerate = true; }
            else k = (bits & 0x80000000u) ? -(int)(bits & 0x7FFFFFFFu) : (int)bits;
        }
        sk[i] = k;
    }
    // rank[e] = #{f < e: sk[f] <= sk[e]} + #{f > e: sk[f] < sk[e]} = e - (earlier entries that are larger) + (later entries that are smaller)
    int rank[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) rank[i] = i;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
#pragma unroll
        for (int j = i + 1; j < 16; ++j) {
            if (sk[j] < sk[i]) { ++rank[i]; --rank[j]; }   // strict: on ties the earlier entry stays first
        }
    }
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if ((ts.new16 >> i) & 1u) {
            const uint32_t p = (uint32_t)__popc(ts.new16 & ((1u << i) - 1u));     // point index of pixel i
            if (rank[i] < 8) lo |= p << (4 * rank[i]); else hi |= p << (4 * (rank[i] - 8));
        }
    }
    const int count = __popc(ts.new16);
    setup[lb] = make_uint4(lo, hi, (uint32_t)count | SETUP_SEARCH | (degenerate ? SETUP_DEGENERATE : 0u) |
                                   (ts.transparent ? SETUP_TRANSPARENT : 0u) | (ts.active16 << 16), 0u);
    if (EMIT) {
        // pixel -> point (colourset.rs:84-88): the point of the first pixel with the same key
        uint32_t rlo = 0, rhi = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            int f = i;
#pragma unroll
            for (int j = i - 1; j >= 0; --j) if (px[j] == px[i]) f = j;
            const uint32_t p = (uint32_t)__popc(ts.new16 & ((1u << f) - 1u));
            if (i < 8) rlo |= p << (4 * i); else rhi |= p << (4 * (i - 8));
        }
        remap[lb] = make_uint2(rlo, rhi);
        // the points of the set, in set order, with their weights (colourset.rs:65-67, :107-109); the search kernels form
        // points_weights = (x, y, z, 1) * w in the order of the current axis themselves (cluster.rs:123-133)
        int p = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if ((ts.new16 >> i) & 1u) {
                ptbuf[pt_index(lb, p)] = make_float4(lut[px[i] & 255u], lut[(px[i] >> 8) & 255u], lut[(px[i] >> 16) & 255u], w[i]);
                ++p;
            }
        }
    }
    return count + (ts.transparent ? 17 : 0);
}

// Blocks [first, first + n) of `src`; the records are indexed by the chunk-local block number.
template <int FMT>
__global__ void __launch_bounds__(128) cluster_setup_kernel(const BlockSource src, const EncodeParams prm,
                                                            uint8_t* __restrict__ out, uint4* __restrict__ setup,
                                                            const uint64_t first, const uint32_t n) {
    __shared__ float lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = fdiv((float)i, 255.0f);   // colourset.rs:65-67
    __syncthreads();
    const uint32_t lb = blockIdx.x * blockDim.x + threadIdx.x;
    if (lb >= n) return;
    cluster_setup_block<FMT, false>(src, prm, out, setup, nullptr, nullptr, lut, first, lb);
}

// Setup for the lane-per-block search kernel.  One CTA owns a window of SETUP_WINDOW consecutive blocks (one block per
// thread and round) and also leaves the window's permutation sorted by (points, punch-through), blocks that need no
// search last: perm[s] = chunk-local block number of the s-th record in sorted order.  The 32 lanes of a search warp take
// 32 consecutive entries of perm, i.e. (mostly) blocks whose loop nests have the same shape.
constexpr int SETUP_WINDOW_ROUNDS = 8;
constexpr int SETUP_WINDOW = 128 * SETUP_WINDOW_ROUNDS;

#ifndef TXP_SETUP_MIN_CTAS
#define TXP_SETUP_MIN_CTAS 4
#endif
template <int FMT>
__global__ void __launch_bounds__(128, TXP_SETUP_MIN_CTAS) cluster_setup_sorted_kernel(const BlockSource src, const EncodeParams prm,
                                                                   uint8_t* __restrict__ out, uint4* __restrict__ setup,
                                                                   uint2* __restrict__ remap, float4* __restrict__ ptbuf,
                                                                   uint32_t* __restrict__ perm, const uint64_t first, const uint32_t n) {
    __shared__ float lut[256];
    __shared__ int hist[SETUP_BINS];
    __shared__ uint16_t s_rank[SETUP_WINDOW];
    __shared__ uint8_t s_key[SETUP_WINDOW];
    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += 128) lut[i] = fdiv((float)i, 255.0f);                  // colourset.rs:65-67
    if (tid < SETUP_BINS) hist[tid] = 0;
    __syncthreads();
    const uint32_t win0 = blockIdx.x * SETUP_WINDOW;
#pragma unroll 1
    for (int r = 0; r < SETUP_WINDOW_ROUNDS; ++r) {
        const uint32_t lb = win0 + r * 128 + tid;
        if (lb >= n) break;
        const int key = cluster_setup_block<FMT, true>(src, prm, out, setup, remap, ptbuf, lut, first, lb);
        s_key[r * 128 + tid] = (uint8_t)key;
        s_rank[r * 128 + tid] = (uint16_t)atomicAdd(&hist[key], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int k = 0; k < SETUP_BINS; ++k) { const int c = hist[k]; hist[k] = run; run += c; }
    }
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < SETUP_WINDOW_ROUNDS; ++r) {
        const uint32_t lb = win0 + r * 128 + tid;
        if (lb >= n) break;
        perm[win0 + (uint32_t)hist[s_key[r * 128 + tid]] + s_rank[r * 128 + tid]] = lb;
    }
}

}  // namespace txp