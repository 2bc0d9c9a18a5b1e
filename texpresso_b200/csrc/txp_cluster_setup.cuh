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
// With EMIT (lane-per-block search kernels, txp_cluster_lane.cuh) the thread also leaves the points of the set (4 bytes each:
// RGB + pixel count) and the pixel -> point map next to the record, so that the search kernels never touch the pixels.  The
// per-pixel loops are rolled over a shared-memory column (txp_block_rolled.cuh).
#pragma once
#include "txp_range.cuh"

namespace txp {

// setup record: .x/.y = ordering word (4 bits per sorted position: point index, 0 for padding), .z = flags
constexpr uint32_t SETUP_SEARCH = 0x100u;       // block needs the partition search (>= 2 points)
constexpr uint32_t SETUP_DEGENERATE = 0x200u;   // some projection is NaN/inf: ordering has repeated entries
constexpr uint32_t SETUP_TRANSPARENT = 0x400u;  // BC1 punch-through pixels present: no 4-colour pass (colourfit.rs:51)
// .z bits 0..4 = number of points, bits 16..31 = pixels that belong to a point (valid and not punched through)

// Lane-per-block search kernels (EMIT): the record of chunk-local block lb is rec_quads() uint4 at rec[lb * quads]:
//   [0]      the setup record above
//   [1..4]   point p of the colour set, p = 0..15 in set order: RGB of its first pixel | number of pixels << 24
//   [5]      .x/.y = pixel -> point map (colourset.rs:130-141), 4 bits per pixel                     -- 6 quads = 96 bytes
// and with Params::weigh_colour_by_alpha (weights are alpha sums, not pixel counts):
//   [1..4]   RGB of point p   [5..8] weight of point p (fp32, colourset.rs:107-109)   [9] pixel -> point map  -- 10 quads
// The search kernels turn a point into (r, g, b) / 255 and sqrt(count) with two small shared-memory tables.  96 + 4 (permutation)
// bytes of scratch per block; round 1 kept 16 float4 points + a remap table = 292 bytes.
__host__ __device__ __forceinline__ int rec_quads(const bool emit, const bool alpha_weighted) { return emit ? (alpha_weighted ? 10 : 6) : 1; }

constexpr int SETUP_BINS = 35;                   // sort key: points (2..16) + 17 if punch-through; 34 = no search needed

// One block, chunk-local number lb.  Returns the block's sort key.  col: the thread's shared-memory column.
template <int FMT, bool EMIT>
__device__ __forceinline__ int cluster_setup_block(const BlockSource& src, const EncodeParams& prm, uint8_t* __restrict__ out,
                                                    uint4* __restrict__ setup, const float* lut, float4* col, const uint64_t first, const uint32_t lb) {
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
    const bool aw = prm.alpha_weighted != 0;
    uint4* rec = setup + (size_t)lb * rec_quads(EMIT, aw);
    const RolledSet ts = rolled_colourset<FMT == BC1>(px, mask, aw, col);
    if (ts.active16 == 0) {                              // lib.rs:223 -> RangeFit on an empty set
        *colour_out = FMT == BC1 ? make_uint2(0u, 0xFFFFFFFFu) : make_uint2(0u, 0u);
        rec[0] = make_uint4(0u, 0u, 0u, 0u);
        return SETUP_BINS - 1;
    }
    if ((ts.new16 & (ts.new16 - 1u)) == 0u) {            // one point: SingleColourFit (lib.rs:217-222)
        *colour_out = single_fit_thread<FMT == BC1>(rolled_single_rgb(px, ts.active16), ts.active16, ts.transparent);
        rec[0] = make_uint4(0u, 0u, 1u, 0u);
        return SETUP_BINS - 1;
    }
    rolled_fill_points<EMIT>(px, ts, aw, lut, col, reinterpret_cast<uint32_t*>(rec + 1), reinterpret_cast<float*>(rec + 5));
    if (EMIT) {                                          // pixel -> point: the point of the first pixel with the same key
        uint32_t rlo = 0, rhi = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t f = ((i < 8 ? ts.first_lo : ts.first_hi) >> (4 * (i & 7))) & 15u;
            const uint32_t q = (uint32_t)__popc(ts.new16 & ((1u << f) - 1u));
            if (i < 8) rlo |= q << (4 * i); else rhi |= q << (4 * (i - 8));
        }
        rec[aw ? 9 : 5] = make_uint4(rlo, rhi, 0u, 0u);
    }
    const float3 axis = rolled_principal_axis(col);

    // ---- construct_ordering (cluster.rs:78-105) on the principal axis --------------------------------------------
    // keys: finite projections in float order < padding (the reference's (0, f32::MAX) entries) < non-finite
    // projections; ties keep index order (stable insertion sort)
    int sk[16];
    bool degenerate = false;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        int k = 0x7FFFFFFE;                              // padding
        if ((ts.new16 >> i) & 1u) {
            const float4 q = col[i * ROLL_THREADS];
            const uint32_t bits = __float_as_uint(add(add(mul(q.x, axis.x), mul(q.y, axis.y)), mul(q.z, axis.z)));
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
    rec[0] = make_uint4(lo, hi, (uint32_t)count | SETUP_SEARCH | (degenerate ? SETUP_DEGENERATE : 0u) |
                                (ts.transparent ? SETUP_TRANSPARENT : 0u) | (ts.active16 << 16), 0u);
    return count + (ts.transparent ? 17 : 0);
}

// Blocks [first, first + n) of `src`; the records are indexed by the chunk-local block number.
#ifndef TXP_SETUP_MIN_CTAS
#define TXP_SETUP_MIN_CTAS 5
#endif
template <int FMT>
__global__ void __launch_bounds__(ROLL_THREADS, TXP_SETUP_MIN_CTAS) cluster_setup_kernel(const BlockSource src, const EncodeParams prm,
                                                            uint8_t* __restrict__ out, uint4* __restrict__ setup,
                                                            const uint64_t first, const uint32_t n) {
    __shared__ float lut[256];
    __shared__ float4 s_col[16][ROLL_THREADS];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = fdiv((float)i, 255.0f);   // colourset.rs:65-67
    __syncthreads();
    const uint32_t lb = blockIdx.x * blockDim.x + threadIdx.x;
    if (lb >= n) return;
    cluster_setup_block<FMT, false>(src, prm, out, setup, lut, &s_col[0][threadIdx.x], first, lb);
}

// Setup for the lane-per-block search kernel.  One CTA owns a window of SETUP_WINDOW consecutive blocks (one block per
// thread and round) and also leaves the window's permutation sorted by (points, punch-through), blocks that need no
// search last: perm[s] = chunk-local block number of the s-th record in sorted order.  The 32 lanes of a search warp take
// 32 consecutive entries of perm, i.e. (mostly) blocks whose loop nests have the same shape.
#ifndef TXP_SETUP_WINDOW_ROUNDS
#define TXP_SETUP_WINDOW_ROUNDS 8    // blocks per window = 128 x rounds (A/B: larger windows group the blocks better but leave the setup kernel fewer CTAs)
#endif
constexpr int SETUP_WINDOW_ROUNDS = TXP_SETUP_WINDOW_ROUNDS;
constexpr int SETUP_WINDOW = ROLL_THREADS * SETUP_WINDOW_ROUNDS;

template <int FMT>
__global__ void __launch_bounds__(ROLL_THREADS, TXP_SETUP_MIN_CTAS) cluster_setup_sorted_kernel(const BlockSource src, const EncodeParams prm,
                                                                   uint8_t* __restrict__ out, uint4* __restrict__ rec,
                                                                   uint32_t* __restrict__ perm, const uint64_t first, const uint32_t n) {
    __shared__ float lut[256];
    __shared__ float4 s_col[16][ROLL_THREADS];
    __shared__ int hist[SETUP_BINS];
    __shared__ uint16_t s_rank[SETUP_WINDOW];
    __shared__ uint8_t s_key[SETUP_WINDOW];
    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += ROLL_THREADS) lut[i] = fdiv((float)i, 255.0f);          // colourset.rs:65-67
    if (tid < SETUP_BINS) hist[tid] = 0;
    __syncthreads();
    const uint32_t win0 = blockIdx.x * SETUP_WINDOW;
#pragma unroll 1
    for (int r = 0; r < SETUP_WINDOW_ROUNDS; ++r) {
        const uint32_t lb = win0 + r * ROLL_THREADS + tid;
        if (lb >= n) break;
        const int key = cluster_setup_block<FMT, true>(src, prm, out, rec, lut, &s_col[0][tid], first, lb);
        s_key[r * ROLL_THREADS + tid] = (uint8_t)key;
        s_rank[r * ROLL_THREADS + tid] = (uint16_t)atomicAdd(&hist[key], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int k = 0; k < SETUP_BINS; ++k) { const int c = hist[k]; hist[k] = run; run += c; }
    }
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < SETUP_WINDOW_ROUNDS; ++r) {
        const uint32_t lb = win0 + r * ROLL_THREADS + tid;
        if (lb >= n) break;
        perm[win0 + (uint32_t)hist[s_key[r * ROLL_THREADS + tid]] + s_rank[r * ROLL_THREADS + tid]] = lb;
    }
}

}  // namespace txp
