// txp_cluster_setup.cuh -- ClusterFit setup kernel ("K1"), one THREAD per 4x4 block.
//
// Everything of ClusterFit that happens once per block before the partition search and is serial in nature is done
// here 32 blocks per warp instead of one block per warp:
//   * BC2 / BC3 alpha half (alpha.rs:27-51 / :187-256) -> written straight to the output
//   * ColourSet (colourset.rs:35-112); blocks with no point or one point are finished here (lib.rs:217-225,
//     single.rs) and flagged as done
//   * Sym3x3::weighted_covariance + principle_component (math.rs:44-97)
//   * construct_ordering for the principal axis (cluster.rs:78-105), which is iteration 0 of compress3 AND compress4
// The search kernel (txp_colour.cuh) receives 16 bytes per block: the ordering word and flags.
#pragma once
#include "txp_range.cuh"

namespace txp {

// setup record: .x/.y = ordering word (4 bits per sorted position: point index, 0 for padding), .z = flags
constexpr uint32_t SETUP_SEARCH = 0x100u;       // block needs the partition search (>= 2 points)
constexpr uint32_t SETUP_DEGENERATE = 0x200u;   // some projection is NaN/inf: ordering has repeated entries

template <int FMT>
__global__ void __launch_bounds__(128) cluster_setup_kernel(const BlockSource src, const EncodeParams prm,
                                                            uint8_t* __restrict__ out, uint4* __restrict__ setup) {
    __shared__ float lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = fdiv((float)i, 255.0f);   // colourset.rs:65-67
    __syncthreads();
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= src.nblocks) return;
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
        setup[b] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    if ((ts.new16 & (ts.new16 - 1u)) == 0u) {            // one point: SingleColourFit (lib.rs:217-222)
        *colour_out = single_fit_thread<FMT == BC1>(thread_single_rgb(px, ts.active16), ts.active16, ts.transparent);
        setup[b] = make_uint4(0u, 0u, 1u, 0u);
        return;
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
    int rank[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) rank[i] = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
#pragma unroll
        for (int j = i + 1; j < 16; ++j) {
            const bool lt = sk[j] < sk[i];               // strict: on ties the earlier entry stays first
            rank[i] += lt ? 1 : 0;
            rank[j] += lt ? 0 : 1;
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
    setup[b] = make_uint4(lo, hi, (uint32_t)__popc(ts.new16) | SETUP_SEARCH | (degenerate ? SETUP_DEGENERATE : 0u), 0u);
}

}  // namespace txp
