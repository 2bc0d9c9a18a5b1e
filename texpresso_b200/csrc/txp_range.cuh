// txp_range.cuh -- RangeFit encoder (BC1 / BC2 / BC3), one THREAD per 4x4 block.
//
// Replaces (reference): lib.rs:188-234 for Algorithm::RangeFit, colourset.rs:35-141, colourfit/range.rs:44-192,
// colourfit/single.rs:58-164 (count == 1), math.rs:44-97, colourblock.rs:28-94, and the alpha half
// (alpha.rs:27-51 / :187-256).
//
// RangeFit is ~2k dependent flops per block with no search to parallelise, so a warp per block (the ClusterFit
// layout) wastes 31 lanes on the serial parts.  Here every lane owns a whole block:
//   * the colour set is not compacted.  A pixel is "new" if no earlier active pixel has its RGB; all
//     order-sensitive sums (math.rs:48-70, range.rs:129) run over the 16 pixels in order with weight / term 0 for
//     pixels that are not new -- x + (+-0) is an exact no-op for accumulators that start at +0 -- which is the same
//     left-to-right order as the reference's loops over the compacted set.
//   * nearest-code indices are computed per pixel (duplicates of a point get the same answer), so no remap table.
//   * c/255 comes from a 256-entry shared table built with IEEE division (colourset.rs:65-67).
#pragma once
#include <cfloat>
#include "txp_common.cuh"
#include "txp_alpha.cuh"
#include "txp_block_rolled.cuh"

namespace txp {

__device__ uint8_t g_single_lut[6144];        // copy of the SingleColourFit table for per-thread (divergent) lookups

// spread the low 16 bits of m to the even bit positions of a 32-bit word
__device__ __forceinline__ uint32_t spread16(uint32_t m) {
    m = (m | (m << 8)) & 0x00FF00FFu;
    m = (m | (m << 4)) & 0x0F0F0F0Fu;
    m = (m | (m << 2)) & 0x33333333u;
    m = (m | (m << 1)) & 0x55555555u;
    return m;
}

// single.rs:58-106 with per-thread table reads
__device__ __forceinline__ void single_endpoints_thread(const uint32_t rgb, const int t0, const int t1, const int t2,
                                                        uint32_t& a565, uint32_t& b565, uint32_t& index, uint32_t& error) {
    const int tabs[3] = {t0, t1, t2};
    const uint32_t col[3] = {rgb & 255u, (rgb >> 8) & 255u, (rgb >> 16) & 255u};
    error = 0xFFFFFFFFu; a565 = 0; b565 = 0; index = 0;
#pragma unroll
    for (int idx = 0; idx < 2; ++idx) {
        uint32_t e = 0, st[3], en[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint8_t* p = &g_single_lut[((tabs[c] * 256 + col[c]) * 2 + idx) * 3];
            st[c] = __ldg(p); en[c] = __ldg(p + 1);
            const uint32_t d = __ldg(p + 2);
            e += d * d;
        }
        if (e < error) {
            a565 = (st[0] << 11) | (st[1] << 5) | st[2];
            b565 = (en[0] << 11) | (en[1] << 5) | en[2];
            index = 2u * idx; error = e;
        }
    }
}

// single.rs:122-164 + colourfit.rs:48-59.  active16: pixels that carry the colour; others get index 3.
template <bool IS_BC1>
__device__ __forceinline__ uint2 single_fit_thread(const uint32_t rgb, const uint32_t active16, const bool transparent) {
    const uint32_t act = spread16(active16), inact = spread16(~active16 & 0xFFFFu) * 3u;
    uint32_t a, b, index, err, best = 0xFFFFFFFFu;
    uint2 block = make_uint2(0u, 0u);
    if (IS_BC1) {
        single_endpoints_thread(rgb, 0, 1, 0, a, b, index, err);
        block = write3_packed(a, b, act * index | inact);
        best = err;
        if (transparent) return block;
    }
    single_endpoints_thread(rgb, 2, 3, 2, a, b, index, err);
    if (err < best) block = write4_packed(a, b, act * index | inact);
    return block;
}

// alpha.rs:27-51, one thread
__device__ __forceinline__ uint2 alpha_bc2_thread(const uint32_t px[16], const uint32_t mask) {
    uint32_t w[2] = {0u, 0u};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float a = mul((float)(px[i] >> 24), 15.0f / 255.0f);
        uint32_t q = (uint32_t)f32_to_i32_clamped(a, 15);
        if (!((mask >> i) & 1u)) q = 0;
        w[i >> 3] |= q << (4 * (i & 7));
    }
    return make_uint2(w[0], w[1]);
}

// ---- per-thread ColourSet (colourset.rs:35-112), shared by the RangeFit kernel and the ClusterFit setup kernel ----
struct ThreadSet {
    uint32_t active16;       // valid and not punched through
    uint32_t new16;          // first occurrence of its RGB among the active pixels  (== the points of the set, in order)
    bool transparent;        // BC1 punch-through present (colourset.rs:54-58)
};

// px[] is rewritten in place to comparison keys (RGB for active pixels, a unique value otherwise);
// gw[i] = integer weight total of the group whose first pixel is i (1 per pixel, or alpha+1: exact sums)
template <bool IS_BC1>
__device__ __forceinline__ ThreadSet thread_colourset(uint32_t px[16], const uint32_t mask, const bool alpha_weighted, uint32_t gw[16]) {
    ThreadSet t;
    t.active16 = 0;
    uint32_t punched = 0;
    uint32_t wgt[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const bool valid = (mask >> i) & 1u;
        const bool pt = IS_BC1 && valid && (px[i] >> 24) < 128u;                  // :54
        if (pt) punched |= 1u << i;
        if (valid && !pt) t.active16 |= 1u << i;
        wgt[i] = alpha_weighted ? (px[i] >> 24) + 1u : 1u;
        px[i] = (valid && !pt) ? (px[i] & 0x00FFFFFFu) : (0x01000000u | (uint32_t)i);
        gw[i] = wgt[i];
    }
    t.transparent = punched != 0;
    // exact-RGB duplicates (:84-88): pixel i is new iff no earlier pixel has its key; group weights are
    // accumulated on every earlier equal pixel (only the first one's total is used)
    uint32_t dup16 = 0;
#pragma unroll
    for (int i = 1; i < 16; ++i) {
        bool dup = false;
#pragma unroll
        for (int j = 0; j < i; ++j) {
            const bool eq = px[i] == px[j];
            dup |= eq;
            if (eq) gw[j] += wgt[i];
        }
        if (dup) dup16 |= 1u << i;
    }
    t.new16 = t.active16 & ~dup16;
    return t;
}

// weights: sqrt of the group totals (colourset.rs:107-109); 0 for pixels that are not new
__device__ __forceinline__ void thread_weights(const uint32_t gw[16], const uint32_t new16, const bool alpha_weighted, float w[16]) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float t = 0.0f;
        if ((new16 >> i) & 1u) {
            t = 1.0f;
            if (gw[i] != 1u || alpha_weighted)
                t = __fsqrt_rn(alpha_weighted ? mul((float)gw[i], 1.0f / 256.0f) : (float)gw[i]);
        }
        w[i] = t;
    }
}

// Sym3x3::weighted_covariance + principle_component (math.rs:44-97) over the 16 pixels with weight 0 for non-points
__device__ __forceinline__ float3 thread_principal_axis(const uint32_t px[16], const float w[16], const float* __restrict__ lut) {
    float total = 0.0f, cx = 0.0f, cy = 0.0f, cz = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float x = lut[px[i] & 255u], y = lut[(px[i] >> 8) & 255u], z = lut[(px[i] >> 16) & 255u];
        total = add(total, w[i]);
        cx = add(cx, mul(x, w[i])); cy = add(cy, mul(y, w[i])); cz = add(cz, mul(z, w[i]));
    }
    if (total > FLT_EPSILON) { cx = fdiv(cx, total); cy = fdiv(cy, total); cz = fdiv(cz, total); }
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f, m5 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float x = lut[px[i] & 255u], y = lut[(px[i] >> 8) & 255u], z = lut[(px[i] >> 16) & 255u];
        const float ax = sub(x, cx), ay = sub(y, cy), az = sub(z, cz);
        const float bx = mul(ax, w[i]), by = mul(ay, w[i]), bz = mul(az, w[i]);
        m0 = add(m0, mul(ax, bx)); m1 = add(m1, mul(ax, by)); m2 = add(m2, mul(ax, bz));
        m3 = add(m3, mul(ay, by)); m4 = add(m4, mul(ay, bz)); m5 = add(m5, mul(az, bz));
    }
    float vx = 1.0f, vy = 1.0f, vz = 1.0f;
#pragma unroll 1
    for (int it = 0; it < 8; ++it) {
        const float tx = add(mul(m2, vz), add(mul(m1, vy), mul(m0, vx)));
        const float ty = add(mul(m4, vz), add(mul(m3, vy), mul(m1, vx)));
        const float tz = add(mul(m5, vz), add(mul(m4, vy), mul(m2, vx)));
        const float ra = rcp(fmaxf(tx, fmaxf(ty, tz)));
        vx = mul(tx, ra); vy = mul(ty, ra); vz = mul(tz, ra);
    }
    return make_float3(vx, vy, vz);
}

// the one colour of a single-colour block: every active pixel carries it
__device__ __forceinline__ uint32_t thread_single_rgb(const uint32_t px[16], const uint32_t active16) {
    uint32_t rgb = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) if ((active16 >> i) & 1u) rgb |= px[i];
    return rgb;
}

// Colour half of one block with Algorithm::RangeFit (range.rs:44-192).  px: 16 RGBA words (rewritten to keys), mask: valid
// bits, lut: c/255 table, col: the thread's shared-memory column (txp_block_rolled.cuh).
template <bool IS_BC1>
__device__ __forceinline__ uint2 range_colour_rolled(uint32_t px[16], const uint32_t mask, const EncodeParams& prm,
                                                     const float* __restrict__ lut, float4* col) {
    const RolledSet ts = rolled_colourset<IS_BC1>(px, mask, prm.alpha_weighted != 0, col);
    const uint32_t active16 = ts.active16, new16 = ts.new16;
    if (active16 == 0)                                   // lib.rs:223, SURVEY Q14
        return IS_BC1 ? make_uint2(0u, 0xFFFFFFFFu) : make_uint2(0u, 0u);
    if ((new16 & (new16 - 1u)) == 0u)                    // exactly one distinct colour: lib.rs:217-222
        return single_fit_thread<IS_BC1>(rolled_single_rgb(px, active16), active16, ts.transparent);
    rolled_fill_points<false>(px, ts, prm.alpha_weighted != 0, lut, col);
    const float3 axis = rolled_principal_axis(col);
    // ---- range.rs:67-86: first point starts both ends; strict < / else-if > over the following points ----------
    int is = 0, ie = 0;
    float mn = 0.f, mx = 0.f;
    if (__all_sync(__activemask(), active16 == 0xFFFFu)) {           // warp-uniform: a warp with both kinds of block would pay for both loops
        // Every pixel carries a colour (the usual case), so pixel 0 is the first point.  A pixel that is not a point repeats the colour of an
        // earlier point, hence its projection bit for bit, and the strict comparisons ignore it: no is-a-point test per pixel.  mn <= mx always,
        // so "d < mn" and "d > mx" exclude each other and the reference's else-if needs no spelling out.  (NaN projections compare false throughout.)
        const float4 p0 = col[0];
        mn = mx = add(add(mul(p0.x, axis.x), mul(p0.y, axis.y)), mul(p0.z, axis.z));
#pragma unroll 5
        for (int i = 1; i < 16; ++i) {
            const float4 p = col[i * ROLL_THREADS];
            const float d = add(add(mul(p.x, axis.x), mul(p.y, axis.y)), mul(p.z, axis.z));
            if (d < mn) { is = i; mn = d; }
            if (d > mx) { ie = i; mx = d; }
        }
    } else {
        bool found = false;
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const float4 p = col[i * ROLL_THREADS];
            const bool is_new = p.w > 0.0f;              // weights of points are sqrt of positive totals
            const float d = add(add(mul(p.x, axis.x), mul(p.y, axis.y)), mul(p.z, axis.z));
            const bool first = is_new && !found;
            const bool lower = is_new && found && d < mn;
            const bool upper = is_new && found && !(d < mn) && d > mx;
            if (first || lower) { is = i; mn = d; }
            if (first || upper) { ie = i; mx = d; }
            found = found || is_new;
        }
    }
    // clamp to [0,1] is the identity on c/255; snap to the 5:6:5 grid (range.rs:88-98)
    const float4 ps = col[is * ROLL_THREADS], pe = col[ie * ROLL_THREADS];
    const float sa[3] = {ps.x, ps.y, ps.z}, ea[3] = {pe.x, pe.y, pe.z};
    const float grid[3] = {31.0f, 63.0f, 31.0f};
    const float gridrcp[3] = {1.0f / 31.0f, 1.0f / 63.0f, 1.0f / 31.0f};
    float sv[3], ev[3];
    uint32_t ks[3], ke[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float a = grid_index(grid[c], sa[c]);
        const float b = grid_index(grid[c], ea[c]);
        ks[c] = (uint32_t)a; ke[c] = (uint32_t)b;
        sv[c] = mul(a, gridrcp[c]); ev[c] = mul(b, gridrcp[c]);
    }
    const uint32_t a565 = (ks[0] << 11) | (ks[1] << 5) | ks[2];
    const uint32_t b565 = (ke[0] << 11) | (ke[1] << 5) | ke[2];
    const float wx = prm.wx, wy = prm.wy, wz = prm.wz;
    const uint32_t act2 = spread16(active16) * 3u, inact = spread16(~active16 & 0xFFFFu) * 3u;   // colourset.rs:134-137

    // ---- compress3 / compress4 (range.rs:103-192, colourfit.rs:48-59) -------------------------------------------
    // Both codebooks start with (start, end): the two passes share those distances, so one loop over the pixels evaluates
    // five codes instead of 3 + 4 (same operations on the same operands => the same bits as the reference's two loops).
    const bool do3 = IS_BC1, do4 = !(IS_BC1 && ts.transparent);
    float m3[3], t4a[3], t4b[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        m3[c] = add(mul(sv[c], 0.5f), mul(ev[c], 0.5f));                           // range.rs:161
        t4a[c] = add(mul(sv[c], 2.0f / 3.0f), mul(ev[c], 1.0f / 3.0f));           // range.rs:179
        t4b[c] = add(mul(sv[c], 1.0f / 3.0f), mul(ev[c], 2.0f / 3.0f));           // range.rs:180
    }
    float err3 = 0.0f, err4 = 0.0f;
    uint32_t idx3 = 0, idx4 = 0;
#pragma unroll 2
    for (int i = 0; i < 16; ++i) {
        const float4 p = col[i * ROLL_THREADS];
        // weights are applied before squaring (range.rs:117); the first minimum wins (range.rs:118, strict <)
        float dx = mul(wx, sub(p.x, sv[0])), dy = mul(wy, sub(p.y, sv[1])), dz = mul(wz, sub(p.z, sv[2]));
        const float d0 = add(add(mul(dx, dx), mul(dy, dy)), mul(dz, dz));
        dx = mul(wx, sub(p.x, ev[0])); dy = mul(wy, sub(p.y, ev[1])); dz = mul(wz, sub(p.z, ev[2]));
        const float d1 = add(add(mul(dx, dx), mul(dy, dy)), mul(dz, dz));
        float dist = d0;
        uint32_t idx = 0;
        if (d1 < dist) { dist = d1; idx = 1u; }
        const bool is_new = p.w > 0.0f;
        if (IS_BC1) {
            dx = mul(wx, sub(p.x, m3[0])); dy = mul(wy, sub(p.y, m3[1])); dz = mul(wz, sub(p.z, m3[2]));
            const float d = add(add(mul(dx, dx), mul(dy, dy)), mul(dz, dz));
            float dist3 = dist;
            uint32_t i3 = idx;
            if (d < dist3) { dist3 = d; i3 = 2u; }
            err3 = add(err3, is_new ? dist3 : 0.0f);                                // range.rs:129, set order
            idx3 |= i3 << (2 * i);
        }
        dx = mul(wx, sub(p.x, t4a[0])); dy = mul(wy, sub(p.y, t4a[1])); dz = mul(wz, sub(p.z, t4a[2]));
        float d = add(add(mul(dx, dx), mul(dy, dy)), mul(dz, dz));
        if (d < dist) { dist = d; idx = 2u; }
        dx = mul(wx, sub(p.x, t4b[0])); dy = mul(wy, sub(p.y, t4b[1])); dz = mul(wz, sub(p.z, t4b[2]));
        d = add(add(mul(dx, dx), mul(dy, dy)), mul(dz, dz));
        if (d < dist) { dist = d; idx = 3u; }
        err4 = add(err4, is_new ? dist : 0.0f);
        idx4 |= idx << (2 * i);
    }
    float best_error = FLT_MAX;
    uint2 block = make_uint2(0u, 0u);
    if (do3 && err3 < best_error) {                                                // range.rs:133
        best_error = err3;
        block = write3_packed(a565, b565, (idx3 & act2) | inact);
    }
    if (do4 && err4 < best_error) block = write4_packed(a565, b565, (idx4 & act2) | inact);
    return block;
}

// ---- kernel ---------------------------------------------------------------------------------------------------
#ifndef TXP_RANGE_MIN_CTAS
#define TXP_RANGE_MIN_CTAS 6         // 6 CTAs x 33 KB of shared memory, 24 warps per SM
#endif
template <int FMT>
__global__ void __launch_bounds__(ROLL_THREADS, TXP_RANGE_MIN_CTAS) range_encode_kernel(const BlockSource src, const EncodeParams prm, uint8_t* __restrict__ out) {
    __shared__ float lut[256];
    __shared__ float4 s_col[16][ROLL_THREADS];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = fdiv((float)i, 255.0f);   // colourset.rs:65-67
    __syncthreads();
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= src.nblocks) return;
    uint32_t px[16];
    uint32_t mask;
    load_block_thread(src, b, px, mask);
    uint2 alpha_half = make_uint2(0u, 0u);
    if (FMT == BC2) alpha_half = alpha_bc2_thread(px, mask);
    if (FMT == BC3) {
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = px[i] >> 24;
        alpha_half = mask == 0xFFFFu ? alpha_fit_full(v) : alpha_fit_thread(v, mask);
    }
    const uint2 colour = range_colour_rolled<FMT == BC1>(px, mask, prm, lut, &s_col[0][threadIdx.x]);
    if (FMT == BC1) reinterpret_cast<uint2*>(out)[b] = colour;
    else reinterpret_cast<uint4*>(out)[b] = make_uint4(alpha_half.x, alpha_half.y, colour.x, colour.y);
}

}  // namespace txp
