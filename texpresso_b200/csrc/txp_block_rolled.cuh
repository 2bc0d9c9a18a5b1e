// txp_block_rolled.cuh -- the serial per-block front end of every BC1/BC2/BC3 encoder, one THREAD per 4x4 block, written as
// ROLLED loops over a shared-memory copy of the block.
//
// Replaces (reference, /root/reference/lib/src): colourset.rs:35-112 (ColourSet::new), math.rs:44-73 (weighted_covariance),
// math.rs:75-97 (principle_component), colourfit/range.rs:44-192 (RangeFit) and the first half of cluster.rs:78-136
// (construct_ordering on the principal axis).
//
// Why rolled (ncu of the straight-line version, profiles/ncu_range_r02_summary.txt): 4 100 warp-instructions per block laid out
// as ~6 000 instructions of straight-line code, `no_instruction` (instruction-cache misses) the top stall reason at 2.5 stalled
// warps per issue, 128 registers (16 warps per SM), 240 table look-ups per block because the pixel's c/255 values were looked up
// again in every loop to save registers.  Here a thread keeps its block as sixteen float4 (r/255, g/255, b/255, weight) in its
// own column of shared memory, built once (48 look-ups), and every order-sensitive loop of the reference (math.rs:48-70,
// range.rs:67-86, :107-131) is a short rolled loop over that column: one LDS.128 per pixel and loop, a few hundred
// instructions of code in total, < 64 registers.
//
// The colour set is not compacted.  A pixel is "new" if no earlier pixel has its key; pixels that are not new carry weight 0
// and contribute x + (+-0), an exact no-op for accumulators that start at +0, so the sums run over the points of the set in
// the reference's order.  Duplicate detection yields first[i] (the earliest pixel with pixel i's colour) with two instructions
// per pair; group weights (colourset.rs:94-97) are integer sums scattered through the column's w slot.
#pragma once
#include <cfloat>
#include "txp_common.cuh"

namespace txp {

constexpr int ROLL_THREADS = 128;                 // threads per CTA of the kernels that use a column: col[i * ROLL_THREADS]
constexpr size_t ROLL_SMEM = 16 * ROLL_THREADS * sizeof(float4);

struct RolledSet {
    uint32_t active16;       // valid and not punched through
    uint32_t new16;          // first occurrence of its RGB among the active pixels (== the points of the set, in order)
    uint32_t first_lo, first_hi;   // 4 bits per pixel: index of the first pixel with the same key (itself if new / inactive)
    bool transparent;        // BC1 punch-through present (colourset.rs:54-58)
};

// colourset.rs:35-112 up to the weights.  px: the block's RGBA words; on return px[i] holds the comparison key (RGB for
// active pixels, a unique value otherwise).  col[i].w receives the integer weight total of the group whose first pixel is i
// (1 per pixel, or alpha + 1 if alpha-weighted: exact integer sums), as raw bits.
template <bool IS_BC1>
__device__ __forceinline__ RolledSet rolled_colourset(uint32_t px[16], const uint32_t mask, const bool alpha_weighted, float4* col) {
    RolledSet t;
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
        reinterpret_cast<uint32_t*>(col + i * ROLL_THREADS)[3] = 0u;
    }
    t.transparent = punched != 0;
    // exact-RGB duplicates (:84-88): first[i] = smallest j <= i with key[j] == key[i]
    uint32_t f[16];
    f[0] = 0;
#pragma unroll
    for (int i = 1; i < 16; ++i) {
        uint32_t fi = (uint32_t)i;
#pragma unroll
        for (int j = i - 1; j >= 0; --j) fi = px[j] == px[i] ? (uint32_t)j : fi;
        f[i] = fi;
    }
    uint32_t dup16 = 0;
    t.first_lo = 0; t.first_hi = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if (f[i] != (uint32_t)i) dup16 |= 1u << i;
        if (i < 8) t.first_lo |= f[i] << (4 * i); else t.first_hi |= f[i] << (4 * (i - 8));
        // group totals: one read-modify-write of the first pixel's w slot per pixel (a thread only touches its own column)
        uint32_t* slot = reinterpret_cast<uint32_t*>(col + f[i] * ROLL_THREADS) + 3;
        *slot += wgt[i];
    }
    t.new16 = t.active16 & ~dup16;
    return t;
}

// col[i] = (r/255, g/255, b/255, weight_i): the points of the set with their sqrt'ed weights (colourset.rs:65-67, :107-109);
// weight 0 marks a pixel that is not a point.  lut[c] = c / 255 (IEEE division, built by the kernel).
// EMIT: also leaves the points in set order for the lane-per-block search kernels (record layout: txp_cluster_setup.cuh) --
// keys[p] = RGB of point p, with its integer weight total in the top byte when weights are pixel counts (<= 16), or the fp32
// weight in wts[p] when they are alpha sums.
template <bool EMIT>
__device__ __forceinline__ void rolled_fill_points(const uint32_t key[16], const RolledSet& ts, const bool alpha_weighted,
                                                   const float* __restrict__ lut, float4* col, uint32_t* __restrict__ keys = nullptr,
                                                   float* __restrict__ wts = nullptr) {
    int p = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float4* e = col + i * ROLL_THREADS;
        const uint32_t gw = reinterpret_cast<const uint32_t*>(e)[3];
        float w = 0.0f;
        if ((ts.new16 >> i) & 1u) {
            w = 1.0f;
            if (gw != 1u || alpha_weighted)
                w = __fsqrt_rn(alpha_weighted ? mul((float)gw, 1.0f / 256.0f) : (float)gw);
            if (EMIT) {
                if (alpha_weighted) { keys[p] = key[i]; wts[p] = w; } else keys[p] = key[i] | (gw << 24);
                ++p;
            }
        }
        *e = make_float4(lut[key[i] & 255u], lut[(key[i] >> 8) & 255u], lut[(key[i] >> 16) & 255u], w);
    }
}

// Sym3x3::weighted_covariance + principle_component (math.rs:44-97) over the column
__device__ __forceinline__ float3 rolled_principal_axis(const float4* col) {
    float total = 0.0f, cx = 0.0f, cy = 0.0f, cz = 0.0f;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const float4 p = col[i * ROLL_THREADS];
        total = add(total, p.w);
        cx = add(cx, mul(p.x, p.w)); cy = add(cy, mul(p.y, p.w)); cz = add(cz, mul(p.z, p.w));
    }
    if (total > FLT_EPSILON) { cx = fdiv(cx, total); cy = fdiv(cy, total); cz = fdiv(cz, total); }
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f, m5 = 0.f;
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
        const float4 p = col[i * ROLL_THREADS];
        const float ax = sub(p.x, cx), ay = sub(p.y, cy), az = sub(p.z, cz);
        const float bx = mul(ax, p.w), by = mul(ay, p.w), bz = mul(az, p.w);
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
__device__ __forceinline__ uint32_t rolled_single_rgb(const uint32_t key[16], const uint32_t active16) {
    uint32_t rgb = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) if ((active16 >> i) & 1u) rgb |= key[i];
    return rgb;
}

}  // namespace txp
