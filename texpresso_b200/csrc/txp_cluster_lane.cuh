// txp_cluster_lane.cuh -- ClusterFit partition search ("K2L"), one LANE per 4x4 block.
//
// Replaces (reference, /root/reference/lib/src): colourfit/cluster.rs:152-274 (compress3), :276-417 (compress4) and the
// dispatch colourfit.rs:48-59 for Algorithm::ClusterFit (one ordering per pass).  The ordering itself
// (cluster.rs:78-136) comes from cluster_setup_kernel<FMT, true> as the ordered weighted points.
//
// Why a lane per block (measured reasoning in profiles/README.md): the warp-per-block kernel (txp_colour.cuh) spreads the
// candidates of ONE block over 32 lanes, so every lane evaluates its candidate from scratch out of a range-sum table
// (3 x LDS.128 + index decoding per candidate), the winner has to be found by a warp reduction and evaluated again
// (1 of 31 loop trips), the last trip is 7/32 full, and table construction costs every block ~800 warp-instructions.
// Here every lane walks the reference's own loop nest for its own block:
//   * part0 / part1 / part2 are running sums exactly as in cluster.rs:303-376 (no table, one LDS.128 per candidate),
//   * everything that depends on (i, j) only -- part1*(2/3,4/9)+part0 and part1*(1/3,1/9) -- is computed once per (i, j)
//     instead of once per candidate (same operations, same operands => same bits),
//   * "first candidate in loop order wins" is the natural strict `<` of a sequential loop, no tie keys,
//   * the winner is evaluated a second time once per block (1 of 967), to get its endpoints.
// Lanes of a warp stay converged as long as their blocks have the same number of points, so the setup kernel leaves its
// a permutation of every window of 1024 blocks sorted by (points, punch-through) (cluster_setup_sorted_kernel).
#pragma once
#include "txp_colour.cuh"
#include "txp_cluster_setup.cuh"

namespace txp {

constexpr int LANE_THREADS = 128;    // one block per thread and CTA: the hardware CTA scheduler balances the load
#ifndef TXP_LANE_UNROLL
#define TXP_LANE_UNROLL 1            // candidates per trip of the k loop
#endif
#ifndef TXP_LANE_MIN_CTAS
#define TXP_LANE_MIN_CTAS 6          // 6 CTAs x 4 warps = 24 warps per SM (shared memory: 6 x 34 KB)
#endif

struct LaneBest { float err; uint32_t key; };          // key: bit 15 = 3-colour pass, i << 10 | j << 5 | k
constexpr uint32_t LANE_NONE = 0xFFFFFFFFu;

// compress3 (cluster.rs:152-274), one ordering.  col[m * LANE_THREADS] = points_weights[m]; col[count] is a zero guard.
__device__ __forceinline__ void lane_pass3(const float4* col, const int count, const float4 xsum,
                                           const EncodeParams& prm, LaneBest& best) {
    float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int i = 0; i < count; ++i) {
        // cluster.rs:180-181: part1 starts at points_weights[0] (== 0 + points_weights[0]) with j from 1 when i == 0
        float4 p1 = make_float4(0.f, 0.f, 0.f, 0.f);
        int j = i;
        if (i == 0) { p1 = col[0]; j = 1; }
#pragma unroll 1
        for (; j <= count; ++j) {
            const float err = eval3_parts(p0, p1, xsum, prm.wx, prm.wy, prm.wz, nullptr, false);
            if (err < best.err) { best.err = err; best.key = 0x8000u | ((uint32_t)i << 5) | (uint32_t)j; }       // :223 strict
            p1 = f4add(p1, col[j * LANE_THREADS]);                                    // :233-235 (guard row at j == count)
        }
        p0 = f4add(p0, col[i * LANE_THREADS]);                                        // :239
    }
}

// compress4 (cluster.rs:276-417), one ordering.
__device__ __forceinline__ void lane_pass4(const float4* col, const int count, const float4 xsum,
                                           const EncodeParams& prm, LaneBest& best) {
    const float c13 = 1.0f / 3.0f, c19 = 1.0f / 9.0f, c23 = 2.0f / 3.0f, c49 = 4.0f / 9.0f, c29 = 2.0f / 9.0f;
    const f32x2 k13xy = pk(c13, c13), k23xy = pk(c23, c23);
    const f32x2 nz = prm.negzero2;
#ifdef TXP_LANE_LITERAL_GRID
    const f32x2 gxy = pk(31.0f, 63.0f), grxy = pk(1.0f / 31.0f, 1.0f / 63.0f);
#else
    const f32x2 gxy = prm.grid_xy, grxy = prm.gridrcp_xy;        // from the parameter bank: no UMOVs inside the loop
#endif
    // z/w products with the lane-asymmetric multipliers (1/3, 1/9) and (2/3, 4/9) as two scalar FMULs with immediates: as
    // an FFMA2 they would read three distinct vector register pairs (2/3 rate, profiles/README.md)
#ifdef TXP_LANE_ZW_PACKED
#define TXP_ZW13(v) mul2c(v, pk(c13, c19), nz)
#define TXP_ZW23(v) mul2c(v, pk(c23, c49), nz)
#else
#define TXP_ZW13(v) mul2s(v, pk(c13, c19))
#define TXP_ZW23(v) mul2s(v, pk(c23, c49))
#endif
    const f32x2 xs_xy = pk(xsum.x, xsum.y), xs_zw = pk(xsum.z, xsum.w);
    const f32x2 zero2 = pk(0.f, 0.f);
    f32x2 p0xy = zero2, p0zw = zero2;
#pragma unroll 1
    for (int i = 0; i < count; ++i) {
        f32x2 p1xy = zero2, p1zw = zero2;
#pragma unroll 1
        for (int j = i; j <= count; ++j) {
            // the (i, j)-only halves of alphax_sum / betax_sum (cluster.rs:323-328)
            const f32x2 Axy = add2(mul2c(p1xy, k23xy, nz), p0xy), Azw = add2(TXP_ZW23(p1zw), p0zw);
            const f32x2 Bxy = mul2c(p1xy, k13xy, nz), Bzw = TXP_ZW13(p1zw);
            float p1z_, p1w;
            upk(p1zw, p1z_, p1w);
            // cluster.rs:314-315: part2 starts at points_weights[0] with k from 1 when j == 0
            f32x2 p2xy = zero2, p2zw = zero2;
            int k0 = j;
            if (j == 0) { const float4 v = col[0]; p2xy = pk(v.x, v.y); p2zw = pk(v.z, v.w); k0 = 1; }
            // the loop counter is the candidate's key (i, j, k) itself; next = points_weights[k]
            const uint32_t ij = ((uint32_t)i << 10) | ((uint32_t)j << 5), key_end = ij | (uint32_t)count;
            const float4* next = col + k0 * LANE_THREADS;
TXP_UNROLL(TXP_LANE_UNROLL)
            for (uint32_t key = ij | (uint32_t)k0; key <= key_end; ++key, next += LANE_THREADS) {
#if defined(TXP_LANE_SCALAR_ALL) || defined(TXP_LANE_SCALAR_SUMS)
                // A/B variants: the sums (and with _ALL the whole candidate) as scalar operations
                float s0x, s0y, s0z, s0w, s1x, s1y, s1z, s1w, s2x, s2y, s2z, s2w, Ax, Ay, Az, Aw, Bx, By, Bz, Bw;
                upk(p0xy, s0x, s0y); upk(p0zw, s0z, s0w); upk(p1xy, s1x, s1y); upk(p1zw, s1z, s1w);
                upk(p2xy, s2x, s2y); upk(p2zw, s2z, s2w); upk(Axy, Ax, Ay); upk(Azw, Az, Aw); upk(Bxy, Bx, By); upk(Bzw, Bz, Bw);
                const float p3x = sub(sub(sub(xsum.x, s2x), s1x), s0x), p3y = sub(sub(sub(xsum.y, s2y), s1y), s0y);
                const float p3z = sub(sub(sub(xsum.z, s2z), s1z), s0z), p3w = sub(sub(sub(xsum.w, s2w), s1w), s0w);
                float4 alphax, betax;
                alphax.x = add(mul(s2x, c13), Ax); alphax.y = add(mul(s2y, c13), Ay);
                alphax.z = add(mul(s2z, c13), Az); alphax.w = add(mul(s2w, c19), Aw);
                betax.x = add(Bx, add(mul(s2x, c23), p3x)); betax.y = add(By, add(mul(s2y, c23), p3y));
                betax.z = add(Bz, add(mul(s2z, c23), p3z)); betax.w = add(Bw, add(mul(s2w, c49), p3w));
                const float ab = mul(c29, add(s1w, s2w));
#ifdef TXP_LANE_SCALAR_ALL
                const float err = solve<false>(alphax, betax, ab, prm.wx, prm.wy, prm.wz, nullptr);
#else
                const float err = solve_packed(pk(alphax.x, alphax.y), alphax.z, alphax.w, pk(betax.x, betax.y), betax.z, betax.w, ab,
                                               prm.wx, prm.wy, prm.wz, nz, gxy, grxy);
#endif
#else
                const f32x2 p3xy = sub2(sub2(sub2(xs_xy, p2xy), p1xy), p0xy);        // :320
                const f32x2 p3zw = sub2(sub2(sub2(xs_zw, p2zw), p1zw), p0zw);
                const f32x2 axy = add2(mul2c(p2xy, k13xy, nz), Axy), azw = add2(TXP_ZW13(p2zw), Azw);
                const f32x2 bxy = add2(Bxy, add2(mul2c(p2xy, k23xy, nz), p3xy));
                const f32x2 bzw = add2(Bzw, add2(TXP_ZW23(p2zw), p3zw));
                float az, alpha2, bz, beta2, p2z_, p2w;
                upk(azw, az, alpha2);
                upk(bzw, bz, beta2);
                upk(p2zw, p2z_, p2w);
                const float ab = mul(c29, add(p1w, p2w));                            // :331
                const float err = solve_packed(axy, az, alpha2, bxy, bz, beta2, ab, prm.wx, prm.wy, prm.wz, nz, gxy, grxy);
#endif
                if (err < best.err) { best.err = err; best.key = key; }              // :356 strict
                const float4 v = *next;                                               // :367-369 (guard row at k == count)
                p2xy = add2(p2xy, pk(v.x, v.y)); p2zw = add2(p2zw, pk(v.z, v.w));
            }
            const float4 v = col[j * LANE_THREADS];                                   // :373-375
            p1xy = add2(p1xy, pk(v.x, v.y)); p1zw = add2(p1zw, pk(v.z, v.w));
        }
        const float4 v = col[i * LANE_THREADS];                                       // :379
        p0xy = add2(p0xy, pk(v.x, v.y)); p0zw = add2(p0zw, pk(v.z, v.w));
    }
#undef TXP_ZW13
#undef TXP_ZW23
}

// sum of points_weights[a..b) accumulated from zero, left to right (the value of a running part sum)
__device__ __forceinline__ float4 lane_range_sum(const float4* col, const int a, const int b) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int m = a; m < b; ++m) acc = f4add(acc, col[m * LANE_THREADS]);
    return acc;
}

// ---- the block's points, from its record (layout: txp_cluster_setup.cuh) -----------------------------------------------------------
// tabs: [0..255] c / 255 (colourset.rs:65-67), [256 + n] sqrt(n) for n = 0..16 (colourset.rs:107-109 with pixel-count weights)
constexpr int LANE_TABS = 256 + 17;
__device__ __forceinline__ void lane_tabs_init(float* tabs, const int tid) {
    for (int i = tid; i < LANE_TABS; i += LANE_THREADS) tabs[i] = i < 256 ? fdiv((float)i, 255.0f) : __fsqrt_rn((float)(i - 256));
}

// point q of the set as (x, y, z, weight)
__device__ __forceinline__ float4 lane_point(const uint4* __restrict__ rec, const bool alpha_weighted, const float* __restrict__ tabs, const uint32_t q) {
    const uint32_t k = __ldg(reinterpret_cast<const uint32_t*>(rec + 1) + q);
    const float w = alpha_weighted ? __ldg(reinterpret_cast<const float*>(rec + 5) + q) : tabs[256 + (k >> 24)];
    return make_float4(tabs[k & 255u], tabs[(k >> 8) & 255u], tabs[(k >> 16) & 255u], w);
}

// points_weights for the ordering `ow` (cluster.rs:123-133): colw[m] = (x, y, z, 1) * w of point ow[m]; zero guard at [count].
__device__ __forceinline__ void lane_build_pw(const uint4* __restrict__ rec, const bool alpha_weighted, const float* __restrict__ tabs,
                                              const unsigned long long ow, const int count, float4* colw) {
    for (int m = 0; m < count; ++m) {
        const float4 q = lane_point(rec, alpha_weighted, tabs, (uint32_t)(ow >> (4 * m)) & 15u);
        colw[m * LANE_THREADS] = make_float4(mul(q.x, q.w), mul(q.y, q.w), mul(q.z, q.w), q.w);
    }
    colw[count * LANE_THREADS] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// pixel -> point map (colourset.rs:130-141), 4 bits per pixel
__device__ __forceinline__ uint2 lane_remap(const uint4* __restrict__ rec, const bool alpha_weighted) {
    const uint4 rm = __ldg(rec + (alpha_weighted ? 9 : 5));
    return make_uint2(rm.x, rm.y);
}

struct LaneWinner { bool three; int bi, bj, bk; };
__device__ __forceinline__ LaneWinner lane_decode_key(const uint32_t key) {
    LaneWinner w;
    w.three = (key & 0x8000u) != 0;
    w.bi = (int)((key >> (w.three ? 5 : 10)) & 31u);
    w.bj = (int)((key >> (w.three ? 0 : 5)) & 31u);
    w.bk = w.three ? w.bj : (int)(key & 31u);             // 3-colour: no third cluster (codes 0, 2, 1)
    return w;
}

// the winning candidate once more, for its endpoints
__device__ __forceinline__ void lane_winner_endpoints(const float4* col, const float4 xsum, const LaneWinner& w, const EncodeParams& prm, Solution& sol) {
    const float4 p0 = lane_range_sum(col, 0, w.bi), p1 = lane_range_sum(col, w.bi, w.bj);
    if (w.three) eval3_parts(p0, p1, xsum, prm.wx, prm.wy, prm.wz, &sol, true);
    else eval4_parts(p0, p1, lane_range_sum(col, w.bj, w.bk), xsum, prm.wx, prm.wy, prm.wz, &sol, true);
}

// pack_565 of k*gridrcp is k itself (checked in tests/test_identities.py)
__device__ __forceinline__ uint32_t lane_565(const float k[3]) { return ((uint32_t)k[0] << 11) | ((uint32_t)k[1] << 5) | (uint32_t)k[2]; }

// remap + write3/write4 (cluster.rs:254-269 / :396-412, colourset.rs:130-141, colourblock.rs:55-94)
__device__ __forceinline__ uint2 lane_finish_block(const unsigned long long ow, const int count, const LaneWinner& w, const uint32_t a565,
                                                   const uint32_t b565, const uint2 rm, const uint32_t active16) {
    // unordered[order[m]] = code(m), m ascending (later writes win)
    uint32_t pc = 0;                                      // 2 bits per point
    for (int m = 0; m < count; ++m) {
        const uint32_t q = (uint32_t)(ow >> (4 * m)) & 15u;
        const uint32_t cm = m < w.bi ? 0u : (m < w.bj ? 2u : (m < w.bk ? 3u : 1u));
        pc = (pc & ~(3u << (2 * q))) | (cm << (2 * q));
    }
    uint32_t idx2 = 0;                                    // pixels without a point get index 3
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t p = ((i < 8 ? rm.x : rm.y) >> (4 * (i & 7))) & 15u;
        const uint32_t c = ((active16 >> i) & 1u) ? ((pc >> (2 * p)) & 3u) : 3u;
        idx2 |= c << (2 * i);
    }
    return w.three ? write3_packed(a565, b565, idx2) : write4_packed(a565, b565, idx2);
}

template <int FMT>
__global__ void __launch_bounds__(LANE_THREADS, TXP_LANE_MIN_CTAS) cluster_lane_kernel(const EncodeParams prm,
                                                                                       const uint4* __restrict__ rec,
                                                                                       const uint32_t* __restrict__ perm,
                                                                                       uint8_t* __restrict__ out,
                                                                                       const uint64_t first, const uint32_t n) {
    __shared__ float4 s_pw[17][LANE_THREADS];
    __shared__ float tabs[LANE_TABS];
    const int tid = threadIdx.x;
    lane_tabs_init(tabs, tid);
    __syncthreads();
    float4* col = &s_pw[0][tid];
    // perm is window-sorted by cluster_setup_sorted_kernel: 32 consecutive entries are blocks of (mostly) the same shape
    const uint32_t slot = blockIdx.x * LANE_THREADS + tid;
    if (slot >= n) return;
    const uint32_t lb = __ldg(perm + slot);               // chunk-local block number
    const bool aw = prm.alpha_weighted != 0;
    const uint4* r = rec + (size_t)lb * rec_quads(true, aw);
    const uint4 su = __ldg(r);
    if (!(su.z & SETUP_SEARCH)) return;                   // finished by the setup kernel (0 or 1 points)
    const int count = (int)(su.z & 31u);
    const unsigned long long ow = (unsigned long long)su.x | ((unsigned long long)su.y << 32);
    lane_build_pw(r, aw, tabs, ow, count, col);
    const float4 xsum = lane_range_sum(col, 0, count);    // xsum_wsum (cluster.rs:125-133)

    LaneBest best;
    best.err = FLT_MAX;                                   // cluster.rs:66
    best.key = LANE_NONE;
    if (FMT == BC1) {                                     // colourfit.rs:48-59
        lane_pass3(col, count, xsum, prm, best);
        if (!(su.z & SETUP_TRANSPARENT)) lane_pass4(col, count, xsum, prm, best);
    } else {
        lane_pass4(col, count, xsum, prm, best);
    }

    uint2 block = make_uint2(0u, 0u);                     // best_compressed starts zeroed (cluster.rs:71)
    if (best.key != LANE_NONE) {
        const LaneWinner w = lane_decode_key(best.key);
        Solution sol;
        lane_winner_endpoints(col, xsum, w, prm, sol);
        block = lane_finish_block(ow, count, w, lane_565(sol.ka), lane_565(sol.kb), lane_remap(r, aw), su.z >> 16);
    }
    uint2* out2 = reinterpret_cast<uint2*>(out);
    const uint64_t b = first + lb;
    if (FMT == BC1) out2[b] = block; else out2[2 * b + 1] = block;                    // lib.rs:213
}

// ---------------------------------------------------------------------------------------------------------------------
// IterativeClusterFit (cluster.rs:33, :59: up to 8 orderings per pass, each from the axis between the previous winner's
// endpoints), one LANE per block.  Blocks need different numbers of orderings, so lanes do not own fixed blocks: every
// warp draws 32-block chunks of the window-sorted permutation from a global counter and hands the next block to whichever
// lane has finished its current one.  All lanes of a warp then run the same search (the loop nest of one pass) in lockstep.
// The two passes of BC1 are two launches (THREE = compress3, then compress4 on the blocks without punch-through), with
// best_error / best_compressed carried through `carry`, so that lanes in different passes never share a warp.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef TXP_LANE_ITER_MIN_CTAS
#define TXP_LANE_ITER_MIN_CTAS 5
#endif

// construct_ordering (cluster.rs:78-105) for one thread: stable sort of (i, p_i . axis) for i < count, padding (0, f32::MAX)
__device__ __forceinline__ unsigned long long lane_ordering(const uint4* __restrict__ rec, const float* __restrict__ tabs, const int count, const float ax, const float ay, const float az) {
    int sk[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        int k = 0x7F7FFFFF;                               // f32::MAX, finite
        if (e < count) {
            const uint32_t kq = __ldg(reinterpret_cast<const uint32_t*>(rec + 1) + e);
            const uint32_t bits = __float_as_uint(add(add(mul(tabs[kq & 255u], ax), mul(tabs[(kq >> 8) & 255u], ay)), mul(tabs[(kq >> 16) & 255u], az)));
            // fcmp (cluster.rs:90-97): non-finite values compare Equal to each other and Greater than finite
            if ((bits & 0x7F800000u) == 0x7F800000u) k = 0x7FFFFFFF;
            else k = (bits & 0x80000000u) ? -(int)(bits & 0x7FFFFFFFu) : (int)bits;
        }
        sk[e] = k;
    }
    // rank[e] = e - (earlier entries that are larger) + (later entries that are smaller)
    int rank[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) rank[e] = e;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
#pragma unroll
        for (int f = e + 1; f < 16; ++f) {
            if (sk[f] < sk[e]) { ++rank[e]; --rank[f]; }  // strict: on ties the earlier entry stays first
        }
    }
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        if (e < count) {                                  // padding entries carry index 0
            if (rank[e] < 8) lo |= (uint32_t)e << (4 * rank[e]); else hi |= (uint32_t)e << (4 * (rank[e] - 8));
        }
    }
    return (unsigned long long)lo | ((unsigned long long)hi << 32);
}

template <int FMT, bool THREE>
__global__ void __launch_bounds__(LANE_THREADS, TXP_LANE_ITER_MIN_CTAS) cluster_lane_iter_kernel(const EncodeParams prm,
                                                                                                 const uint4* __restrict__ rec,
                                                                                                 const uint32_t* __restrict__ perm,
                                                                                                 uint32_t* __restrict__ carry,
                                                                                                 uint8_t* __restrict__ out,
                                                                                                 uint32_t* __restrict__ counter,
                                                                                                 const uint64_t first, const uint32_t n) {
    __shared__ float4 s_pw[17][LANE_THREADS];
    __shared__ unsigned long long s_seen[8][LANE_THREADS];
    __shared__ float tabs[LANE_TABS];
    const int tid = threadIdx.x, lane = tid & 31;
    lane_tabs_init(tabs, tid);
    __syncthreads();
    const bool aw = prm.alpha_weighted != 0;
    const int rq = rec_quads(true, aw);
    const uint4* r = rec;                                 // record of the block in progress
    float4* col = &s_pw[0][tid];
    uint2* out2 = reinterpret_cast<uint2*>(out);

    uint32_t pool = 0, pool_end = 0;                      // warp-uniform: the warp's current chunk of the permutation
    bool exhausted = false;
    // per-lane state of the block in progress
    bool have = false;
    uint32_t lb = 0, zflags = 0, best_key = LANE_NONE, a565 = 0, b565 = 0;
    int count = 0, it = 0, best_it = 0;
    float start_best = FLT_MAX, run_best = FLT_MAX;
    float bsx = 0.f, bsy = 0.f, bsz = 0.f, bex = 0.f, bey = 0.f, bez = 0.f;
    unsigned long long ow = 0, best_ow = 0;
    float4 xsum = make_float4(0.f, 0.f, 0.f, 0.f);

    for (;;) {
        // ---- give every idle lane its next block -------------------------------------------------------------------
        for (;;) {
            const uint32_t need = __ballot_sync(FULL, !have);
            if (need == 0) break;
            if (pool == pool_end) {
                if (exhausted) break;
                uint32_t c = 0;
                if (lane == 0) c = atomicAdd(counter, 1u);
                c = __shfl_sync(FULL, c, 0);
                if ((uint64_t)c * 32 >= n) { exhausted = true; break; }
                pool = c * 32;
                pool_end = min(pool + 32u, n);
            }
            const uint32_t avail = pool_end - pool, idx = (uint32_t)__popc(need & ((1u << lane) - 1u));
            if (!have && idx < avail) {
                lb = __ldg(perm + pool + idx);
                const uint4 su = __ldg(rec + (size_t)lb * rq);
                bool ok = (su.z & SETUP_SEARCH) != 0;
                if (FMT == BC1 && !THREE && (su.z & SETUP_TRANSPARENT)) ok = false;      // colourfit.rs:51
                if (ok) {
                    zflags = su.z;
                    count = (int)(su.z & 31u);
                    ow = (unsigned long long)su.x | ((unsigned long long)su.y << 32);  // iteration 0: the principal axis
                    start_best = (FMT == BC1 && !THREE) ? __uint_as_float(__ldg(carry + lb)) : FLT_MAX;    // self.best_error
                    run_best = start_best;
                    it = 0; best_it = 0; best_key = LANE_NONE;
                    bsx = bsy = bsz = bex = bey = bez = 0.f;                            // best_start = best_end = zero (:165-166 / :290-291)
                    s_seen[0][tid] = ow;
                    r = rec + (size_t)lb * rq;
                    lane_build_pw(r, aw, tabs, ow, count, col);
                    xsum = lane_range_sum(col, 0, count);
                    have = true;
                }
            }
            pool += min((uint32_t)__popc(need), avail);
        }
        if (!__any_sync(FULL, have)) break;

        // ---- one search over the current ordering, all lanes in lockstep ---------------------------------------------
        LaneBest best;
        best.err = run_best;
        best.key = LANE_NONE;
        if (have) {
            if (THREE) lane_pass3(col, count, xsum, prm, best);
            else lane_pass4(col, count, xsum, prm, best);
        }

        // ---- bookkeeping of the reference's iteration loop (cluster.rs:171-249 / :298-389) ---------------------------
        if (have) {
            if (best.key != LANE_NONE) {                  // strictly better than everything before
                run_best = best.err;
                best_it = it; best_key = best.key; best_ow = ow;
                Solution sol;
                lane_winner_endpoints(col, xsum, lane_decode_key(best.key), prm, sol);
                bsx = sol.ax; bsy = sol.ay; bsz = sol.az; bex = sol.bx; bey = sol.by; bez = sol.bz;
                a565 = lane_565(sol.ka); b565 = lane_565(sol.kb);
            }
            bool finished = best_it != it;                // :243 / :383 (incl. quirk Q9: best_iteration starts at 0)
            if (!finished) {
                ++it;
                if (it == 8) {
                    finished = true;
                } else {
                    ow = lane_ordering(r, tabs, count, sub(bex, bsx), sub(bey, bsy), sub(bez, bsz));    // :248 / :388
                    for (int p = 0; p < it; ++p) finished |= (s_seen[p][tid] == ow);    // :108-120
                    if (!finished) {
                        s_seen[it][tid] = ow;
                        lane_build_pw(r, aw, tabs, ow, count, col);
                        xsum = lane_range_sum(col, 0, count);
                    }
                }
            }
            if (finished) {
                const uint64_t b = first + lb;
                uint2* dst = FMT == BC1 ? out2 + b : out2 + 2 * b + 1;                   // lib.rs:213
                uint2 block = make_uint2(0u, 0u);          // best_compressed starts zeroed (cluster.rs:71)
                const bool improved = run_best < start_best;                             // :252 / :392
                if (improved)
                    block = lane_finish_block(best_ow, count, lane_decode_key(best_key), a565, b565, lane_remap(r, aw), zflags >> 16);
                if (improved || THREE || FMT != BC1) *dst = block;                       // compress4 of BC1 keeps compress3's block otherwise
                if (FMT == BC1 && THREE) carry[lb] = __float_as_uint(run_best);
                have = false;
            }
        }
    }
}

}  // namespace txp
