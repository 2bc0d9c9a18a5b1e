// txp_colour.cuh -- BC1 / BC2 / BC3 encoder kernel: one warp per 4x4 block.
//
// Replaces (reference, /root/reference/lib/src): lib.rs:188-234 (block dispatch), colourset.rs:35-141,
// colourfit.rs:48-59, colourfit/cluster.rs:49-417, colourfit/range.rs:44-192, colourfit/single.rs:58-164,
// colourblock.rs:28-94, math.rs:44-102 and, for the BC2/BC3 alpha half, alpha.rs:27-51 / :187-256.
//
// Work decomposition (B200-first, not the reference's nested scalar loops):
//   * lanes 0..15 own the 16 pixels; the minimal colour set is built with MATCH.ANY + ballots.
//   * sequential fp32 reductions whose order is part of the numeric contract (centroid, covariance,
//     range sums) are kept in the reference's left-to-right order, one lane per vector component.
//   * the ordered partition search is evaluated from a table of range sums S[a][b] held in shared
//     memory.  Every value the reference's running `part0/part1/part2` can take is S[0][i], S[i][j],
//     S[j][k] (left-to-right sums from the range start), so candidates can be evaluated in
//     any order: the (i,j,k) simplex is flattened into one list shared by all colour counts and spread
//     over the 32 lanes, followed by a REDUX.MIN argmin with the reference's "first in loop order wins"
//     tie rule.
#pragma once
#include <cfloat>
#include "txp_common.cuh"

#ifndef TXP_SEARCH_UNROLL
#define TXP_SEARCH_UNROLL 1          // candidates per lane and loop trip in the partition search (interleaved A/B: profiles/README.md)
#endif
#define TXP_PRAGMA_(x) _Pragma(#x)
#define TXP_UNROLL(n) TXP_PRAGMA_(unroll n)

namespace txp {

#ifndef TXP_COLOUR_WARPS
#define TXP_COLOUR_WARPS 4          // 4 warps per CTA: same speed on uniform work, 5-10 % faster when per-block work varies
#endif
constexpr int COLOUR_WARPS = TXP_COLOUR_WARPS;       // warps (= blocks) per CTA
constexpr int TAB4_N = 967;                          // 4-colour candidates at 16 points
constexpr int TAB3_N = 151;                          // 3-colour candidates at 16 points
constexpr int TAB4_PAD = 968, TAB3_PAD = 152;

// Universal candidate lists (filled by the host once per device, see txp_api.cu: build_tables()).
//  g_tab4: k-major, then j, then i  -> for `count` points the valid candidates are the first
//          (count+1)(count+2)(count+3)/6 - 2 entries.   entry = i<<18 | (17i+j)<<9 | (17j+k)
//  g_tab3: j-major, then i          -> first count(count+3)/2 - 1 entries.   entry = i<<9 | (17i+j)
// The entry doubles as the tie-break key: it orders candidates like the reference's loop nest.
__device__ uint32_t g_tab4[TAB4_PAD];
__device__ uint32_t g_tab3[TAB3_PAD];
__constant__ uint8_t c_single_lut[6144];             // single_lut.rs data, layout in single_lut_data.h

struct __align__(16) WarpScratch {
    float4 S[17 * 17];       // S[a*17+b] = sum of ordered weighted points a..b-1 (left to right)
    float4 PW[16];           // ordered weighted points  (cluster.rs:126-133)
    float4 UW[16];           // weighted points in set order
    float4 PT[16];           // points (x,y,z,weight) in set order
    float prod[16][8];       // covariance products
    int keys[16];
    unsigned long long seen[8];
};

constexpr size_t COLOUR_SMEM = COLOUR_WARPS * sizeof(WarpScratch);

__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(add(a.x, b.x), add(a.y, b.y), add(a.z, b.z), add(a.w, b.w)); }
__device__ __forceinline__ float4 f4sub(float4 a, float4 b) { return make_float4(sub(a.x, b.x), sub(a.y, b.y), sub(a.z, b.z), sub(a.w, b.w)); }

struct Solution { float ax, ay, az, bx, by, bz; float ka[3], kb[3]; float err; };

// Least-squares endpoints for one partition + its error: cluster.rs:201-220 == :334-353.
// alphax/betax carry alpha2_sum / beta2_sum in .w.
template <bool WANT_ENDPOINTS>
__device__ __forceinline__ float solve(const float4 alphax, const float4 betax, const float ab,
                                       const float wx, const float wy, const float wz, Solution* out) {
    const float alpha2 = alphax.w, beta2 = betax.w;
    const float factor = rcp_normal(sub(mul(alpha2, beta2), mul(ab, ab)));
    const float av[3] = {alphax.x, alphax.y, alphax.z}, bv[3] = {betax.x, betax.y, betax.z};
    const float grid[3] = {31.0f, 63.0f, 31.0f};
    const float gridrcp[3] = {1.0f / 31.0f, 1.0f / 63.0f, 1.0f / 31.0f};
    const float mw[3] = {wx, wy, wz};
    float e5[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float a = mul(sub(mul(av[c], beta2), mul(bv[c], ab)), factor);
        float b = mul(sub(mul(bv[c], alpha2), mul(av[c], ab)), factor);
        const float ka = grid_index(grid[c], clamp01(a));
        const float kb = grid_index(grid[c], clamp01(b));
        a = mul(ka, gridrcp[c]);
        b = mul(kb, gridrcp[c]);
        const float e1 = add(mul(mul(a, a), alpha2), mul(mul(b, b), beta2));
        const float e2 = sub(mul(mul(a, b), ab), mul(a, av[c]));
        const float e3 = sub(e2, mul(b, bv[c]));
        const float e4 = __fmaf_rn(2.0f, e3, e1);      // 2*e3 is exact, so this equals (2*e3)+e1 bit for bit
        e5[c] = mul(e4, mw[c]);
        if (WANT_ENDPOINTS) {
            out->ka[c] = ka; out->kb[c] = kb;
            if (c == 0) { out->ax = a; out->bx = b; } else if (c == 1) { out->ay = a; out->by = b; } else { out->az = a; out->bz = b; }
        }
    }
    return add(add(e5[0], e5[1]), e5[2]);
}

// One 3-colour candidate from its cluster sums: p0 = part0, p1 = part1 (cluster.rs:189-220).
__device__ __forceinline__ float eval3_parts(const float4 p0, const float4 p1, const float4 xsum,
                                             const float wx, const float wy, const float wz, Solution* out, bool want) {
    const float4 p2 = f4sub(f4sub(xsum, p1), p0);                                   // cluster.rs:189
    const float4 p1h = make_float4(mul(p1.x, 0.5f), mul(p1.y, 0.5f), mul(p1.z, 0.5f), mul(p1.w, 0.25f));
    const float4 alphax = f4add(p1h, p0);                                            // :192
    const float4 betax = f4add(p1h, p2);                                             // :195
    return want ? solve<true>(alphax, betax, p1h.w, wx, wy, wz, out) : solve<false>(alphax, betax, p1h.w, wx, wy, wz, out);
}

__device__ __forceinline__ float eval3(const float4* S, const uint32_t E, const float4 xsum,
                                       const float wx, const float wy, const float wz, Solution* out, bool want) {
    return eval3_parts(S[E >> 9] /* S[0][i] */, S[E & 511u] /* S[i][j] */, xsum, wx, wy, wz, out, want);
}

// One 4-colour candidate from its cluster sums p0, p1, p2 = part0, part1, part2 (cluster.rs:320-353).
__device__ __forceinline__ float eval4_parts(const float4 p0, const float4 p1, const float4 p2, const float4 xsum,
                                             const float wx, const float wy, const float wz, Solution* out, bool want) {
    const float c13 = 1.0f / 3.0f, c19 = 1.0f / 9.0f, c23 = 2.0f / 3.0f, c49 = 4.0f / 9.0f, c29 = 2.0f / 9.0f;
    const float4 p3 = f4sub(f4sub(f4sub(xsum, p2), p1), p0);                         // cluster.rs:320
    float4 alphax, betax;                                                            // :323-328
    alphax.x = add(mul(p2.x, c13), add(mul(p1.x, c23), p0.x));
    alphax.y = add(mul(p2.y, c13), add(mul(p1.y, c23), p0.y));
    alphax.z = add(mul(p2.z, c13), add(mul(p1.z, c23), p0.z));
    alphax.w = add(mul(p2.w, c19), add(mul(p1.w, c49), p0.w));
    betax.x = add(mul(p1.x, c13), add(mul(p2.x, c23), p3.x));
    betax.y = add(mul(p1.y, c13), add(mul(p2.y, c23), p3.y));
    betax.z = add(mul(p1.z, c13), add(mul(p2.z, c23), p3.z));
    betax.w = add(mul(p1.w, c19), add(mul(p2.w, c49), p3.w));
    const float ab = mul(c29, add(p1.w, p2.w));                                      // :331
    return want ? solve<true>(alphax, betax, ab, wx, wy, wz, out) : solve<false>(alphax, betax, ab, wx, wy, wz, out);
}

__device__ __forceinline__ float eval4(const float4* S, const uint32_t E, const float4 xsum,
                                       const float wx, const float wy, const float wz, Solution* out, bool want) {
    return eval4_parts(S[E >> 18] /* S[0][i] */, S[(E >> 9) & 511u] /* S[i][j] */, S[E & 511u] /* S[j][k] */, xsum, wx, wy, wz, out, want);
}

// Second half of a 4-colour candidate with the x/y lanes carried as fp32x2 pairs: endpoints, grid snap and error
// (cluster.rs:334-353) from alphax = (axy, az, alpha2), betax = (bxy, bz, beta2) and ab = alphabeta_sum.
__device__ __forceinline__ float solve_packed(const f32x2 axy, const float az, const float alpha2, const f32x2 bxy, const float bz,
                                              const float beta2, const float ab, const float wx, const float wy, const float wz,
                                              const f32x2 nz, const f32x2 gxy, const f32x2 grxy) {
    const float factor = rcp_normal(sub(mul(alpha2, beta2), mul(ab, ab)));           // :334-335
    float nax, nay, nbx, nby;                                                        // :336-337
    upk(sub2(mul2s(axy, beta2), mul2s(bxy, ab)), nax, nay);
    upk(sub2(mul2s(bxy, alpha2), mul2s(axy, ab)), nbx, nby);
    const float cax = clamp01(mul(nax, factor)), cay = clamp01(mul(nay, factor));    // :340-341
    const float cbx = clamp01(mul(nbx, factor)), cby = clamp01(mul(nby, factor));
    const float caz = clamp01(mul(sub(mul(az, beta2), mul(bz, ab)), factor));
    const float cbz = clamp01(mul(sub(mul(bz, alpha2), mul(az, ab)), factor));
    // grid snap :342-343  (x,y) pairs use (31,63); the z values of a and b share one pair
    const f32x2 hh = pk(0.5f, 0.5f);
    const f32x2 g31 = pk(31.0f, 31.0f), gr31 = pk(1.0f / 31.0f, 1.0f / 31.0f);
    float t0, t1;
    upk(add2(mul2c(pk(cax, cay), gxy, nz), hh), t0, t1);
    const f32x2 a2 = mul2m(pk(truncf(t0), truncf(t1)), grxy);                        // feeds multiplies only
    upk(add2(mul2c(pk(cbx, cby), gxy, nz), hh), t0, t1);
    const f32x2 b2 = mul2m(pk(truncf(t0), truncf(t1)), grxy);
    upk(add2(mul2c(pk(caz, cbz), g31, nz), hh), t0, t1);
    float qaz, qbz;
    upk(mul2m(pk(truncf(t0), truncf(t1)), gr31), qaz, qbz);
    // error terms :346-353, x/y packed
    const f32x2 e1 = add2(mul2s(mul2m(a2, a2), alpha2), mul2s(mul2m(b2, b2), beta2));
    const f32x2 e2 = sub2(mul2s(mul2m(a2, b2), ab), mul2s(a2, axy));
    const f32x2 e3 = sub2(e2, mul2s(b2, bxy));
    const f32x2 e4 = fma2(pk(2.0f, 2.0f), e3, e1);        // 2*e3 exact -> same as (2*e3)+e1
    float e5x, e5y;
    upk(mul2c(e4, pk(wx, wy), nz), e5x, e5y);
    // z scalar
    const float e1z = add(mul(mul(qaz, qaz), alpha2), mul(mul(qbz, qbz), beta2));
    const float e2z = sub(mul(mul(qaz, qbz), ab), mul(qaz, az));
    const float e3z = sub(e2z, mul(qbz, bz));
    const float e5z = mul(__fmaf_rn(2.0f, e3z, e1z), wz);
    return add(add(e5x, e5y), e5z);
}

// eval4 with the x/y lanes (and the z/w lanes of the sums) carried as fp32x2 pairs.  Same operations, same
// association and rounding as eval4 -- only the issue slots are shared (error value is bit-identical).
__device__ __forceinline__ float eval4_packed(const float4* S, const uint32_t E, const f32x2 xs_xy, const f32x2 xs_zw,
                                              const float wx, const float wy, const float wz, const f32x2 nz) {
    const float c13 = 1.0f / 3.0f, c19 = 1.0f / 9.0f, c23 = 2.0f / 3.0f, c49 = 4.0f / 9.0f, c29 = 2.0f / 9.0f;
    const float4 p0 = S[E >> 18];                // S[0][i]
    const float4 p1 = S[(E >> 9) & 511u];        // S[i][j]
    const float4 p2 = S[E & 511u];               // S[j][k]
    const f32x2 p0xy = pk(p0.x, p0.y), p0zw = pk(p0.z, p0.w);
    const f32x2 p1xy = pk(p1.x, p1.y), p1zw = pk(p1.z, p1.w);
    const f32x2 p2xy = pk(p2.x, p2.y), p2zw = pk(p2.z, p2.w);
    const f32x2 p3xy = sub2(sub2(sub2(xs_xy, p2xy), p1xy), p0xy);                    // cluster.rs:320
    const f32x2 p3zw = sub2(sub2(sub2(xs_zw, p2zw), p1zw), p0zw);
    const f32x2 k13xy = pk(c13, c13), k13zw = pk(c13, c19), k23xy = pk(c23, c23), k23zw = pk(c23, c49);
    const f32x2 axy = add2(mul2c(p2xy, k13xy, nz), add2(mul2c(p1xy, k23xy, nz), p0xy)); // :323-324
    const f32x2 azw = add2(mul2c(p2zw, k13zw, nz), add2(mul2c(p1zw, k23zw, nz), p0zw));
    const f32x2 bxy = add2(mul2c(p1xy, k13xy, nz), add2(mul2c(p2xy, k23xy, nz), p3xy)); // :327-328
    const f32x2 bzw = add2(mul2c(p1zw, k13zw, nz), add2(mul2c(p2zw, k23zw, nz), p3zw));
    float az, alpha2, bz, beta2;
    upk(azw, az, alpha2);
    upk(bzw, bz, beta2);
    const float ab = mul(c29, add(p1.w, p2.w));                                      // :331
    return solve_packed(axy, az, alpha2, bxy, bz, beta2, ab, wx, wy, wz, nz, pk(31.0f, 63.0f), pk(1.0f / 31.0f, 1.0f / 63.0f));
}

// The per-block state every fit needs.
struct SetInfo {
    int count;               // distinct colours (uniform)
    bool transparent;        // BC1 punch-through present (uniform)
    int remap;               // lane<16: point index of my pixel, -1 if masked / punched
    float px, py, pz, pw;    // lane<count: point `lane` and its (sqrt'ed) weight
};

// 2-bit index word from per-point codes: pixel l takes the code of point remap[l], 3 if remap<0
// (colourset.rs:130-141).
__device__ __forceinline__ uint32_t pixel_indices(const SetInfo& s, const int point_code, const int lane) {
    const int src = s.remap < 0 ? 0 : s.remap;
    int code = __shfl_sync(FULL, point_code, src);
    if (s.remap < 0) code = 3;
    return __reduce_or_sync(FULL, lane < 16 ? (uint32_t)code << (2 * lane) : 0u);
}

// ---------------------------------------------------------------------------------------------------
// ClusterFit pass (cluster.rs:152-274 for THREE, :276-417 otherwise).  Updates best_error/best_block.
// ---------------------------------------------------------------------------------------------------
// USE_OW0: the ordering of iteration 0 (principal axis) comes from the setup kernel as `ow0`.
// table_ow / table_valid: ordering for which ws->PW / ws->S currently hold the range sums, so that BC1's 3- and
// 4-colour passes (both start from the principal axis, cluster.rs:172 / :299) build them once.
template <bool THREE, bool ITERATE, bool USE_OW0>
__device__ void cluster_pass(const SetInfo& s, const EncodeParams& prm, const float3 principle,
                             const unsigned long long ow0, const bool ow0_degenerate,
                             unsigned long long& table_ow, bool& table_valid,
                             WarpScratch* ws, const uint32_t* tab, const int lane,
                             float& best_error, uint2& best_block) {
    const int count = s.count;
    const int ncand = THREE ? (count * (count + 3)) / 2 - 1 : ((count + 1) * (count + 2) * (count + 3)) / 6 - 2;
    float run_best = best_error;
    int best_iteration = 0;
    uint32_t best_E = 0;
    unsigned long long best_ow = 0;
    int best_rank = 0;
    bool best_degenerate = false;
    float bka[3] = {0, 0, 0}, bkb[3] = {0, 0, 0};
    float bsx = 0, bsy = 0, bsz = 0, bex = 0, bey = 0, bez = 0;
    float axx = principle.x, axy = principle.y, axz = principle.z;

    constexpr int niter = ITERATE ? 8 : 1;                // cluster.rs:33, :59
#pragma unroll 1
    for (int it = 0; it < niter; ++it) {
        // ---- construct_ordering (cluster.rs:78-136) ------------------------------------------------
        unsigned long long ow;
        int rank;
        bool degenerate;
        if (USE_OW0 && it == 0) {
            ow = ow0;
            degenerate = ow0_degenerate;
            // rank of point `lane` = its position in the ordering (inverse permutation through shared memory)
            // (a degenerate ordering repeats entries, SURVEY Q7: several lanes would write one slot -- its ranks are never used, see best_degenerate)
            if (lane < count && !degenerate) ws->keys[(ow >> (4 * lane)) & 15ull] = lane;
            __syncwarp();
            rank = lane < 16 ? ws->keys[lane] : 0;
            __syncwarp();
        } else {
            const float dp = lane < count ? add(add(mul(s.px, axx), mul(s.py, axy)), mul(s.pz, axz)) : FLT_MAX;
            const int idv = lane < count ? lane : 0;
            // fcmp (cluster.rs:90-97): non-finite values compare Equal to each other and Greater than finite.
            const uint32_t bits = __float_as_uint(dp);
            int sk;
            if ((bits & 0x7F800000u) == 0x7F800000u) sk = 0x7FFFFFFF;
            else sk = (bits & 0x80000000u) ? -(int)(bits & 0x7FFFFFFFu) : (int)bits;
            if (lane < 16) ws->keys[lane] = sk;
            __syncwarp();
            rank = 0;                                    // stable rank == insertion sort position
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int kj = ws->keys[j];
                rank += (kj < sk || (kj == sk && j < lane)) ? 1 : 0;
            }
            uint32_t lo = 0, hi = 0;
            if (lane < 16) {
                if (rank < 8) lo = (uint32_t)idv << (4 * rank); else hi = (uint32_t)idv << (4 * (rank - 8));
            }
            lo = __reduce_or_sync(FULL, lo);
            hi = __reduce_or_sync(FULL, hi);
            ow = (unsigned long long)lo | ((unsigned long long)hi << 32);
            degenerate = __any_sync(FULL, lane < count && sk == 0x7FFFFFFF);
        }
        if (ITERATE) {
            bool dup = false;                            // cluster.rs:108-120
            for (int p = 0; p < it; ++p) dup |= (ws->seen[p] == ow);
            if (dup) break;
            __syncwarp();
            if (lane == 0) ws->seen[it] = ow;
        }
        if (!(table_valid && table_ow == ow)) {
            // ordered weighted points: PW[m] = UW[order[m]]  (cluster.rs:126-132)
            __syncwarp();
            if (lane < count) ws->PW[lane] = ws->UW[(ow >> (4 * lane)) & 15ull];
            if (lane <= count) ws->S[lane * 17 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();
            // ---- range-sum table: row `lane` accumulated left to right -------------------------------
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int b = 0; b < count; ++b) {
                const float4 v = ws->PW[b];
                if (lane <= b) {
                    acc = f4add(acc, v);
                    ws->S[lane * 17 + b + 1] = acc;
                }
            }
            __syncwarp();
            table_ow = ow; table_valid = true;
        }
        const float4 xsum = ws->S[count];                // == xsum_wsum (cluster.rs:125-133)
        const f32x2 xs_xy = pk(xsum.x, xsum.y), xs_zw = pk(xsum.z, xsum.w);

        // ---- partition search ----------------------------------------------------------------------
        float lbest = __int_as_float(0x7F800000);
        uint32_t lE = 0xFFFFFFFFu;
        TXP_UNROLL(TXP_SEARCH_UNROLL)
        for (int c = lane; c < ncand; c += 32) {
            const uint32_t E = __ldg(tab + c);
            const float err = THREE ? eval3(ws->S, E, xsum, prm.wx, prm.wy, prm.wz, nullptr, false)
                                    : eval4_packed(ws->S, E, xs_xy, xs_zw, prm.wx, prm.wy, prm.wz, prm.negzero2);
            if (err < lbest || (err == lbest && E < lE)) { lbest = err; lE = E; }
        }
        // lexicographic (error, loop order) argmin over the warp
        const uint32_t o = orderable(add(lbest, 0.0f));
        const uint32_t omin = __reduce_min_sync(FULL, o);
        const uint32_t Ewin = __reduce_min_sync(FULL, o == omin ? lE : 0xFFFFFFFFu);
        const float err_win = __uint_as_float((omin & 0x80000000u) ? (omin & 0x7FFFFFFFu) : ~omin);

        if (Ewin != 0xFFFFFFFFu && err_win < run_best) {  // cluster.rs:223 / :356 (strict)
            Solution sol;
            if (THREE) eval3(ws->S, Ewin, xsum, prm.wx, prm.wy, prm.wz, &sol, true);
            else eval4(ws->S, Ewin, xsum, prm.wx, prm.wy, prm.wz, &sol, true);
            run_best = err_win;
            best_iteration = it;
            best_E = Ewin;
            best_ow = ow;
            best_rank = rank;
            best_degenerate = degenerate;
#pragma unroll
            for (int c = 0; c < 3; ++c) { bka[c] = sol.ka[c]; bkb[c] = sol.kb[c]; }
            bsx = sol.ax; bsy = sol.ay; bsz = sol.az; bex = sol.bx; bey = sol.by; bez = sol.bz;
        }
        if (!ITERATE) break;
        if (best_iteration != it) break;                 // cluster.rs:243 / :383 (incl. quirk Q9)
        axx = sub(bex, bsx); axy = sub(bey, bsy); axz = sub(bez, bsz);      // :248 / :388
        __syncwarp();
    }

    if (run_best < best_error) {                         // cluster.rs:252 / :392
        int bi, bj, bk;
        if (THREE) { bi = (int)(best_E >> 9); bj = (int)(best_E & 511u) - 17 * bi; bk = count; }
        else { bi = (int)(best_E >> 18); bj = (int)((best_E >> 9) & 511u) - 17 * bi; bk = (int)(best_E & 511u) - 17 * bj; }
        // unordered[order[m]] = code(m) (cluster.rs:254-262 / :396-405).  With finite keys the ordering is a
        // permutation and point `lane` sits at position best_rank; with NaN/inf keys entries repeat and later m
        // overwrite earlier ones, which needs the sequential form.
        int code = 0;
        if (!best_degenerate) {
            const int m = best_rank;
            if (THREE) code = m < bi ? 0 : (m < bj ? 2 : 1);
            else code = m < bi ? 0 : (m < bj ? 2 : (m < bk ? 3 : 1));
        } else {
            for (int m = 0; m < count; ++m) {
                const int q = (int)((best_ow >> (4 * m)) & 15ull);
                int cm;
                if (THREE) cm = m < bi ? 0 : (m < bj ? 2 : 1);
                else cm = m < bi ? 0 : (m < bj ? 2 : (m < bk ? 3 : 1));
                if (q == lane) code = cm;
            }
        }
        const uint32_t idx2 = pixel_indices(s, code, lane);
        // pack_565 of k*gridrcp is k itself (checked in tests/test_identities.py)
        const uint32_t a = ((uint32_t)bka[0] << 11) | ((uint32_t)bka[1] << 5) | (uint32_t)bka[2];
        const uint32_t b = ((uint32_t)bkb[0] << 11) | ((uint32_t)bkb[1] << 5) | (uint32_t)bkb[2];
        best_block = THREE ? write3_packed(a, b, idx2) : write4_packed(a, b, idx2);
        best_error = run_best;
    }
}

// ---------------------------------------------------------------------------------------------------
// RangeFit (range.rs:44-192)
// ---------------------------------------------------------------------------------------------------
template <bool IS_BC1>
__device__ uint2 range_fit(const SetInfo& s, const EncodeParams& prm, const float3 principle, const int lane) {
    const int count = s.count;
    const float dp = add(add(mul(s.px, principle.x), mul(s.py, principle.y)), mul(s.pz, principle.z));
    int imin = 0, imax = 0;                               // range.rs:69-85 (sequential scan semantics)
    float mn = __shfl_sync(FULL, dp, 0), mx = mn;
    for (int i = 1; i < count; ++i) {
        const float d = __shfl_sync(FULL, dp, i);
        if (d < mn) { imin = i; mn = d; }
        else if (d > mx) { imax = i; mx = d; }
    }
    float sv[3] = {__shfl_sync(FULL, s.px, imin), __shfl_sync(FULL, s.py, imin), __shfl_sync(FULL, s.pz, imin)};
    float ev[3] = {__shfl_sync(FULL, s.px, imax), __shfl_sync(FULL, s.py, imax), __shfl_sync(FULL, s.pz, imax)};
    const float grid[3] = {31.0f, 63.0f, 31.0f};
    const float gridrcp[3] = {1.0f / 31.0f, 1.0f / 63.0f, 1.0f / 31.0f};
    uint32_t ks[3], ke[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {                        // range.rs:88-98
        const float a = grid_index(grid[c], clamp01(sv[c]));
        const float b = grid_index(grid[c], clamp01(ev[c]));
        ks[c] = (uint32_t)a; ke[c] = (uint32_t)b;
        sv[c] = mul(a, gridrcp[c]); ev[c] = mul(b, gridrcp[c]);
    }
    const uint32_t a565 = (ks[0] << 11) | (ks[1] << 5) | ks[2];
    const uint32_t b565 = (ke[0] << 11) | (ke[1] << 5) | ke[2];
    const float p[3] = {s.px, s.py, s.pz};
    const float mw[3] = {prm.wx, prm.wy, prm.wz};

    float best_error = FLT_MAX;
    uint2 block = make_uint2(0u, 0u);
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const bool three = pass == 0;
        if (three && !IS_BC1) continue;                   // colourfit.rs:48-56
        if (!three && IS_BC1 && s.transparent) continue;
        float codes[4][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            codes[0][c] = sv[c]; codes[1][c] = ev[c];
            if (three) {                                  // range.rs:161
                codes[2][c] = add(mul(sv[c], 0.5f), mul(ev[c], 0.5f));
                codes[3][c] = 0.f;
            } else {                                      // range.rs:176-181
                codes[2][c] = add(mul(sv[c], 2.0f / 3.0f), mul(ev[c], 1.0f / 3.0f));
                codes[3][c] = add(mul(sv[c], 1.0f / 3.0f), mul(ev[c], 2.0f / 3.0f));
            }
        }
        float dist = FLT_MAX; int idx = 0;                // range.rs:111-123
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (three && j == 3) continue;
            const float dx = mul(mw[0], sub(p[0], codes[j][0]));
            const float dy = mul(mw[1], sub(p[1], codes[j][1]));
            const float dz = mul(mw[2], sub(p[2], codes[j][2]));
            const float d = add(add(mul(dx, dx), mul(dy, dy)), mul(dz, dz));
            if (d < dist) { dist = d; idx = j; }
        }
        float error = 0.0f;                               // range.rs:129 (sequential)
        for (int i = 0; i < count; ++i) error = add(error, __shfl_sync(FULL, dist, i));
        if (error < best_error) {                         // range.rs:133
            const uint32_t idx2 = pixel_indices(s, idx, lane);
            best_error = error;
            block = three ? write3_packed(a565, b565, idx2) : write4_packed(a565, b565, idx2);
        }
    }
    return block;
}

// ---------------------------------------------------------------------------------------------------
// SingleColourFit (single.rs:58-164).  rgb = the block's only colour, as bytes.
// ---------------------------------------------------------------------------------------------------
struct SingleEnds { uint32_t a565, b565, index, error; };

__device__ __forceinline__ SingleEnds single_endpoints(const uint32_t rgb, const int t0, const int t1, const int t2) {
    const int tabs[3] = {t0, t1, t2};
    const uint32_t col[3] = {rgb & 255u, (rgb >> 8) & 255u, (rgb >> 16) & 255u};
    SingleEnds r; r.a565 = 0; r.b565 = 0; r.index = 0; r.error = 0xFFFFFFFFu;
#pragma unroll
    for (int index = 0; index < 2; ++index) {
        uint32_t error = 0, st[3], en[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint8_t* e = &c_single_lut[((tabs[c] * 256 + col[c]) * 2 + index) * 3];
            st[c] = e[0]; en[c] = e[1];
            error += (uint32_t)e[2] * (uint32_t)e[2];
        }
        if (error < r.error) {                            // single.rs:91 (strict)
            r.a565 = (st[0] << 11) | (st[1] << 5) | st[2];   // pack_565(s/31) == s
            r.b565 = (en[0] << 11) | (en[1] << 5) | en[2];
            r.index = 2u * index;
            r.error = error;
        }
    }
    return r;
}

template <bool IS_BC1>
__device__ uint2 single_fit(const SetInfo& s, const uint32_t rgb, const int lane) {
    const bool active = s.remap >= 0;
    uint2 block = make_uint2(0u, 0u);
    uint32_t best = 0xFFFFFFFFu;
    if (IS_BC1) {
        const SingleEnds r = single_endpoints(rgb, 0, 1, 0);
        const uint32_t idx2 = __reduce_or_sync(FULL, lane < 16 ? (active ? r.index : 3u) << (2 * lane) : 0u);
        block = write3_packed(r.a565, r.b565, idx2);
        best = r.error;
        if (s.transparent) return block;                  // colourfit.rs:51
    }
    const SingleEnds r = single_endpoints(rgb, 2, 3, 2);
    if (r.error < best) {
        const uint32_t idx2 = __reduce_or_sync(FULL, lane < 16 ? (active ? r.index : 3u) << (2 * lane) : 0u);
        block = write4_packed(r.a565, r.b565, idx2);
    }
    return block;
}

// ---------------------------------------------------------------------------------------------------
// Colour half of one block: ColourSet (colourset.rs:35-112) + dispatch (lib.rs:208-231).
// `pix`/`valid` are meaningful for lanes 0..15; result is uniform across the warp.
// ---------------------------------------------------------------------------------------------------
// HAVE_SETUP: the block went through cluster_setup_kernel (txp_cluster_setup.cuh): it has >= 2 points, and the
// principal-axis ordering `ow0` is given, so covariance / power iteration / first sort are skipped here.
template <bool IS_BC1, bool HAVE_SETUP>
__device__ uint2 colour_block(const uint32_t pix, const bool valid, const EncodeParams& prm,
                              WarpScratch* ws, const uint32_t* tab3, const uint32_t* tab4, const int lane,
                              const unsigned long long ow0 = 0, const bool ow0_degenerate = false) {
    const uint32_t rgb = pix & 0x00FFFFFFu, alpha = pix >> 24;
    const bool punched = IS_BC1 && valid && alpha < 128u;                         // colourset.rs:54
    const bool active = valid && !punched;
    SetInfo s;
    s.transparent = IS_BC1 && __any_sync(FULL, punched);
    // exact-RGB duplicates among active pixels (colourset.rs:84-88): one MATCH.ANY
    const uint32_t grp = __match_any_sync(FULL, active ? rgb : (0x01000000u | (uint32_t)lane));
    const int first = __ffs(grp) - 1;
    const bool is_new = active && first == lane;
    const uint32_t newmask = __ballot_sync(FULL, is_new);
    s.count = __popc(newmask);
    s.remap = active ? __popc(newmask & ((1u << first) - 1u)) : -1;

    if (s.count == 0)                                      // lib.rs:223 -> RangeFit on an empty set
        return IS_BC1 ? make_uint2(0u, 0xFFFFFFFFu) : make_uint2(0u, 0u);
    if (s.count == 1)                                      // lib.rs:217-222
        return single_fit<IS_BC1>(s, __shfl_sync(FULL, rgb, __ffs(newmask) - 1), lane);

    // weights: sums of 1 or (alpha+1)/256 are exact in fp32 in any order (multiples of 2^-8 below 2^5)
    uint32_t wsum;
    if (prm.alpha_weighted) {
        wsum = 0;
        for (int j = 0; j < 16; ++j) {
            const uint32_t aj = __shfl_sync(FULL, alpha + 1u, j);
            if ((grp >> j) & 1u) wsum += aj;
        }
    } else {
        wsum = (uint32_t)__popc(grp);
    }
    if (is_new) {
        const float w = prm.alpha_weighted ? fdiv((float)wsum, 256.0f) : (float)wsum;
        ws->PT[s.remap] = make_float4(fdiv((float)(rgb & 255u), 255.0f), fdiv((float)((rgb >> 8) & 255u), 255.0f),
                                      fdiv((float)((rgb >> 16) & 255u), 255.0f), __fsqrt_rn(w));
    }
    __syncwarp();
    const float4 pt = lane < s.count ? ws->PT[lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    s.px = pt.x; s.py = pt.y; s.pz = pt.z; s.pw = pt.w;
    if (lane < s.count) ws->UW[lane] = make_float4(mul(pt.x, pt.w), mul(pt.y, pt.w), mul(pt.z, pt.w), mul(1.0f, pt.w));
    __syncwarp();

    float3 principle = make_float3(0.f, 0.f, 0.f);
    if (!HAVE_SETUP) {
        // ---- Sym3x3::weighted_covariance (math.rs:44-73), sums in set order -----------------------------
        float acc = 0.0f;
        if (lane < 4) {
            const float* uw = reinterpret_cast<const float*>(ws->UW);
            for (int k = 0; k < s.count; ++k) acc = add(acc, uw[4 * k + lane]);
        }
        const float total = __shfl_sync(FULL, acc, 3);
        float cx = __shfl_sync(FULL, acc, 0), cy = __shfl_sync(FULL, acc, 1), cz = __shfl_sync(FULL, acc, 2);
        if (total > FLT_EPSILON) { cx = fdiv(cx, total); cy = fdiv(cy, total); cz = fdiv(cz, total); }
        if (lane < s.count) {
            const float ax = sub(pt.x, cx), ay = sub(pt.y, cy), az = sub(pt.z, cz);
            const float bx = mul(ax, pt.w), by = mul(ay, pt.w), bz = mul(az, pt.w);
            float* pr = ws->prod[lane];
            pr[0] = mul(ax, bx); pr[1] = mul(ax, by); pr[2] = mul(ax, bz);
            pr[3] = mul(ay, by); pr[4] = mul(ay, bz); pr[5] = mul(az, bz);
        }
        __syncwarp();
        acc = 0.0f;
        if (lane < 6) for (int k = 0; k < s.count; ++k) acc = add(acc, ws->prod[k][lane]);
        const float m0 = __shfl_sync(FULL, acc, 0), m1 = __shfl_sync(FULL, acc, 1), m2 = __shfl_sync(FULL, acc, 2);
        const float m3 = __shfl_sync(FULL, acc, 3), m4 = __shfl_sync(FULL, acc, 4), m5 = __shfl_sync(FULL, acc, 5);

        // ---- principle_component (math.rs:75-97): 8 power iterations, uniform ---------------------------
        float vx = 1.0f, vy = 1.0f, vz = 1.0f;
    #pragma unroll 1
        for (int it = 0; it < 8; ++it) {
            const float wx = add(mul(m2, vz), add(mul(m1, vy), mul(m0, vx)));
            const float wy = add(mul(m4, vz), add(mul(m3, vy), mul(m1, vx)));
            const float wz = add(mul(m5, vz), add(mul(m4, vy), mul(m2, vx)));
            const float a = fmaxf(wx, fmaxf(wy, wz));
            const float ra = rcp(a);
            vx = mul(wx, ra); vy = mul(wy, ra); vz = mul(wz, ra);
        }
        principle = make_float3(vx, vy, vz);

        if (prm.algorithm == RANGE_FIT) return range_fit<IS_BC1>(s, prm, principle, lane);

    }

    // ---- ClusterFit (cluster.rs:49-76 + colourfit.rs:48-59) ----------------------------------------
    float best_error = FLT_MAX;
    uint2 block = make_uint2(0u, 0u);
    unsigned long long table_ow = 0;
    bool table_valid = false;
#define TXP_PASS(THREE, ITER, TAB) cluster_pass<THREE, ITER, HAVE_SETUP>(s, prm, principle, ow0, ow0_degenerate, table_ow, table_valid, ws, TAB, lane, best_error, block)
    if (prm.algorithm == ITERATIVE_CLUSTER_FIT) {
        if (IS_BC1) {
            TXP_PASS(true, true, tab3);
            __syncwarp();
            if (!s.transparent) TXP_PASS(false, true, tab4);
        } else {
            TXP_PASS(false, true, tab4);
        }
    } else {
        if (IS_BC1) {
            TXP_PASS(true, false, tab3);
            __syncwarp();
            if (!s.transparent) TXP_PASS(false, false, tab4);
        } else {
            TXP_PASS(false, false, tab4);
        }
    }
#undef TXP_PASS
    return block;
}

}  // namespace txp
