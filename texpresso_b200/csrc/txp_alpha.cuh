// txp_alpha.cuh -- alpha / single-channel codecs.
//
// Replaces (reference, /root/reference/lib/src/alpha.rs): compress_bc2 :27-51, fix_range :70-77,
// fit_codes :79-119, write_alpha_block{,5,7} :121-185, compress_bc3 :187-256.
//
// Two shapes of the same integer algorithm:
//   * warp_alpha_bc3 / warp_alpha_bc2 : 16 lanes = 16 pixels, used for the alpha half of BC2/BC3 inside
//     the warp-per-block colour kernel (REDUX min/max/add/or instead of loops).
//   * alpha_fit_thread                : one thread per 8-byte output, used by the BC4 / BC5 kernel, which
//     is the HBM-bound path (64 B in, 8/16 B out per block).
//
// Exact identity used throughout: squared distance is monotone in |v - c|, so the
// reference's first-minimum argmin over (v-c)^2 equals the minimum of the key |v-c|*8 + j.
#pragma once
#include "txp_common.cuh"

namespace txp {

// alpha.rs:70-77
__device__ __forceinline__ void fix_range(int& mn, int& mx, const int steps) {
    if (mx - mn < steps) mx = min(mn + steps, 255);
    if (mx - mn < steps) mn = max(mx - steps, 0);
}

// codebooks of alpha.rs:226-242 (including the min5/max5 quirk of the 7-point book)
__device__ __forceinline__ void alpha_codebooks(const int min5, const int max5, const int min7, const int max7,
                                                int codes5[8], int codes7[8]) {
    codes5[0] = min5; codes5[1] = max5;
#pragma unroll
    for (int i = 1; i < 5; ++i) codes5[1 + i] = ((5 - i) * min5 + i * max5) / 5;
    codes5[6] = 0; codes5[7] = 255;
    codes7[0] = min5; codes7[1] = max5;
#pragma unroll
    for (int i = 1; i < 7; ++i) codes7[1 + i] = ((7 - i) * min7 + i * max7) / 7;
}

// nearest code, first minimum wins: returns |v-c|*8 + j minimised over j (alpha.rs:101-111)
__device__ __forceinline__ uint32_t nearest_key(const uint32_t v8, const int codes[8]) {
    uint32_t k[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) k[j] = __usad(v8, (uint32_t)codes[j] << 3, (uint32_t)j);
    const uint32_t a = __vimin3_u32(k[0], k[1], k[2]);
    const uint32_t b = __vimin3_u32(k[3], k[4], k[5]);
    return __vimin3_u32(a, b, min(k[6], k[7]));
}

// index remaps of write_alpha_block5 / write_alpha_block7 (alpha.rs:146-185) for a swapped pair
__device__ __forceinline__ uint32_t swap5(uint32_t x) { return x == 0 ? 1u : x == 1 ? 0u : (x <= 5 ? 7u - x : x); }
__device__ __forceinline__ uint32_t swap7(uint32_t x) { return x == 0 ? 1u : x == 1 ? 0u : 9u - x; }

// ---- warp-cooperative 5-/7-point fit (alpha.rs:187-256).  value/valid meaningful for lanes 0..15 ----
__device__ __forceinline__ uint2 warp_alpha_bc3(const uint32_t value, const bool valid, const int lane) {
    const bool v = valid && lane < 16;
    int min7 = (int)__reduce_min_sync(FULL, v ? value : 255u);
    int max7 = (int)__reduce_max_sync(FULL, v ? value : 0u);
    int min5 = (int)__reduce_min_sync(FULL, (v && value != 0u) ? value : 255u);
    int max5 = (int)__reduce_max_sync(FULL, (v && value != 255u) ? value : 0u);
    if (min5 > max5) min5 = max5;
    if (min7 > max7) min7 = max7;
    fix_range(min5, max5, 5);
    fix_range(min7, max7, 7);
    int codes5[8], codes7[8];
    alpha_codebooks(min5, max5, min7, max7, codes5, codes7);
    const uint32_t k5 = nearest_key(value << 3, codes5), k7 = nearest_key(value << 3, codes7);
    const uint32_t d5 = k5 >> 3, d7 = k7 >> 3;
    const uint32_t err5 = __reduce_add_sync(FULL, v ? d5 * d5 : 0u);
    const uint32_t err7 = __reduce_add_sync(FULL, v ? d7 * d7 : 0u);
    uint32_t a0, a1, idx;
    if (err5 <= err7) {                                   // alpha.rs:251
        idx = v ? (k5 & 7u) : 0u;
        a0 = (uint32_t)min5; a1 = (uint32_t)max5;
        if (a0 > a1) { idx = swap5(idx); const uint32_t t = a0; a0 = a1; a1 = t; }     // dead after fix_range (Q3b)
    } else {
        idx = v ? (k7 & 7u) : 0u;
        a0 = (uint32_t)min7; a1 = (uint32_t)max7;
        if (a0 < a1) { idx = swap7(idx); const uint32_t t = a0; a0 = a1; a1 = t; }     // always taken (Q3b)
    }
    // alpha.rs:121-144: two groups of 8 x 3 bits
    const uint32_t g0 = __reduce_or_sync(FULL, lane < 8 ? idx << (3 * lane) : 0u);
    const uint32_t g1 = __reduce_or_sync(FULL, (lane >= 8 && lane < 16) ? idx << (3 * (lane - 8)) : 0u);
    return make_uint2(a0 | (a1 << 8) | (g0 << 16), (g0 >> 16) | (g1 << 8));
}

// ---- warp-cooperative BC2 alpha (alpha.rs:27-51) --------------------------------------------------
__device__ __forceinline__ uint2 warp_alpha_bc2(const uint32_t value, const bool valid, const int lane) {
    const float alpha = mul((float)value, 15.0f / 255.0f);
    uint32_t q = (uint32_t)f32_to_i32_clamped(alpha, 15);
    if (!valid) q = 0;
    const uint32_t lo = __reduce_or_sync(FULL, lane < 8 ? q << (4 * lane) : 0u);
    const uint32_t hi = __reduce_or_sync(FULL, (lane >= 8 && lane < 16) ? q << (4 * (lane - 8)) : 0u);
    return make_uint2(lo, hi);
}

// ---- one thread, one channel: 16 values (v[i], i = 4*py+px) and the 16-bit valid mask --------------
__device__ __forceinline__ uint2 alpha_fit_thread(const uint32_t v[16], const uint32_t mask) {
    int min5 = 255, max5 = 0, min7 = 255, max7 = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if ((mask >> i) & 1u) {
            const int x = (int)v[i];
            min7 = min(min7, x); max7 = max(max7, x);
            if (x != 0) min5 = min(min5, x);
            if (x != 255) max5 = max(max5, x);
        }
    }
    if (min5 > max5) min5 = max5;
    if (min7 > max7) min7 = max7;
    fix_range(min5, max5, 5);
    fix_range(min7, max7, 7);
    int codes5[8], codes7[8];
    alpha_codebooks(min5, max5, min7, max7, codes5, codes7);
    uint32_t err5 = 0, err7 = 0;
    unsigned long long i5 = 0, i7 = 0;                   // 16 x 3-bit indices
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        if ((mask >> i) & 1u) {
            const uint32_t k5 = nearest_key(v[i] << 3, codes5), k7 = nearest_key(v[i] << 3, codes7);
            err5 += (k5 >> 3) * (k5 >> 3);
            err7 += (k7 >> 3) * (k7 >> 3);
            i5 |= (unsigned long long)(k5 & 7u) << (3 * i);
            i7 |= (unsigned long long)swap7(k7 & 7u) << (3 * i);   // pre-swapped; see below
        } else {
            i7 |= 1ull << (3 * i);                        // masked pixels hold index 0, which swaps to 1
        }
    }
    uint32_t a0, a1; unsigned long long idx;
    if (err5 <= err7) {
        a0 = (uint32_t)min5; a1 = (uint32_t)max5; idx = i5;    // min5 <= max5 always: no swap (Q3b)
    } else {
        // after fix_range max7 - min7 >= 7, so write_alpha_block7 always swaps (Q3b)
        a0 = (uint32_t)max7; a1 = (uint32_t)min7; idx = i7;
    }
    const uint32_t g0 = (uint32_t)(idx & 0xFFFFFFull), g1 = (uint32_t)(idx >> 24);
    return make_uint2(a0 | (a1 << 8) | (g0 << 16), (g0 >> 16) | (g1 << 8));
}

// ---- fast path for fully valid blocks (mask == 0xFFFF) ------------------------------------------------------
// The nearest-code search is moved from the ALU pipe (|v-c| keys) to the FMA pipe:
//   (v-c)^2 = v^2 - 2vc + c^2, so  argmin_j ((v-c_j)^2, j)  ==  argmin_j  key_j,
//   key_j = 8*(c_j^2 - 2*v*c_j) + j = v*(-16*c_j) + (8*c_j^2 + j)        -- one IMAD per code and pixel,
// the same lexicographic (distance, first index) rule as alpha.rs:101-111.  key_min & 7 is the index and
// sum(key_min - j_min) = 8*(err - sum v^2), so err5 <= err7 is decided on the key sums directly.
struct KeyBook { int a[8], b[8]; };

__device__ __forceinline__ void make_keybook(const int codes[8], KeyBook& kb) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { kb.a[j] = -16 * codes[j]; kb.b[j] = 8 * codes[j] * codes[j] + j; }
}

// interpolated codes without integer division: ((N-i)*lo + i*hi)/N == lo + floor(i*(hi-lo)/N), and for
// 0 <= x <= 1530 floor(x/5) == (x*205)>>10 (x <= 1020) and floor(x/7) == (x*9363)>>16  (checked exhaustively
// in tests/test_identities.py)
__device__ __forceinline__ void alpha_codebooks_fast(const int min5, const int max5, const int min7, const int max7,
                                                     int codes5[8], int codes7[8]) {
    const int r5 = max5 - min5, r7 = max7 - min7;
    codes5[0] = min5; codes5[1] = max5;
#pragma unroll
    for (int i = 1; i < 5; ++i) codes5[1 + i] = min5 + ((i * 205 * r5) >> 10);
    codes5[6] = 0; codes5[7] = 255;
    codes7[0] = min5; codes7[1] = max5;
#pragma unroll
    for (int i = 1; i < 7; ++i) codes7[1 + i] = min7 + ((i * 9363 * r7) >> 16);
}

// 24-bit words of eight 3-bit fields; EVEN3 selects fields 0,2,4,6 (6-bit lanes at bits 0,6,12,18)
constexpr uint32_t EVEN3 = 0x1C71C7u;

// sum of the sixteen 3-bit fields held in two 24-bit words
__device__ __forceinline__ int sum_fields3(const uint32_t lo, const uint32_t hi) {
    const uint32_t t = (lo & EVEN3) + ((lo >> 3) & EVEN3) + (hi & EVEN3) + ((hi >> 3) & EVEN3);   // 4 lanes, each <= 28
    const uint32_t u = (t & 0x03F03Fu) + ((t >> 6) & 0x03F03Fu);                                  // 2 lanes (bits 0, 12), each <= 56
    return (int)((u & 0xFFFu) + (u >> 12));
}

// write_alpha_block7's index swap (alpha.rs:172-178) on all eight fields at once:
// 0->1, 1->0, x->9-x  is  f -> (9 - f) & 7   (9-0 = 9 -> 1, 9-1 = 8 -> 0)
__device__ __forceinline__ uint32_t swap7_fields(const uint32_t w) {
    const uint32_t nine = 0x249249u;                       // 9 in every 6-bit lane
    const uint32_t e = (nine - (w & EVEN3)) & EVEN3;
    const uint32_t o = (nine - ((w >> 3) & EVEN3)) & EVEN3;
    return e | (o << 3);
}

__device__ __forceinline__ uint2 alpha_fit_full(const uint32_t v[16]) {
    // alpha.rs:194-212 on a full mask
    uint32_t mn = __vimin3_u32(v[0], v[1], v[2]), mx = __vimax3_u32(v[0], v[1], v[2]);
#pragma unroll
    for (int i = 3; i < 15; i += 2) { mn = __vimin3_u32(mn, v[i], v[i + 1]); mx = __vimax3_u32(mx, v[i], v[i + 1]); }
    mn = min(mn, v[15]); mx = max(mx, v[15]);
    uint32_t a5 = 0xFFFFFFFFu;            // min over (v-1) with 0 -> 0xFFFFFFFF, i.e. zeros ignored
    int b5 = (int)0x80000000;             // max over v + 0x7FFFFF01 with 255 -> INT_MIN, i.e. 255 ignored
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        a5 = __viaddmin_u32(v[i], 0xFFFFFFFFu, a5);
        b5 = __viaddmax_s32((int)v[i], 0x7FFFFF01, b5);
    }
    int min7 = (int)mn, max7 = (int)mx;
    int min5 = a5 == 0xFFFFFFFFu ? 255 : (int)(a5 + 1u);
    int max5 = b5 == (int)0x80000000 ? 0 : b5 - 0x7FFFFF01;
    if (min5 > max5) min5 = max5;
    fix_range(min5, max5, 5);
    fix_range(min7, max7, 7);
    int codes5[8], codes7[8];
    alpha_codebooks_fast(min5, max5, min7, max7, codes5, codes7);
    KeyBook k5, k7;
    make_keybook(codes5, k5);
    make_keybook(codes7, k7);
    int s5 = 0, s7 = 0;                                   // sums of the winning keys
    uint32_t w5lo = 0, w5hi = 0, w7lo = 0, w7hi = 0;      // 3-bit indices, 8 per word
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        // codes 0 and 1 (min5, max5) are the same in both books (alpha.rs:227-228, :238-239): shared keys
        const int x = (int)v[i];
        const int m01 = min(x * k5.a[0] + k5.b[0], x * k5.a[1] + k5.b[1]);
        int k[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) k[j] = x * k5.a[2 + j] + k5.b[2 + j];
        const int q5 = __vimin3_s32(__vimin3_s32(m01, k[0], k[1]), __vimin3_s32(k[2], k[3], k[4]), k[5]);
#pragma unroll
        for (int j = 0; j < 6; ++j) k[j] = x * k7.a[2 + j] + k7.b[2 + j];
        const int q7 = __vimin3_s32(__vimin3_s32(m01, k[0], k[1]), __vimin3_s32(k[2], k[3], k[4]), k[5]);
        s5 += q5; s7 += q7;
        const uint32_t sh = 1u << (3 * (i & 7));
        if (i < 8) { w5lo += ((uint32_t)q5 & 7u) * sh; w7lo += ((uint32_t)q7 & 7u) * sh; }
        else       { w5hi += ((uint32_t)q5 & 7u) * sh; w7hi += ((uint32_t)q7 & 7u) * sh; }
    }
    const int e5 = s5 - sum_fields3(w5lo, w5hi), e7 = s7 - sum_fields3(w7lo, w7hi);   // 8*(err - sum v^2)
    uint32_t a0, a1, g0, g1;
    if (e5 <= e7) {                                       // alpha.rs:251
        a0 = (uint32_t)min5; a1 = (uint32_t)max5; g0 = w5lo; g1 = w5hi;                // no swap: min5 <= max5 (Q3b)
    } else {
        // write_alpha_block7 always swaps after fix_range (Q3b)
        a0 = (uint32_t)max7; a1 = (uint32_t)min7;
        g0 = swap7_fields(w7lo); g1 = swap7_fields(w7hi);
    }
    return make_uint2(a0 | (a1 << 8) | (g0 << 16), (g0 >> 16) | (g1 << 8));
}

// ---- BC4 / BC5 encoder: one thread per block ---------------------------------------------------------
// Loads: 4 x 16-byte row segments per thread; consecutive threads read consecutive 16 B, so every warp
// load instruction covers 512 contiguous bytes per image row (fully coalesced without staging).
// Stores: 8 B (BC4) / 16 B (BC5) per thread, consecutive across the warp.
template <int FMT, int THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) alpha_encode_kernel(const BlockSource src, uint8_t* __restrict__ out) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= src.nblocks) return;
    uint32_t px[16];
    uint32_t mask;
    load_block_thread(src, b, px, mask);
    uint32_t v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = px[i] & 255u;    // channel 0 (lib.rs:200, :202)
    const bool full = mask == 0xFFFFu;                    // all blocks except image edges
    const uint2 r0 = full ? alpha_fit_full(v) : alpha_fit_thread(v, mask);
    if (FMT == BC4) {
        reinterpret_cast<uint2*>(out)[b] = r0;
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (px[i] >> 8) & 255u;   // channel 1 (lib.rs:203)
        const uint2 r1 = full ? alpha_fit_full(v) : alpha_fit_thread(v, mask);
        reinterpret_cast<uint4*>(out)[b] = make_uint4(r0.x, r0.y, r1.x, r1.y);
    }
}

}  // namespace txp
