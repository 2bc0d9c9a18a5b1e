/*
 * txp_oracle.c -- CPU restatement of the texpresso BC1..BC5 block codec.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path: it may be loaded by
 * tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, and by
 * nothing else.  The product (texpresso_b200/) never links, imports or falls back to it.
 *
 * It restates, in plain scalar C (fp32, no FMA contraction, no fast-math), the algorithm of the
 * reference at /root/reference/lib/src (Rust, cannot be compiled in this image: no cargo/rustc).
 * Every function cites the reference lines it follows.  Build: see oracle/Makefile
 * (gcc -O2 -ffp-contract=off -fno-fast-math; SSE2 scalar float arithmetic == Rust f32 arithmetic).
 *
 * Pinning: checked by tests/test_oracle_kat.py against every known-answer vector the reference's own
 * tests hold (lib/src/test_data.rs, lib/src/lib.rs:350-504).  What those vectors do NOT pin
 * (tie order of the 16-entry sort, NaN axes, 7-point alpha mode, perceptual weights ...) is
 * "parity unpinned" and is stated as such in DESIGN.md.
 *
 * Third-party arithmetic used by the reference: libm 0.2 sqrtf / truncf / roundf
 * (colourset.rs:108, math/vec3.rs:77-79, math/vec4.rs:104-107, math.rs:101).  All three are exactly
 * specified (correctly rounded sqrt, trunc, round-half-away) so the C library versions are equivalent.
 */
#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#include "single_lut_data.h"

#define TXO_API __attribute__((visibility("default")))

enum { TXO_BC1 = 0, TXO_BC2 = 1, TXO_BC3 = 2, TXO_BC4 = 3, TXO_BC5 = 4 };      /* lib.rs:40-46 */
enum { TXO_RANGE = 0, TXO_CLUSTER = 1, TXO_ITERATIVE = 2 };                    /* lib.rs:50-59 */

typedef struct {
    uint32_t algorithm;
    float weights[3];
    uint32_t weigh_colour_by_alpha;
} txo_params;                                                                   /* lib.rs:77-90 */

/* per-call instrumentation (used for flop accounting in bench.py; not part of the reference) */
typedef struct {
    uint64_t blocks;
    uint64_t single_blocks, range_blocks, cluster_blocks;
    uint64_t cand3, cand4;          /* ClusterFit candidates evaluated (cluster.rs:187 / :318 loop bodies) */
    uint64_t orderings3, orderings4;
    uint64_t count_hist[17];
} txo_stats;

static const uint8_t SINGLE_LUT[TXP_SINGLE_LUT_BYTES] = TXP_SINGLE_LUT_INIT;

/* ------------------------------------------------------------------------------------------------
 * scalar helpers with Rust semantics
 * ---------------------------------------------------------------------------------------------- */
typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } v4;

/* Rust f32::max / f32::min: if one operand is NaN the other is returned (math/vec4.rs:76-92) */
static inline float rmax(float a, float b) { if (a != a) return b; if (b != b) return a; return a > b ? a : b; }
static inline float rmin(float a, float b) { if (a != a) return b; if (b != b) return a; return a < b ? a : b; }

/* math.rs:100-102  roundf(a).max(0).min(limit) as i32 */
static inline int f32_to_i32_clamped(float a, int limit) {
    float r = roundf(a);
    r = rmax(r, 0.0f);
    r = rmin(r, (float)limit);
    return (int)r;
}

static inline float dot3(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  /* vec3.rs:51-53 */

/* ------------------------------------------------------------------------------------------------
 * ColourSet  (colourset.rs:26-142)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int count;
    v3 points[16];
    float weights[16];
    int8_t remap[16];
    int transparent;
} colourset;

/* colourset.rs:35-112 */
static void colourset_new(colourset *s, const uint8_t rgba[64], uint32_t mask, int fmt, int alpha_weighted) {
    memset(s, 0, sizeof *s);
    for (int i = 0; i < 16; ++i) {
        uint32_t bit = 1u << i;
        if ((mask & bit) == 0) { s->remap[i] = -1; continue; }                  /* :48-51 */
        if (fmt == TXO_BC1 && rgba[4 * i + 3] < 128) {                          /* :54-58 */
            s->remap[i] = -1; s->transparent = 1; continue;
        }
        for (int j = 0;; ++j) {
            if (j == i) {                                                       /* :63-80 */
                float x = (float)rgba[4 * i + 0] / 255.0f;
                float y = (float)rgba[4 * i + 1] / 255.0f;
                float z = (float)rgba[4 * i + 2] / 255.0f;
                float w = (float)((int)rgba[4 * i + 3] + 1) / 256.0f;
                s->points[s->count].x = x; s->points[s->count].y = y; s->points[s->count].z = z;
                s->weights[s->count] = alpha_weighted ? w : 1.0f;
                s->remap[i] = (int8_t)s->count;
                s->count += 1;
                break;
            }
            uint32_t oldbit = 1u << j;                                          /* :83-88 */
            int dup = ((mask & oldbit) != 0)
                && rgba[4 * i + 0] == rgba[4 * j + 0]
                && rgba[4 * i + 1] == rgba[4 * j + 1]
                && rgba[4 * i + 2] == rgba[4 * j + 2]
                && (fmt != TXO_BC1 || rgba[4 * j + 3] >= 128);
            if (dup) {                                                          /* :89-102 */
                int index = s->remap[j];
                float w = (float)((int)rgba[4 * i + 3] + 1) / 256.0f;
                s->weights[index] += alpha_weighted ? w : 1.0f;
                s->remap[i] = (int8_t)index;
                break;
            }
        }
    }
    for (int i = 0; i < 16; ++i) s->weights[i] = sqrtf(s->weights[i]);          /* :107-109 (all 16) */
}

/* colourset.rs:130-141 */
static void remap_indices(const colourset *s, const uint8_t source[16], uint8_t target[16]) {
    for (int i = 0; i < 16; ++i) {
        int j = s->remap[i];
        target[i] = (j == -1) ? 3 : source[j];
    }
}

/* ------------------------------------------------------------------------------------------------
 * math.rs: Sym3x3
 * ---------------------------------------------------------------------------------------------- */
/* math.rs:44-73 */
static void weighted_covariance(const v3 *points, const float *weights, int n, float cov[6]) {
    float total = 0.0f;
    for (int i = 0; i < n; ++i) total = total + weights[i];                     /* :48 iter().sum() */
    v3 c = {0.0f, 0.0f, 0.0f};
    for (int i = 0; i < n; ++i) {                                               /* :49, vec3.rs:341-345 */
        c.x = c.x + points[i].x * weights[i];
        c.y = c.y + points[i].y * weights[i];
        c.z = c.z + points[i].z * weights[i];
    }
    if (total > FLT_EPSILON) {                                                  /* :51-55, vec3.rs:316-322 */
        c.x = c.x / total; c.y = c.y / total; c.z = c.z / total;
    }
    for (int k = 0; k < 6; ++k) cov[k] = 0.0f;
    for (int i = 0; i < n; ++i) {                                               /* :60-70 */
        float ax = points[i].x - c.x, ay = points[i].y - c.y, az = points[i].z - c.z;
        float bx = ax * weights[i], by = ay * weights[i], bz = az * weights[i];
        cov[0] += ax * bx;
        cov[1] += ax * by;
        cov[2] += ax * bz;
        cov[3] += ay * by;
        cov[4] += ay * bz;
        cov[5] += az * bz;
    }
}

/* math.rs:75-97 */
static v3 principle_component(const float m[6]) {
    float r0[3] = {m[0], m[1], m[2]}, r1[3] = {m[1], m[3], m[4]}, r2[3] = {m[2], m[4], m[5]};
    float v[3] = {1.0f, 1.0f, 1.0f};
    for (int it = 0; it < 8; ++it) {
        float w[3];
        for (int k = 0; k < 3; ++k) {
            float t = r0[k] * v[0];                                             /* :85 */
            t = r1[k] * v[1] + t;                                               /* :86 */
            t = r2[k] * v[2] + t;                                               /* :87 */
            w[k] = t;
        }
        float a = rmax(w[0], rmax(w[1], w[2]));                                 /* :90 */
        float ra = 1.0f / a;                                                    /* :93, vec4.rs:94-96 */
        v[0] = w[0] * ra; v[1] = w[1] * ra; v[2] = w[2] * ra;
    }
    v3 r = {v[0], v[1], v[2]};
    return r;
}

/* ------------------------------------------------------------------------------------------------
 * colourblock.rs
 * ---------------------------------------------------------------------------------------------- */
/* colourblock.rs:28-34 */
static uint16_t pack_565(v3 c) {
    uint16_t r = (uint16_t)f32_to_i32_clamped(31.0f * c.x, 31);
    uint16_t g = (uint16_t)f32_to_i32_clamped(63.0f * c.y, 63);
    uint16_t b = (uint16_t)f32_to_i32_clamped(31.0f * c.z, 31);
    return (uint16_t)((r << 11) | (g << 5) | b);
}

/* colourblock.rs:36-53 */
static void write_block(uint16_t a, uint16_t b, const uint8_t idx[16], uint8_t block[8]) {
    block[0] = (uint8_t)(a & 0xFF); block[1] = (uint8_t)(a >> 8);
    block[2] = (uint8_t)(b & 0xFF); block[3] = (uint8_t)(b >> 8);
    for (int i = 0; i < 4; ++i)
        block[4 + i] = (uint8_t)(((idx[4 * i + 3] & 3) << 6) | ((idx[4 * i + 2] & 3) << 4)
                                 | ((idx[4 * i + 1] & 3) << 2) | (idx[4 * i] & 3));
}

/* colourblock.rs:55-74 */
static void write3(v3 start, v3 end, const uint8_t indices[16], uint8_t block[8]) {
    uint16_t a = pack_565(start), b = pack_565(end);
    uint8_t remapped[16];
    memcpy(remapped, indices, 16);
    if (a > b) {
        uint16_t t = a; a = b; b = t;
        for (int i = 0; i < 16; ++i) {
            if (remapped[i] == 0) remapped[i] = 1;
            else if (remapped[i] == 1) remapped[i] = 0;
        }
    }
    write_block(a, b, remapped, block);
}

/* colourblock.rs:76-94 */
static void write4(v3 start, v3 end, const uint8_t indices[16], uint8_t block[8]) {
    uint16_t a = pack_565(start), b = pack_565(end);
    uint8_t remapped[16];
    memset(remapped, 0, 16);
    if (a < b) {
        uint16_t t = a; a = b; b = t;
        for (int i = 0; i < 16; ++i) remapped[i] = (uint8_t)((indices[i] ^ 1) & 3);
    } else if (a > b) {
        memcpy(remapped, indices, 16);
    }
    write_block(a, b, remapped, block);
}

/* colourblock.rs:97-113 */
static void unpack_565(const uint8_t p[2], uint8_t out[4]) {
    uint16_t value = (uint16_t)(p[0] | (p[1] << 8));
    uint8_t r = (uint8_t)((value >> 11) & 0x1F), g = (uint8_t)((value >> 5) & 0x3F), b = (uint8_t)(value & 0x1F);
    out[0] = (uint8_t)((r << 3) | (r >> 2));
    out[1] = (uint8_t)((g << 2) | (g >> 4));
    out[2] = (uint8_t)((b << 3) | (b >> 2));
    out[3] = 255;
}

/* colourblock.rs:116-169 */
static void colour_decompress(const uint8_t bytes[8], int is_bc1, uint8_t rgba[64]) {
    uint8_t codes[16];
    uint16_t a = (uint16_t)(bytes[0] | (bytes[1] << 8));
    uint16_t b = (uint16_t)(bytes[2] | (bytes[3] << 8));
    unpack_565(bytes, codes);
    unpack_565(bytes + 2, codes + 4);
    for (int i = 0; i < 4; ++i) {
        uint32_t c = codes[i], d = codes[4 + i];
        if (is_bc1 && a <= b) {
            codes[8 + i] = (uint8_t)((c + d) / 2);
            codes[12 + i] = 0;
        } else {
            codes[8 + i] = (uint8_t)((2 * c + d) / 3);
            codes[12 + i] = (uint8_t)((c + 2 * d) / 3);
        }
    }
    codes[8 + 3] = 255;
    codes[12 + 3] = (is_bc1 && a <= b) ? 0 : 255;
    for (int i = 0; i < 4; ++i) {
        uint8_t packed = bytes[4 + i];
        for (int k = 0; k < 4; ++k) {
            int ind = (packed >> (2 * k)) & 3;
            memcpy(rgba + 4 * (4 * i + k), codes + 4 * ind, 4);
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * alpha.rs
 * ---------------------------------------------------------------------------------------------- */
/* alpha.rs:27-51 */
static void compress_bc2(const uint8_t rgba[64], uint32_t mask, uint8_t block[8]) {
    for (int i = 0; i < 8; ++i) {
        float alpha1 = (float)rgba[4 * (2 * i) + 3] * (15.0f / 255.0f);
        float alpha2 = (float)rgba[4 * (2 * i + 1) + 3] * (15.0f / 255.0f);
        uint8_t q1 = (uint8_t)f32_to_i32_clamped(alpha1, 15);
        uint8_t q2 = (uint8_t)f32_to_i32_clamped(alpha2, 15);
        if ((mask & (1u << (2 * i))) == 0) q1 = 0;
        if ((mask & (1u << (2 * i + 1))) == 0) q2 = 0;
        block[i] = (uint8_t)(q1 | (q2 << 4));
    }
}

/* alpha.rs:53-68 */
static void decompress_bc2(uint8_t rgba[64], const uint8_t bytes[8]) {
    for (int i = 0; i < 8; ++i) {
        uint8_t q = bytes[i], lo = q & 0x0F, hi = q & 0xF0;
        rgba[4 * (2 * i) + 3] = (uint8_t)(lo | (lo << 4));
        rgba[4 * (2 * i + 1) + 3] = (uint8_t)(hi | (hi >> 4));
    }
}

/* alpha.rs:70-77 */
static void fix_range(uint8_t *mn, uint8_t *mx, int steps) {
    if ((int)*mx - (int)*mn < steps) { int v = (int)*mn + steps; *mx = (uint8_t)(v < 255 ? v : 255); }
    if ((int)*mx - (int)*mn < steps) { int v = (int)*mx - steps; *mn = (uint8_t)(v > 0 ? v : 0); }
}

/* alpha.rs:79-119 */
static uint32_t fit_codes(const uint8_t rgba[64], int channel, uint32_t mask, const uint8_t codes[8], uint8_t indices[16]) {
    uint32_t err = 0;
    for (int i = 0; i < 16; ++i) {
        if ((mask & (1u << i)) == 0) { indices[i] = 0; continue; }
        int value = rgba[4 * i + channel];
        uint32_t least = UINT32_MAX; uint8_t index = 0;
        for (int j = 0; j < 8; ++j) {
            int dist = value - (int)codes[j];
            uint32_t d2 = (uint32_t)(dist * dist);
            if (d2 < least) { least = d2; index = (uint8_t)j; }
        }
        indices[i] = index;
        err += least;
    }
    return err;
}

/* alpha.rs:121-144 */
static void write_alpha_block(uint8_t a0, uint8_t a1, const uint8_t indices[16], uint8_t block[8]) {
    block[0] = a0; block[1] = a1;
    for (int i = 0; i < 2; ++i) {
        uint32_t value = 0;
        for (int j = 0; j < 8; ++j) value |= (uint32_t)indices[8 * i + j] << (3 * j);
        for (int j = 0; j < 3; ++j) block[2 + 3 * i + j] = (uint8_t)((value >> (8 * j)) & 0xFF);
    }
}

/* alpha.rs:146-165 */
static void write_alpha_block5(uint8_t a0, uint8_t a1, const uint8_t indices[16], uint8_t block[8]) {
    if (a0 > a1) {
        uint8_t sw[16];
        for (int i = 0; i < 16; ++i) {
            uint8_t x = indices[i];
            sw[i] = x == 0 ? 1 : x == 1 ? 0 : (x <= 5 ? (uint8_t)(7 - x) : x);
        }
        write_alpha_block(a1, a0, sw, block);
    } else write_alpha_block(a0, a1, indices, block);
}

/* alpha.rs:167-185 */
static void write_alpha_block7(uint8_t a0, uint8_t a1, const uint8_t indices[16], uint8_t block[8]) {
    if (a0 < a1) {
        uint8_t sw[16];
        for (int i = 0; i < 16; ++i) {
            uint8_t x = indices[i];
            sw[i] = x == 0 ? 1 : x == 1 ? 0 : (uint8_t)(9 - x);
        }
        write_alpha_block(a1, a0, sw, block);
    } else write_alpha_block(a0, a1, indices, block);
}

/* alpha.rs:187-256 */
static void compress_bc3(const uint8_t rgba[64], int channel, uint32_t mask, uint8_t block[8]) {
    uint8_t min5 = 255, max5 = 0, min7 = 255, max7 = 0;
    for (int i = 0; i < 16; ++i) {
        if ((mask & (1u << i)) == 0) continue;
        uint8_t v = rgba[4 * i + channel];
        if (v < min7) min7 = v;
        if (v > max7) max7 = v;
        if (v != 0 && v < min5) min5 = v;
        if (v != 255 && v > max5) max5 = v;
    }
    if (min5 > max5) min5 = max5;                                               /* :215-217 */
    if (min7 > max7) min7 = max7;                                               /* :218-220 */
    fix_range(&min5, &max5, 5);
    fix_range(&min7, &max7, 7);

    uint8_t codes5[8], codes7[8];
    codes5[0] = min5; codes5[1] = max5;
    for (int i = 1; i < 5; ++i) codes5[1 + i] = (uint8_t)(((5 - i) * (int)min5 + i * (int)max5) / 5);
    codes5[6] = 0; codes5[7] = 255;
    codes7[0] = min5; codes7[1] = max5;                                         /* :238-239 (sic: min5/max5) */
    for (int i = 1; i < 7; ++i) codes7[1 + i] = (uint8_t)(((7 - i) * (int)min7 + i * (int)max7) / 7);

    uint8_t ind5[16], ind7[16];
    uint32_t err5 = fit_codes(rgba, channel, mask, codes5, ind5);
    uint32_t err7 = fit_codes(rgba, channel, mask, codes7, ind7);
    if (err5 <= err7) write_alpha_block5(min5, max5, ind5, block);
    else write_alpha_block7(min7, max7, ind7, block);
}

/* alpha.rs:258-304 */
static void decompress_bc3(uint8_t rgba[64], int channel, const uint8_t bytes[8]) {
    int a0 = bytes[0], a1 = bytes[1];
    uint8_t codes[8];
    codes[0] = bytes[0]; codes[1] = bytes[1];
    if (a0 <= a1) {
        for (int i = 1; i < 5; ++i) codes[1 + i] = (uint8_t)(((5 - i) * a0 + i * a1) / 5);
        codes[6] = 0; codes[7] = 255;
    } else {
        for (int i = 1; i < 7; ++i) codes[1 + i] = (uint8_t)(((7 - i) * a0 + i * a1) / 7);
    }
    for (int i = 0; i < 2; ++i) {
        int value = 0;
        for (int j = 0; j < 3; ++j) value |= (int)bytes[2 + 3 * i + j] << (8 * j);
        for (int j = 0; j < 8; ++j) rgba[4 * (8 * i + j) + channel] = codes[(value >> (3 * j)) & 7];
    }
}

/* ------------------------------------------------------------------------------------------------
 * SingleColourFit  (colourfit/single.rs)
 * ---------------------------------------------------------------------------------------------- */
typedef struct { v3 start, end; uint8_t index; uint32_t error; } single_result;

/* single.rs:58-106; table ids: 0=5_3 1=6_3 2=5_4 3=6_4 */
static single_result single_compute_endpoints(const colourset *s, const int lut[3]) {
    int colour[3] = {
        f32_to_i32_clamped(s->points[0].x * 255.0f, 255),
        f32_to_i32_clamped(s->points[0].y * 255.0f, 255),
        f32_to_i32_clamped(s->points[0].z * 255.0f, 255),
    };
    single_result r; memset(&r, 0, sizeof r); r.error = UINT32_MAX;
    for (int index = 0; index < 2; ++index) {
        uint32_t error = 0;
        const uint8_t *src[3];
        for (int ch = 0; ch < 3; ++ch) {
            src[ch] = &SINGLE_LUT[((lut[ch] * 256 + colour[ch]) * 2 + index) * 3];
            uint32_t diff = src[ch][2];
            error += diff * diff;
        }
        if (error < r.error) {
            r.start.x = (float)src[0][0] / 31.0f; r.start.y = (float)src[1][0] / 63.0f; r.start.z = (float)src[2][0] / 31.0f;
            r.end.x = (float)src[0][1] / 31.0f;   r.end.y = (float)src[1][1] / 63.0f;   r.end.z = (float)src[2][1] / 31.0f;
            r.index = (uint8_t)(2 * index);
            r.error = error;
        }
    }
    return r;
}

/* single.rs:122-164 + colourfit.rs:48-59 */
static void single_compress(const colourset *s, int fmt, uint8_t block[8]) {
    uint32_t best_error = UINT32_MAX;
    memset(block, 0, 8);
    uint8_t src[16], indices[16];
    if (fmt == TXO_BC1) {
        const int lut3[3] = {0, 1, 0};
        single_result r = single_compute_endpoints(s, lut3);
        if (r.error < best_error) {
            memset(src, r.index, 16);
            remap_indices(s, src, indices);
            write3(r.start, r.end, indices, block);
            best_error = r.error;
        }
        if (s->transparent) return;
    }
    const int lut4[3] = {2, 3, 2};
    single_result r = single_compute_endpoints(s, lut4);
    if (r.error < best_error) {
        memset(src, r.index, 16);
        remap_indices(s, src, indices);
        write4(r.start, r.end, indices, block);
    }
}

/* ------------------------------------------------------------------------------------------------
 * RangeFit  (colourfit/range.rs)
 * ---------------------------------------------------------------------------------------------- */
static inline float snap(float grid, float v, float gridrcp) { return truncf(grid * v + 0.5f) * gridrcp; }

/* range.rs:44-192 + colourfit.rs:48-59 */
static void range_compress(const colourset *s, int fmt, const float mw[3], uint8_t block[8]) {
    int count = s->count;
    float cov[6];
    weighted_covariance(s->points, s->weights, count, cov);
    v3 principle = principle_component(cov);
    v3 start = {0, 0, 0}, end = {0, 0, 0};
    if (count > 0) {                                                            /* :67-86 */
        start = s->points[0]; end = start;
        float mn = dot3(start, principle), mx = mn;
        for (int i = 1; i < count; ++i) {
            float d = dot3(s->points[i], principle);
            if (d < mn) { start = s->points[i]; mn = d; }
            else if (d > mx) { end = s->points[i]; mx = d; }
        }
    }
    /* :88-98 */
    const float g[3] = {31.0f, 63.0f, 31.0f}, gr[3] = {1.0f / 31.0f, 1.0f / 63.0f, 1.0f / 31.0f};
    float sv[3] = {start.x, start.y, start.z}, ev[3] = {end.x, end.y, end.z};
    for (int k = 0; k < 3; ++k) {
        sv[k] = rmin(1.0f, rmax(0.0f, sv[k]));
        ev[k] = rmin(1.0f, rmax(0.0f, ev[k]));
        sv[k] = snap(g[k], sv[k], gr[k]);
        ev[k] = snap(g[k], ev[k], gr[k]);
    }
    v3 fs = {sv[0], sv[1], sv[2]}, fe = {ev[0], ev[1], ev[2]};

    float best_error = FLT_MAX;
    memset(block, 0, 8);
    for (int pass = 0; pass < 2; ++pass) {
        int three = (pass == 0);
        if (three && fmt != TXO_BC1) continue;                                  /* colourfit.rs:48-56 */
        if (!three && fmt == TXO_BC1 && s->transparent) continue;
        float codes[4][3]; int ncodes;
        for (int k = 0; k < 3; ++k) { codes[0][k] = sv[k]; codes[1][k] = ev[k]; }
        if (three) {                                                            /* :161 */
            ncodes = 3;
            for (int k = 0; k < 3; ++k) codes[2][k] = sv[k] * 0.5f + ev[k] * 0.5f;
        } else {                                                                /* :176-181 */
            ncodes = 4;
            for (int k = 0; k < 3; ++k) {
                codes[2][k] = sv[k] * (2.0f / 3.0f) + ev[k] * (1.0f / 3.0f);
                codes[3][k] = sv[k] * (1.0f / 3.0f) + ev[k] * (2.0f / 3.0f);
            }
        }
        /* compression_helper :103-143 */
        uint8_t closest[16]; memset(closest, 0, 16);
        float error = 0.0f;
        for (int i = 0; i < count; ++i) {
            float dist = FLT_MAX; int idx = 0;
            float p[3] = {s->points[i].x, s->points[i].y, s->points[i].z};
            for (int j = 0; j < ncodes; ++j) {
                float dx = mw[0] * (p[0] - codes[j][0]);
                float dy = mw[1] * (p[1] - codes[j][1]);
                float dz = mw[2] * (p[2] - codes[j][2]);
                float d = dx * dx + dy * dy + dz * dz;                          /* vec3.rs:55-57 */
                if (d < dist) { dist = d; idx = j; }
            }
            closest[i] = (uint8_t)idx;
            error += dist;
        }
        if (error < best_error) {
            uint8_t indices[16];
            remap_indices(s, closest, indices);
            best_error = error;
            if (three) write3(fs, fe, indices, block); else write4(fs, fe, indices, block);
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * ClusterFit  (colourfit/cluster.rs)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const colourset *set;
    int fmt;
    float mw[3];                /* metric weights */
    int num_iterations;
    v3 principle;
    uint8_t order[8][16];
    v4 pw[16];                  /* points_weights */
    v4 xsum_wsum;
    float best_error;           /* all four lanes of the reference's Vec4 are equal */
    uint8_t best_compressed[8];
    txo_stats *stats;
} clusterfit;

/* cluster.rs:90-97 */
static int fcmp_less(float a, float b) {
    int fa = isfinite(a), fb = isfinite(b);
    if (!fa && !fb) return 0;       /* Equal */
    if (!fa) return 0;              /* Greater */
    if (!fb) return 1;              /* Less */
    return a < b;
}

/* cluster.rs:78-136 */
static int construct_ordering(clusterfit *f, v3 axis, int iteration) {
    int count = f->set->count;
    int idx[16]; float dp[16];
    for (int i = 0; i < 16; ++i) { idx[i] = 0; dp[i] = FLT_MAX; }
    for (int i = 0; i < count; ++i) { idx[i] = i; dp[i] = dot3(f->set->points[i], axis); }
    /* sort_unstable_by on 16 elements == insertion sort (stable) in every std version */
    for (int i = 1; i < 16; ++i) {
        int ti = idx[i]; float td = dp[i];
        int j = i;
        while (j > 0 && fcmp_less(td, dp[j - 1])) { idx[j] = idx[j - 1]; dp[j] = dp[j - 1]; --j; }
        idx[j] = ti; dp[j] = td;
    }
    for (int i = 0; i < 16; ++i) f->order[iteration][i] = (uint8_t)idx[i];
    for (int it = 0; it < iteration; ++it)                                       /* :108-120 */
        if (memcmp(f->order[it], f->order[iteration], 16) == 0) return 0;
    v4 sum = {0, 0, 0, 0};                                                      /* :123-133 */
    for (int i = 0; i < count; ++i) {
        int j = f->order[iteration][i];
        float w = f->set->weights[j];
        v4 x = { f->set->points[j].x * w, f->set->points[j].y * w, f->set->points[j].z * w, 1.0f * w };
        f->pw[i] = x;
        sum.x += x.x; sum.y += x.y; sum.z += x.z; sum.w += x.w;
    }
    f->xsum_wsum = sum;
    return 1;
}

typedef struct { float ax, ay, az, bx, by, bz, error; } lsq;

/* the shared tail of both searches: cluster.rs:201-220 == :334-353.
 * alphax/betax are (x,y,z,w) with w = alpha2_sum / beta2_sum. */
static inline lsq solve(v4 alphax, v4 betax, float alphabeta, const float mw[3]) {
    float alpha2 = alphax.w, beta2 = betax.w;
    float factor = 1.0f / (alpha2 * beta2 - alphabeta * alphabeta);
    float av[3] = {alphax.x, alphax.y, alphax.z}, bv[3] = {betax.x, betax.y, betax.z};
    const float g[3] = {31.0f, 63.0f, 31.0f}, gr[3] = {1.0f / 31.0f, 1.0f / 63.0f, 1.0f / 31.0f};
    float a[3], b[3], e5[3];
    for (int k = 0; k < 3; ++k) {
        a[k] = ((av[k] * beta2) - bv[k] * alphabeta) * factor;
        b[k] = ((bv[k] * alpha2) - av[k] * alphabeta) * factor;
        a[k] = rmin(1.0f, rmax(0.0f, a[k]));
        b[k] = rmin(1.0f, rmax(0.0f, b[k]));
        a[k] = truncf(g[k] * a[k] + 0.5f) * gr[k];
        b[k] = truncf(g[k] * b[k] + 0.5f) * gr[k];
        float e1 = (a[k] * a[k]) * alpha2 + (b[k] * b[k] * beta2);
        float e2 = (a[k] * b[k] * alphabeta) - a[k] * av[k];
        float e3 = e2 - b[k] * bv[k];
        float e4 = 2.0f * e3 + e1;
        e5[k] = e4 * mw[k];
    }
    lsq r = { a[0], a[1], a[2], b[0], b[1], b[2], (e5[0] + e5[1]) + e5[2] };
    return r;
}

static inline v4 v4add(v4 a, v4 b) { v4 r = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; return r; }
static inline v4 v4sub(v4 a, v4 b) { v4 r = {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; return r; }
static inline v4 v4mul(v4 a, v4 b) { v4 r = {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; return r; }

/* cluster.rs:152-274 */
static void cluster_compress3(clusterfit *f) {
    const int count = f->set->count;
    const v4 zero = {0, 0, 0, 0};
    const v4 half_half2 = {0.5f, 0.5f, 0.5f, 0.25f};
    v3 best_start = {0, 0, 0}, best_end = {0, 0, 0};
    float best_error = f->best_error;
    int best_iteration = 0, best_i = 0, best_j = 0;
    v3 axis = f->principle;

    for (int it = 0; it < f->num_iterations; ++it) {
        if (!construct_ordering(f, axis, it)) break;
        if (f->stats) f->stats->orderings3++;
        v4 part0 = zero;
        for (int i = 0; i < count; ++i) {
            v4 part1 = (i == 0) ? f->pw[0] : zero;
            int jmin = (i == 0) ? 1 : i;
            for (int j = jmin; j <= count; ++j) {
                v4 part2 = v4sub(v4sub(f->xsum_wsum, part1), part0);
                v4 p1h = v4mul(part1, half_half2);
                v4 alphax = v4add(p1h, part0);
                v4 betax = v4add(p1h, part2);
                float alphabeta = p1h.w;
                lsq r = solve(alphax, betax, alphabeta, f->mw);
                if (f->stats) f->stats->cand3++;
                if (r.error < best_error) {
                    best_start.x = r.ax; best_start.y = r.ay; best_start.z = r.az;
                    best_end.x = r.bx; best_end.y = r.by; best_end.z = r.bz;
                    best_i = i; best_j = j; best_error = r.error; best_iteration = it;
                }
                if (j < count) part1 = v4add(part1, f->pw[j]);
            }
            part0 = v4add(part0, f->pw[i]);
        }
        if (best_iteration != it) break;
        axis.x = best_end.x - best_start.x; axis.y = best_end.y - best_start.y; axis.z = best_end.z - best_start.z;
    }

    if (best_error < f->best_error) {
        const uint8_t *order = f->order[best_iteration];
        uint8_t unordered[16], best_indices[16];
        memset(unordered, 0, 16);
        for (int m = best_i; m < best_j; ++m) unordered[order[m]] = 2;
        for (int m = best_j; m < count; ++m) unordered[order[m]] = 1;
        remap_indices(f->set, unordered, best_indices);
        write3(best_start, best_end, best_indices, f->best_compressed);
        f->best_error = best_error;
    }
}

/* cluster.rs:276-417 */
static void cluster_compress4(clusterfit *f) {
    const int count = f->set->count;
    const v4 zero = {0, 0, 0, 0};
    const v4 c13 = {1.0f / 3.0f, 1.0f / 3.0f, 1.0f / 3.0f, 1.0f / 9.0f};
    const v4 c23 = {2.0f / 3.0f, 2.0f / 3.0f, 2.0f / 3.0f, 4.0f / 9.0f};
    const float twoninths = 2.0f / 9.0f;
    v3 best_start = {0, 0, 0}, best_end = {0, 0, 0};
    float best_error = f->best_error;
    int best_iteration = 0, best_i = 0, best_j = 0, best_k = 0;
    v3 axis = f->principle;

    for (int it = 0; it < f->num_iterations; ++it) {
        if (!construct_ordering(f, axis, it)) break;
        if (f->stats) f->stats->orderings4++;
        v4 part0 = zero;
        for (int i = 0; i < count; ++i) {
            v4 part1 = zero;
            for (int j = i; j <= count; ++j) {
                v4 part2 = (j == 0) ? f->pw[0] : zero;
                int kmin = (j == 0) ? 1 : j;
                for (int k = kmin; k <= count; ++k) {
                    v4 part3 = v4sub(v4sub(v4sub(f->xsum_wsum, part2), part1), part0);
                    v4 alphax = v4add(v4mul(part2, c13), v4add(v4mul(part1, c23), part0));
                    v4 betax = v4add(v4mul(part1, c13), v4add(v4mul(part2, c23), part3));
                    float alphabeta = twoninths * (part1.w + part2.w);
                    lsq r = solve(alphax, betax, alphabeta, f->mw);
                    if (f->stats) f->stats->cand4++;
                    if (r.error < best_error) {
                        best_start.x = r.ax; best_start.y = r.ay; best_start.z = r.az;
                        best_end.x = r.bx; best_end.y = r.by; best_end.z = r.bz;
                        best_i = i; best_j = j; best_k = k; best_error = r.error; best_iteration = it;
                    }
                    if (k < count) part2 = v4add(part2, f->pw[k]);
                }
                if (j < count) part1 = v4add(part1, f->pw[j]);
            }
            part0 = v4add(part0, f->pw[i]);
        }
        if (best_iteration != it) break;
        axis.x = best_end.x - best_start.x; axis.y = best_end.y - best_start.y; axis.z = best_end.z - best_start.z;
    }

    if (best_error < f->best_error) {
        const uint8_t *order = f->order[best_iteration];
        uint8_t unordered[16], best_indices[16];
        memset(unordered, 0, 16);
        for (int m = best_i; m < best_j; ++m) unordered[order[m]] = 2;
        for (int m = best_j; m < count; ++m) unordered[order[m]] = 3;
        for (int m = best_k; m < count; ++m) unordered[order[m]] = 1;
        remap_indices(f->set, unordered, best_indices);
        write4(best_start, best_end, best_indices, f->best_compressed);
        f->best_error = best_error;
    }
}

/* cluster.rs:49-76 + colourfit.rs:48-59 */
static void cluster_compress(const colourset *s, int fmt, const float mw[3], int iterate, uint8_t block[8], txo_stats *st) {
    clusterfit f;
    memset(&f, 0, sizeof f);
    f.set = s; f.fmt = fmt; f.mw[0] = mw[0]; f.mw[1] = mw[1]; f.mw[2] = mw[2];
    f.num_iterations = iterate ? 8 : 1;
    f.best_error = FLT_MAX;
    f.stats = st;
    float cov[6];
    weighted_covariance(s->points, s->weights, s->count, cov);
    f.principle = principle_component(cov);
    if (fmt == TXO_BC1) {
        cluster_compress3(&f);
        if (!s->transparent) cluster_compress4(&f);
    } else {
        cluster_compress4(&f);
    }
    memcpy(block, f.best_compressed, 8);
}

/* ------------------------------------------------------------------------------------------------
 * lib.rs: block dispatch and image loops
 * ---------------------------------------------------------------------------------------------- */
static size_t block_size(int fmt) { return (fmt == TXO_BC1 || fmt == TXO_BC4) ? 8 : 16; }   /* lib.rs:159-168 */

/* lib.rs:188-234 */
static void compress_block_masked(int fmt, const uint8_t rgba[64], uint32_t mask, const txo_params *p, uint8_t *out, txo_stats *st) {
    switch (fmt) {
    case TXO_BC2: compress_bc2(rgba, mask, out); break;
    case TXO_BC3: compress_bc3(rgba, 3, mask, out); break;
    case TXO_BC4: compress_bc3(rgba, 0, mask, out); break;
    case TXO_BC5: compress_bc3(rgba, 0, mask, out); compress_bc3(rgba, 1, mask, out + 8); break;
    default: break;
    }
    if (st) st->blocks++;
    if (fmt == TXO_BC1 || fmt == TXO_BC2 || fmt == TXO_BC3) {
        colourset set;
        colourset_new(&set, rgba, mask, fmt, p->weigh_colour_by_alpha != 0);
        uint8_t *cb = out + (fmt == TXO_BC1 ? 0 : 8);
        if (st) st->count_hist[set.count]++;
        if (set.count == 1) {
            single_compress(&set, fmt, cb);
            if (st) st->single_blocks++;
        } else if (p->algorithm == TXO_RANGE || set.count == 0) {
            range_compress(&set, fmt, p->weights, cb);
            if (st) st->range_blocks++;
        } else {
            cluster_compress(&set, fmt, p->weights, p->algorithm == TXO_ITERATIVE, cb, st);
            if (st) st->cluster_blocks++;
        }
    }
}

/* lib.rs:240-277 */
static void decompress_block(int fmt, const uint8_t *block, uint8_t rgba[64]) {
    if (fmt == TXO_BC1 || fmt == TXO_BC2 || fmt == TXO_BC3) {
        colour_decompress(block + (fmt == TXO_BC1 ? 0 : 8), fmt == TXO_BC1, rgba);
    } else {
        for (int i = 0; i < 16; ++i) { rgba[4 * i] = 0; rgba[4 * i + 1] = 0; rgba[4 * i + 2] = 0; rgba[4 * i + 3] = 255; }
    }
    switch (fmt) {
    case TXO_BC2: decompress_bc2(rgba, block); break;
    case TXO_BC3: decompress_bc3(rgba, 3, block); break;
    case TXO_BC4:
        decompress_bc3(rgba, 0, block);
        for (int i = 0; i < 16; ++i) { rgba[4 * i + 1] = rgba[4 * i]; rgba[4 * i + 2] = rgba[4 * i]; }
        break;
    case TXO_BC5: decompress_bc3(rgba, 0, block); decompress_bc3(rgba, 1, block + 8); break;
    default: break;
    }
}

typedef struct {
    int fmt; const uint8_t *rgba; size_t w, h; const txo_params *p; uint8_t *out; size_t out_len;
    size_t row0, row1; txo_stats stats; int want_stats;
} job;

/* lib.rs:305-334, one call per block row y (the rayon task grain) */
static void compress_rows(job *jb) {
    size_t bs = block_size(jb->fmt), bw = (jb->w + 3) / 4, rowbytes = bw * bs;
    for (size_t y = jb->row0; y < jb->row1; ++y) {
        uint8_t src[64];
        memset(src, 0, 64);                                                     /* :306 (per row, stale across blocks) */
        size_t avail = jb->out_len - y * rowbytes;
        size_t nblk = (avail >= rowbytes) ? bw : avail / bs;
        for (size_t x = 0; x < nblk; ++x) {
            uint32_t mask = 0;
            for (int py = 0; py < 4; ++py)
                for (int px = 0; px < 4; ++px) {
                    size_t sx = 4 * x + px, sy = 4 * y + py;
                    if (sx < jb->w && sy < jb->h) {
                        memcpy(src + 4 * (4 * py + px), jb->rgba + 4 * (jb->w * sy + sx), 4);
                        mask |= 1u << (4 * py + px);
                    }
                }
            compress_block_masked(jb->fmt, src, mask, jb->p, jb->out + y * rowbytes + x * bs, jb->want_stats ? &jb->stats : NULL);
        }
    }
}

static void *compress_thread(void *arg) { compress_rows((job *)arg); return NULL; }

TXO_API size_t txo_block_size(int fmt) { return block_size(fmt); }
TXO_API size_t txo_compressed_size(int fmt, size_t w, size_t h) { return ((w + 3) / 4) * ((h + 3) / 4) * block_size(fmt); }   /* lib.rs:175-179 */

/* lib.rs:287-335.  threads<=1: serial.  Returns 0, or -1 where the reference would panic.
 * Rows beyond compressed_size that fit in out_len are encoded as fully masked blocks. */
TXO_API int txo_compress(int fmt, const uint8_t *rgba, size_t w, size_t h, const txo_params *p,
                         uint8_t *out, size_t out_len, int threads, txo_stats *stats) {
    if (fmt < 0 || fmt > 4 || w == 0) return -1;
    if (out_len < txo_compressed_size(fmt, w, h)) return -1;                    /* lib.rs:295 */
    size_t bs = block_size(fmt), bw = (w + 3) / 4, rowbytes = bw * bs;
    if (out_len % bs) return -1;                                                /* partial block slice panics */
    size_t rows = (out_len + rowbytes - 1) / rowbytes;
    if (stats) memset(stats, 0, sizeof *stats);
    if (threads < 1) threads = 1;
    if ((size_t)threads > rows) threads = rows ? (int)rows : 1;
    job *jobs = (job *)calloc((size_t)threads, sizeof(job));
    pthread_t *tids = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; ++t) {
        job *jb = &jobs[t];
        jb->fmt = fmt; jb->rgba = rgba; jb->w = w; jb->h = h; jb->p = p; jb->out = out; jb->out_len = out_len;
        jb->row0 = rows * (size_t)t / (size_t)threads; jb->row1 = rows * (size_t)(t + 1) / (size_t)threads;
        jb->want_stats = stats != NULL;
        if (threads > 1) pthread_create(&tids[t], NULL, compress_thread, jb); else compress_rows(jb);
    }
    for (int t = 0; t < threads; ++t) {
        if (threads > 1) pthread_join(tids[t], NULL);
        if (stats) {
            const uint64_t *s = (const uint64_t *)&jobs[t].stats; uint64_t *d = (uint64_t *)stats;
            for (size_t k = 0; k < sizeof(txo_stats) / 8; ++k) d[k] += s[k];
        }
    }
    free(jobs); free(tids);
    return 0;
}

/* lib.rs:124-156 */
TXO_API int txo_decompress(int fmt, const uint8_t *data, size_t data_len, size_t w, size_t h, uint8_t *out, size_t out_len) {
    if (fmt < 0 || fmt > 4 || w == 0) return -1;
    size_t bs = block_size(fmt), bw = (w + 3) / 4;
    size_t chunk = w * 16;
    size_t rows = (out_len + chunk - 1) / chunk;
    for (size_t y = 0; y < rows; ++y) {
        size_t rowlen = out_len - y * chunk; if (rowlen > chunk) rowlen = chunk;
        for (size_t x = 0; x < bw; ++x) {
            size_t bidx = (x + y * bw) * bs;
            if (bidx + bs > data_len) return -1;                                /* slice panic lib.rs:138 */
            uint8_t px[64];
            decompress_block(fmt, data + bidx, px);
            for (size_t py = 0; py < 4; ++py)
                for (size_t pxx = 0; pxx < 4; ++pxx) {
                    size_t sx = 4 * x + pxx, sy = py;
                    if (sx < w && 4 * y + sy < h) {
                        size_t o = 4 * (sx + sy * w);
                        if (o + 4 > rowlen) return -1;                          /* index panic lib.rs:149 */
                        memcpy(out + y * chunk + o, px + 4 * (pxx + 4 * py), 4);
                    }
                }
        }
    }
    return 0;
}

TXO_API void txo_compress_block_masked(int fmt, const uint8_t rgba[64], uint32_t mask, const txo_params *p, uint8_t *out) {
    compress_block_masked(fmt, rgba, mask, p, out, NULL);
}

TXO_API void txo_decompress_block(int fmt, const uint8_t *block, uint8_t out[64]) { decompress_block(fmt, block, out); }

/* Weighted squared error of an encoded colour block in the measure ClusterFit minimises
 * (cluster.rs:213-220 / :346-353 plus the dropped constant sum w*x^2), in fp64.
 * Used for the "differing ClusterFit block is no worse than the reference's" check. */
TXO_API double txo_colour_block_error(int fmt, const uint8_t rgba[64], uint32_t mask, const txo_params *p, const uint8_t block8[8]) {
    colourset s;
    colourset_new(&s, rgba, mask, fmt, p->weigh_colour_by_alpha != 0);
    uint16_t a = (uint16_t)(block8[0] | (block8[1] << 8)), b = (uint16_t)(block8[2] | (block8[3] << 8));
    const float gr[3] = {1.0f / 31.0f, 1.0f / 63.0f, 1.0f / 31.0f};
    double ea[3] = { (double)((float)((a >> 11) & 31) * gr[0]), (double)((float)((a >> 5) & 63) * gr[1]), (double)((float)(a & 31) * gr[2]) };
    double eb[3] = { (double)((float)((b >> 11) & 31) * gr[0]), (double)((float)((b >> 5) & 63) * gr[1]), (double)((float)(b & 31) * gr[2]) };
    int three = (fmt == TXO_BC1) && (a <= b);
    double err = 0.0;
    int seen[16]; memset(seen, 0, sizeof seen);
    for (int i = 0; i < 16; ++i) {
        int j = s.remap[i];
        if (j < 0 || seen[j]) continue;
        seen[j] = 1;
        int idx = (block8[4 + i / 4] >> (2 * (i % 4))) & 3;
        double wa;  /* weight of endpoint a in the code */
        if (three) wa = idx == 0 ? 1.0 : idx == 1 ? 0.0 : 0.5;   /* idx 3 cannot be a valid point */
        else wa = idx == 0 ? 1.0 : idx == 1 ? 0.0 : idx == 2 ? 2.0 / 3.0 : 1.0 / 3.0;
        double px[3] = { s.points[j].x, s.points[j].y, s.points[j].z };
        double e = 0.0;
        for (int k = 0; k < 3; ++k) {
            double c = wa * ea[k] + (1.0 - wa) * eb[k];
            e += (double)p->weights[k] * (c - px[k]) * (c - px[k]);
        }
        err += (double)s.weights[j] * e;
    }
    return err;
}

/* n independent compress_block_masked calls (test convenience; same code path as above) */
TXO_API void txo_compress_blocks(int fmt, const uint8_t *rgba_blocks, const uint32_t *masks, size_t n, const txo_params *p, uint8_t *out) {
    size_t bs = block_size(fmt);
    for (size_t i = 0; i < n; ++i) compress_block_masked(fmt, rgba_blocks + 64 * i, masks[i], p, out + bs * i, NULL);
}

TXO_API void txo_decompress_blocks(int fmt, const uint8_t *blocks, size_t n, uint8_t *out) {
    size_t bs = block_size(fmt);
    for (size_t i = 0; i < n; ++i) decompress_block(fmt, blocks + bs * i, out + 64 * i);
}
