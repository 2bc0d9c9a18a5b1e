// txp_decode.cuh -- BC1..BC5 decoders, one thread per block.
//
// Replaces (reference): lib.rs:124-156 (image loop + bounds-checked scatter), lib.rs:240-277
// (decompress_block), colourblock.rs:97-169, alpha.rs:53-68 (BC2), alpha.rs:258-304 (BC3/4/5).
// HBM-bound: 8/16 B read and 64 B written per block; each thread writes four 16-byte row segments, so a
// warp store covers 512 contiguous bytes of one image row.
#pragma once
#include "txp_common.cuh"

namespace txp {

// colourblock.rs:97-113 : 5:6:5 -> 8:8:8 by bit replication, packed as 0xAABBGGRR with A = 255
__device__ __forceinline__ uint32_t unpack_565(const uint32_t v) {
    const uint32_t r = (v >> 11) & 31u, g = (v >> 5) & 63u, b = v & 31u;
    return ((r << 3) | (r >> 2)) | (((g << 2) | (g >> 4)) << 8) | (((b << 3) | (b >> 2)) << 16) | 0xFF000000u;
}

// 8 bits holding four 2-bit indices -> PRMT selector with the indices in its four nibbles
__device__ __forceinline__ uint32_t spread2to4(const uint32_t x) {
    const uint32_t y = (x | (x << 4)) & 0x0F0Fu;          // two fields per byte
    return (y | (y << 2)) & 0x3333u;                      // one field per nibble
}

// colourblock.rs:116-169.  The four codes are held as channel planes (byte k of plane c = channel c of code k), so
// one PRMT (a 4-entry byte LUT indexed by the selector nibbles) decodes a channel for the four pixels of a row; a
// two-stage PRMT transpose then interleaves the planes into RGBA words.
__device__ __forceinline__ void decode_colour(const uint2 blk, const bool is_bc1, uint32_t px[16]) {
    const uint32_t a = blk.x & 0xFFFFu, b = blk.x >> 16;
    const uint32_t e0 = unpack_565(a), e1 = unpack_565(b);
    const bool three = is_bc1 && a <= b;
    uint32_t plane[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const uint32_t c = (e0 >> (8 * ch)) & 255u, d = (e1 >> (8 * ch)) & 255u;
        const uint32_t c2 = three ? (c + d) / 2u : (2u * c + d) / 3u;
        const uint32_t c3 = three ? 0u : (c + 2u * d) / 3u;
        plane[ch] = c | (d << 8) | (c2 << 16) | (c3 << 24);
    }
    const uint32_t plane_a = three ? 0x00FFFFFFu : 0xFFFFFFFFu;   // alpha 255, except code 3 of the 3-colour mode
#pragma unroll
    for (int row = 0; row < 4; ++row) {
        const uint32_t sel = spread2to4((blk.y >> (8 * row)) & 255u);
        const uint32_t r4 = __byte_perm(plane[0], 0, sel), g4 = __byte_perm(plane[1], 0, sel);
        const uint32_t b4 = __byte_perm(plane[2], 0, sel), a4 = __byte_perm(plane_a, 0, sel);
        const uint32_t rg_lo = __byte_perm(r4, g4, 0x5140), rg_hi = __byte_perm(r4, g4, 0x7362);
        const uint32_t ba_lo = __byte_perm(b4, a4, 0x5140), ba_hi = __byte_perm(b4, a4, 0x7362);
        px[4 * row + 0] = __byte_perm(rg_lo, ba_lo, 0x5410);
        px[4 * row + 1] = __byte_perm(rg_lo, ba_lo, 0x7632);
        px[4 * row + 2] = __byte_perm(rg_hi, ba_hi, 0x5410);
        px[4 * row + 3] = __byte_perm(rg_hi, ba_hi, 0x7632);
    }
}

// 12 bits holding four 3-bit indices -> PRMT selector with the indices in its four nibbles
__device__ __forceinline__ uint32_t spread3to4(const uint32_t x) {
    const uint32_t y = (x & 0x03Fu) | ((x & 0xFC0u) << 2);            // two 6-bit groups in bytes 0 and 1
    return (y & 0x0707u) | ((y & 0x3838u) << 1);                       // 0abc0def per byte
}

// alpha.rs:258-304 : writes the decoded value into byte `channel` of each pixel.
// The 8-entry codebook lives in two registers; PRMT is a byte LUT indexed by the selector nibbles, so one
// PRMT looks up four pixels and one more PRMT per pixel inserts the byte into the pixel word.
__device__ __forceinline__ void decode_alpha3(const uint2 blk, const int channel, uint32_t px[16]) {
    const int a0 = (int)(blk.x & 255u), a1 = (int)((blk.x >> 8) & 255u);
    uint32_t c[8];
    c[0] = (uint32_t)a0; c[1] = (uint32_t)a1;
    if (a0 <= a1) {
#pragma unroll
        for (int i = 1; i < 5; ++i) c[1 + i] = (uint32_t)(((5 - i) * a0 + i * a1) / 5);
        c[6] = 0u; c[7] = 255u;
    } else {
#pragma unroll
        for (int i = 1; i < 7; ++i) c[1 + i] = (uint32_t)(((7 - i) * a0 + i * a1) / 7);
    }
    const uint32_t lut_lo = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
    const uint32_t lut_hi = c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24);
    const unsigned long long bits = ((unsigned long long)blk.y << 16) | (blk.x >> 16);   // 48 index bits
    // selector that replaces byte `channel` of a pixel (first PRMT source) with byte k of the second source
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t four = __byte_perm(lut_lo, lut_hi, spread3to4((uint32_t)(bits >> (12 * q)) & 0xFFFu));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t sel = (0x3210u & ~(0xFu << (4 * channel))) | ((4u + k) << (4 * channel));
            px[4 * q + k] = __byte_perm(px[4 * q + k], four, sel);
        }
    }
}

// alpha.rs:53-68: nibble n -> n | n<<4 (= 17 n), four pixels per step
__device__ __forceinline__ void decode_alpha2(const uint2 blk, uint32_t px[16]) {
#pragma unroll
    for (int row = 0; row < 4; ++row) {
        const uint32_t n4 = ((row < 2 ? blk.x : blk.y) >> (16 * (row & 1))) & 0xFFFFu;
        const uint32_t y = (n4 & 0x00FFu) | ((n4 & 0xFF00u) << 8);            // two nibbles in bytes 0 and 2
        const uint32_t four = ((y | (y << 4)) & 0x0F0F0F0Fu) * 17u;            // one nibble per byte, expanded to 8 bits
#pragma unroll
        for (int k = 0; k < 4; ++k) px[4 * row + k] = __byte_perm(px[4 * row + k], four, 0x0210u | ((4u + k) << 12));
    }
}

// lib.rs:240-277
template <int FMT>
__device__ __forceinline__ void decode_block(const uint8_t* __restrict__ data, const uint64_t b, uint32_t px[16]) {
    if (FMT == BC1) {
        decode_colour(__ldg(reinterpret_cast<const uint2*>(data) + b), true, px);
    } else if (FMT == BC4) {
        const uint2 q = __ldg(reinterpret_cast<const uint2*>(data) + b);
#pragma unroll
        for (int i = 0; i < 16; ++i) px[i] = 0xFF000000u;
        decode_alpha3(q, 0, px);
#pragma unroll
        for (int i = 0; i < 16; ++i) px[i] = __byte_perm(px[i], 0, 0x3000);     // (r, r, r, 255): lib.rs:264-268
    } else {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(data) + b);
        if (FMT == BC5) {
#pragma unroll
            for (int i = 0; i < 16; ++i) px[i] = 0xFF000000u;
            decode_alpha3(make_uint2(q.x, q.y), 0, px);
            decode_alpha3(make_uint2(q.z, q.w), 1, px);
        } else {
            decode_colour(make_uint2(q.z, q.w), false, px);
            if (FMT == BC2) decode_alpha2(make_uint2(q.x, q.y), px);
            else decode_alpha3(make_uint2(q.x, q.y), 3, px);
        }
    }
}

// image mode: scatter into a w x h RGBA8 image (lib.rs:141-153); list mode (w == 0): 64 B per block.
template <int FMT>
__global__ void __launch_bounds__(256) decode_kernel(const uint8_t* __restrict__ data, const uint64_t nblocks,
                                                     const uint32_t w, const uint32_t h, const uint32_t bw,
                                                     const int vec_ok, uint8_t* __restrict__ out) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    uint32_t px[16];
    decode_block<FMT>(data, b, px);
    if (w == 0) {
        uint4* o = reinterpret_cast<uint4*>(out) + b * 4;
#pragma unroll
        for (int r = 0; r < 4; ++r) o[r] = make_uint4(px[4 * r], px[4 * r + 1], px[4 * r + 2], px[4 * r + 3]);
        return;
    }
    const uint32_t b32 = (uint32_t)b;                     // nblocks < 2^31 (checked by the host)
    const uint32_t by = b32 / bw, bx = b32 - by * bw;
    const uint32_t x0 = 4 * bx, y0 = 4 * by;
    if (vec_ok && y0 + 4 <= h) {
        uint8_t* base = out + ((size_t)y0 * w + x0) * 4;
#pragma unroll
        for (int r = 0; r < 4; ++r)
            *reinterpret_cast<uint4*>(base + (size_t)r * w * 4) = make_uint4(px[4 * r], px[4 * r + 1], px[4 * r + 2], px[4 * r + 3]);
    } else {
        uint32_t* img = reinterpret_cast<uint32_t*>(out);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t sx = x0 + (i & 3), sy = y0 + (i >> 2);
            if (sx < w && sy < h) img[(size_t)sy * w + sx] = px[i];
        }
    }
}

}  // namespace txp
