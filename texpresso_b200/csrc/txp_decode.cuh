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

// colourblock.rs:116-169
__device__ __forceinline__ void decode_colour(const uint2 blk, const bool is_bc1, uint32_t px[16]) {
    const uint32_t a = blk.x & 0xFFFFu, b = blk.x >> 16;
    uint32_t codes[4];
    codes[0] = unpack_565(a);
    codes[1] = unpack_565(b);
    uint32_t c2 = 0, c3 = 0;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const uint32_t c = (codes[0] >> (8 * ch)) & 255u, d = (codes[1] >> (8 * ch)) & 255u;
        if (is_bc1 && a <= b) {
            c2 |= ((c + d) / 2u) << (8 * ch);
        } else {
            c2 |= ((2u * c + d) / 3u) << (8 * ch);
            c3 |= ((c + 2u * d) / 3u) << (8 * ch);
        }
    }
    codes[2] = c2 | 0xFF000000u;
    codes[3] = (is_bc1 && a <= b) ? 0u : (c3 | 0xFF000000u);
#pragma unroll
    for (int i = 0; i < 16; ++i) px[i] = codes[(blk.y >> (2 * i)) & 3u];
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

// alpha.rs:53-68
__device__ __forceinline__ void decode_alpha2(const uint2 blk, uint32_t px[16]) {
    const unsigned long long bits = ((unsigned long long)blk.y << 32) | blk.x;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t n = (uint32_t)(bits >> (4 * i)) & 15u;
        px[i] = (px[i] & 0x00FFFFFFu) | ((n | (n << 4)) << 24);
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
