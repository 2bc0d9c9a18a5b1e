// txp_alpha_lattice.cuh -- BC4 / BC5 encoder, closed-form fast path + compacted generic path.
//
// Replaces (reference, /root/reference/lib/src/alpha.rs) compress_bc3 :187-256 as called from lib.rs:200-203.
//
// Why: the literal nearest-code search (8 distance keys + a 7-input minimum per pixel and codebook, txp_alpha.cuh)
// costs ~740 lane-instructions per channel, 2-4x the budget of an HBM-bound kernel.
//
// "Regular" block := fully valid, no 0 and no 255 value, max - min >= 7.  Then min5 == min7 == lo and
// max5 == max7 == hi (alpha.rs:194-212), fix_range (alpha.rs:70-77) changes nothing, the codes 0 / 255 of the 5-point
// book can never win (every value is strictly nearer to lo resp. hi), and both books are the lattices
// lo + floor(i*r/N), i = 0..N (alpha.rs:231, :241).  For such a lattice the reference's first-minimum rule
// (alpha.rs:101-111) is a monotone step function of the pixel value, and for every r = hi - lo in 7..255 there is an
// fp32 pair (a, beta) with   slot(x) == rint((x + beta) * a)   for all x in 0..r  (tools/gen_alpha_lattice.py finds
// the pairs and proves them exhaustively; tests/test_alpha_lattice.py re-checks against the C oracle).  Per pixel and
// book that is ONE packed FADD2 + ONE packed FFMA2 for two pixels:
//     vm   = 1.5*2^15 + v            (PRMT: the byte dropped into mantissa bits 8..15, ulp 2^-8)
//     d    = vm - (1.5*2^15 + lo - beta)                      exact
//     t    = fma(d, a, 1.5*2^23)     -> mantissa low bits = slot
// Four slots are merged into one PRMT selector (t0 + 16 t1 + 256 t2 + 4096 t3, low 16 bits), one PRMT looks up
// four code bytes from the 8-byte codebook register pair, VABSDIFF4 + IDP.4A accumulate the squared error of four pixels.
// The 7-point book is evaluated from the hi end (x = hi - v) because its block is written with swapped end points
// (alpha.rs:167-185, always); with that both books share the slot -> 3-bit index map 0 -> 0, N -> 1, s -> s+1.
//
// Every other block (edge masks, a 0 or 255 present, range < 7) takes the literal path of txp_alpha.cuh.  So that a
// few such blocks per warp do not make the whole warp execute both paths, each warp runs persistently over many
// 32-block tiles, pushes its irregular (block, channel) items on a private shared-memory queue and drains the queue
// 32 items at a time with all lanes busy.
#pragma once
#include "txp_alpha.cuh"

namespace txp {

__device__ uint4 g_alpha_lattice[512];            // TXP_ALPHA_LATTICE (alpha_lattice_data.h), copied at context creation

// tuning knobs (tools/micro/alpha_ab.cu measures them; defaults = best measured)
#ifndef TXP_LAT_ERR
#define TXP_LAT_ERR 0      // 1: sum c^2 - 2 sum c v with two IDP.4A;  0: VABSDIFF4 + IDP.4A
#endif
#ifndef TXP_LAT_VMG
#define TXP_LAT_VMG 0      // G channel -> fp32 mantissa: 0 = PRMT, 1 = LOP3 (two of them: LOP3 takes one immediate), 2 = LOP3 with the magic in a register
#endif
#ifndef TXP_LAT_VMR
#define TXP_LAT_VMR 0      // R channel -> fp32 mantissa: 0 = PRMT, 1 = SHF + LOP3, 2 = LOP3 + IMAD
#endif
#ifndef TXP_LAT_SEMI
#define TXP_LAT_SEMI 1      // 1: one-sided (zeros or 255s present) blocks take the semi-lattice path; 0: literal search
#endif
#ifndef TXP_LAT_MM
#define TXP_LAT_MM 0       // min / max: 0 = VIMNMX3 trees, 1 = two-input VIMNMX
#endif

#ifndef TXP_LAT_SCALAR
#define TXP_LAT_SCALAR 0   // slot arithmetic of the regular path: 0 = packed FADD2 + FFMA2 (two pixels each), 1 = scalar FADD + FFMA
#endif
#ifndef TXP_LAT_PACK
#define TXP_LAT_PACK 0     // index bytes -> 48-bit field: 0 = shift / or / mask (ALU pipe), 1 = multiply-gather (IMAD with immediates, FMA pipe) + masks
#endif
#ifndef TXP_LAT_LIFT
#define TXP_LAT_LIFT 1     // regular path, pixel -> fp32: 1 = IDP.4A (FMA pipe): 1.5*2^16 + v with the byte at mantissa bits 7..14; 0 = PRMT (ALU pipe): 1.5*2^15 + v, bits 8..15
#endif
constexpr uint32_t MAGIC15 = 0x47400000u;         // 1.5 * 2^15 as fp32 bits
constexpr uint32_t MAGIC16 = 0x47C00000u;         // 1.5 * 2^16 as fp32 bits
constexpr int LIFT_SHIFT = TXP_LAT_LIFT ? 7 : 8;  // position of the pixel value inside the lifted word
constexpr uint32_t LIFT_MAGIC = TXP_LAT_LIFT ? MAGIC16 : MAGIC15;
constexpr float MAGIC23 = 12582912.0f;            // 1.5 * 2^23

// PRMT with the hardware selector semantics (bit 3 of a nibble = sign replication).  __byte_perm() promises to ignore
// that bit, so nvcc masks every run-time selector with 0x7777 first; all selectors here have bit 3 clear by construction.
__device__ __forceinline__ uint32_t prmt(const uint32_t a, const uint32_t b, const uint32_t sel) {
    uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d;
}

// One book over 16 pixels.  vm: the pixels as 1.5*2^15 + v; V: the same values as packed bytes (4 per word);
// origin / a: d = vm - origin, slot = rint(d * a) (a < 0 for the 7-point book: x = hi - v);
// clo/chi: the codes of slots 0..7 as bytes.  sel[k] holds the slots of pixels 4k..4k+3 as nibbles.
// Returns sum(c^2) - 2 sum(c v) = sum (v - c)^2 - sum v^2 over the chosen codes c: the sum v^2 is the same for both
// books, so err5 <= err7 (alpha.rs:251) is decided on these values.  (Two IDP.4A on the FMA pipe instead of
// VABSDIFF4 + IDP.4A: the ALU pipe, where PRMT / VIMNMX3 / VABSDIFF4 all run at half rate, is the busier one.)
__device__ __forceinline__ int lattice_book(const uint32_t vm[16], const uint32_t V[4], const float origin, const float a,
                                            const uint32_t clo, const uint32_t chi, uint32_t sel[4]) {
    const f32x2 o2 = pk(origin, origin), a2 = pk(a, a), m2 = pk(MAGIC23, MAGIC23);
    uint32_t cc = 0, cv = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float t0, t1, t2, t3;
#if TXP_LAT_SCALAR
        // scalar FADD + FFMA per pixel: one dispatch slot each, and the ALU pipe keeps issuing next to them; a packed FADD2 / FFMA2 holds the
        // dispatch port for two cycles and serialises with the ALU-pipe instructions around it (tools/micro/half2_rate.cu: FFMA2 / PRMT 1:1
        // issues 1.6 warp-instructions per clock per SM, IMAD / PRMT 1:1 issues 3.9)
        t0 = __fmaf_rn(__fsub_rn(__uint_as_float(vm[4 * k]), origin), a, MAGIC23);
        t1 = __fmaf_rn(__fsub_rn(__uint_as_float(vm[4 * k + 1]), origin), a, MAGIC23);
        t2 = __fmaf_rn(__fsub_rn(__uint_as_float(vm[4 * k + 2]), origin), a, MAGIC23);
        t3 = __fmaf_rn(__fsub_rn(__uint_as_float(vm[4 * k + 3]), origin), a, MAGIC23);
#else
        const f32x2 d01 = sub2(pk(__uint_as_float(vm[4 * k]), __uint_as_float(vm[4 * k + 1])), o2);
        const f32x2 d23 = sub2(pk(__uint_as_float(vm[4 * k + 2]), __uint_as_float(vm[4 * k + 3])), o2);
        upk(fma2(d01, a2, m2), t0, t1);
        upk(fma2(d23, a2, m2), t2, t3);
#endif
        // low 16 bits: four slot nibbles (the magic's low 22 bits are zero, so nothing else reaches them)
        const uint32_t s = ((__float_as_uint(t3) * 16u + __float_as_uint(t2)) * 16u + __float_as_uint(t1)) * 16u + __float_as_uint(t0);
        sel[k] = s;
        const uint32_t c4 = prmt(clo, chi, s);
#if TXP_LAT_ERR
        cc = __dp4a(c4, c4, cc);
        cv = __dp4a(c4, V[k], cv);
#else
        const uint32_t e4 = __vabsdiffu4(c4, V[k]);
        cc = __dp4a(e4, e4, cc);
#endif
    }
    return (int)cc - 2 * (int)cv;
}

// four byte-sized 3-bit indices per word -> 48-bit index field + end points (alpha.rs:121-144)
__device__ __forceinline__ uint2 pack_alpha_block(const uint32_t a0, const uint32_t a1, const uint32_t w[4]) {
#if TXP_LAT_PACK
    // Multiply-gather: the index bytes (i0, i1, i2, i3) of a word occupy bits 0-2, 8-10, 16-18, 24-26, so w * 33 = w | w << 5 puts
    // i0 next to i1 (bits 5-10) and i2 next to i3 (bits 21-26); after masking, y * 1025 = y | y << 10 puts the two 6-bit halves
    // next to each other (bits 15-26 = the 12-bit field z of four pixels).  A further power of two in the second multiplier
    // places z where the output word wants it, so the only shifts left are the two right shifts.  Products of disjoint bit
    // patterns: no carries, the multiplications are exact ORs and run on the FMA pipe (IMAD with an immediate).
    uint32_t y[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) y[k] = (w[k] * 33u) & 0x07E007E0u;
    const uint32_t u0 = y[0] * (1025u << 1);          // z0 at bits 16-27
    const uint32_t u1a = y[1] * (1025u << 13);        // low 4 bits of z1 at bits 28-31
    const uint32_t u1b = (y[1] * 1025u) >> 19;        // z1 >> 4 at bits 0-7 (bit 12: a stray copy, masked below)
    const uint32_t u2 = ((y[2] * 1025u) & 0x07FF8000u) >> 7;   // z2 at bits 8-19
    const uint32_t u3 = y[3] * (1025u << 5);          // z3 at bits 20-31
    const uint32_t x = (u1a & 0xF0000000u) | ((u0 & 0x0FFF0000u) | (a0 | (a1 << 8)));
    const uint32_t yy = (u3 & 0xFFF00000u) | ((u1b & 0xFFu) | u2);
    return make_uint2(x, yy);
#else
    uint32_t z[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t y = (w[k] | (w[k] >> 5)) & 0x003F003Fu;
        z[k] = (y | (y >> 10)) & 0xFFFu;
    }
    const uint32_t g0 = z[0] | (z[1] << 12), g1 = z[2] | (z[3] << 12);
    return make_uint2(a0 | (a1 << 8) | (g0 << 16), (g0 >> 16) | (g1 << 8));
#endif
}

// Flat and narrow-range blocks (what flat regions of real textures are made of), also in closed form:
//  * all 0 / all 255: constant encodings (min5/max5 of empty sets, alpha.rs:215-224): 00 05 00.. / 00 05 FF..
//  * max - min <= 6, no 0 / 255, min <= 248: fix_range (alpha.rs:70-77) widens the ranges to (lo, lo + max(r, 5)) and
//    (lo, lo + 7); the 7-point book then holds every integer lo..lo+7 (err7 == 0) and the 5-point book lo..lo+5, or
//    lo + {0,1,2,3,4,6} when r == 6, so err5 == 0 unless r == 6 and some pixel sits at lo + 5 -- the only case in which
//    the 7-point block is written (alpha.rs:251).  Index = a byte LUT of x = v - lo.  tests/test_alpha_lattice.py
//    checks this restatement against the oracle for every lo and r.
// Everything else (0 / 255 mixed with other values, min > 248) takes the literal search.
__device__ __forceinline__ bool alpha_is_narrow(const uint32_t lo, const uint32_t hi) {
    return hi == 0u || lo == 255u || (lo >= 1u && hi <= 254u && hi - lo <= 6u && lo <= 248u);
}
__device__ __forceinline__ uint2 alpha_fit_narrow(const uint32_t lo, const uint32_t hi, const uint32_t V[4]) {
    if (hi == 0u) return make_uint2(0x00000500u, 0u);
    if (lo == 255u) return make_uint2(0xFFFF0500u, 0xFFFFFFFFu);
    const uint32_t r = hi - lo;
    uint32_t sel[4], any5 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t x = V[k] - lo * 0x01010101u;                       // x = v - lo per byte, 0..6
        const uint32_t n = x | (x >> 4);                                  // byte 0: x0 | x1 << 4, byte 2: x2 | x3 << 4
        sel[k] = prmt(n, n, 0x0020u);
        any5 |= prmt(0u, 0x00000100u, sel[k]);                            // 1 where x == 5
    }
    const bool seven = r == 6u && any5 != 0u;
    const uint32_t mlo = seven ? 0x05060701u : 0x04030200u, mhi = seven ? 0x00000304u : 0x00010105u;
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = prmt(mlo, mhi, sel[k]);
    return pack_alpha_block(seven ? lo + 7u : lo, seven ? lo : (r < 5u ? lo + 5u : lo + r), w);
}

// One channel of one fully valid block, given as vm[i] = 1.5*2^15 + v_i (fp32 bits) and V = the same 16 values as
// packed bytes.  Returns true (out written) for a regular block, false for one that goes to the queue.
__device__ __forceinline__ bool alpha_fit_lattice(const uint32_t vm[16], const uint32_t V[4], const uint4* __restrict__ tab, uint2& out) {
    // min / max on the raw bits (positive floats order like integers)
#if TXP_LAT_MM == 2
    // Both at once: with q = v << 7 (vm = 1.5 * 2^16 + v as bits = MAGIC16 + q) one IMAD (FMA pipe) forms the half-word pair
    // (q, 32640 - q):  vm * (1 - 2^16) + C = q + ((32640 - q) << 16)  (mod 2^32; MAGIC16 << 16 == 0), and ONE unsigned 16x2 minimum
    // tree yields min q in the low half and 32640 - max q in the high half: 8 ALU-pipe instructions instead of 16.
    static_assert(TXP_LAT_LIFT == 1, "TXP_LAT_MM == 2 needs the IDP.4A lift");
    uint32_t pq[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pq[i] = vm[i] * 0xFFFF0001u + ((32640u << 16) - MAGIC16);
    uint32_t m2 = __vimin3_u16x2(pq[0], pq[1], pq[2]);
#pragma unroll
    for (int i = 3; i < 15; i += 2) m2 = __vimin3_u16x2(m2, pq[i], pq[i + 1]);
    m2 = __vminu2(m2, pq[15]);
    const uint32_t qmin = m2 & 0xFFFFu, hq = m2 >> 16;                 // hq = 32640 - q_max
    const uint32_t span = 32640u - hq - qmin;
    if (!(qmin != 0u && hq != 0u && span >= (7u << 7))) return false;
    const uint32_t mn = MAGIC16 + qmin, mx = (MAGIC16 + 32640u) - hq;
#else
#if TXP_LAT_MM
    uint32_t mn = min(vm[0], vm[1]), mx = max(vm[0], vm[1]);
#pragma unroll
    for (int i = 2; i < 16; ++i) { mn = min(mn, vm[i]); mx = max(mx, vm[i]); }
#else
    uint32_t mn = __vimin3_u32(vm[0], vm[1], vm[2]), mx = __vimax3_u32(vm[0], vm[1], vm[2]);
#pragma unroll
    for (int i = 3; i < 15; i += 2) { mn = __vimin3_u32(mn, vm[i], vm[i + 1]); mx = __vimax3_u32(mx, vm[i], vm[i + 1]); }
    mn = min(mn, vm[15]); mx = max(mx, vm[15]);
#endif
    const uint32_t span = mx - mn;                                     // r << LIFT_SHIFT
    if (!(mn > LIFT_MAGIC && mx < (LIFT_MAGIC | (255u << LIFT_SHIFT)) && span >= (7u << LIFT_SHIFT))) return false;
#endif

    const uint4 e5 = tab[span >> (LIFT_SHIFT - 1)], e7 = tab[(span >> (LIFT_SHIFT - 1)) + 1];        // row r = two uint4
    const uint32_t lo = (mn >> LIFT_SHIFT) & 255u, hi = (mx >> LIFT_SHIFT) & 255u;
    uint32_t s5[4], s7[4];
    // 5-point book from lo: x = v - lo, origin = lo - beta5, codes lo + offs
    const int err5 = lattice_book(vm, V, __fsub_rn(__uint_as_float(mn), __uint_as_float(e5.y)), __uint_as_float(e5.x),
                                  e5.z + lo * 0x01010101u, e5.w + lo * 0x01010101u, s5);
    // 7-point book from hi: x = hi - v = -(v - (hi + beta7)), codes hi - offs
    const int err7 = lattice_book(vm, V, __fadd_rn(__uint_as_float(mx), __uint_as_float(e7.y)), -__uint_as_float(e7.x),
                                  hi * 0x01010101u - e7.z, hi * 0x01010101u - e7.w, s7);
    const bool five = err5 <= err7;                                    // alpha.rs:251
    const uint32_t a0 = five ? lo : hi, a1 = five ? hi : lo;          // write_alpha_block5 as is / write_alpha_block7 swapped
    const uint32_t mhi = five ? 0x00000105u : 0x01070605u;            // slot -> index: 0 -> 0, N -> 1, s -> s + 1
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = prmt(0x04030200u, mhi, five ? s5[k] : s7[k]);   // four 3-bit indices, one per byte
    out = pack_alpha_block(a0, a1, w);
    return true;
}

// Both channel fits of one fully valid block.  R (byte 0) and G (byte 1) are lifted into the fp32 mantissa with one
// PRMT each; the packed-byte forms of both channels (VR, VG: also what an irregular channel is queued as) share
// their first PRMT level.
template <int FMT>
__device__ __forceinline__ void alpha_fit_block(const uint32_t px[16], const uint4* __restrict__ tab, bool& ok0, uint2& r0, uint32_t VR[4],
                                                bool& ok1, uint2& r1, uint32_t VG[4]) {
    uint32_t vm[16];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (FMT == BC4) {
            VR[k] = prmt(prmt(px[4 * k], px[4 * k + 1], 0x0040u), prmt(px[4 * k + 2], px[4 * k + 3], 0x0040u), 0x5410u);
        } else {
            const uint32_t t0 = prmt(px[4 * k], px[4 * k + 1], 0x5140u), t1 = prmt(px[4 * k + 2], px[4 * k + 3], 0x5140u);   // (R, R', G, G')
            VR[k] = prmt(t0, t1, 0x5410u);
            VG[k] = prmt(t0, t1, 0x7632u);
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
#if TXP_LAT_LIFT
        vm[i] = __dp4a(px[i], 0x00000080u, MAGIC16);                                 // 1.5 * 2^16 + R, as an integer add on the FMA pipe
#elif TXP_LAT_VMR == 0
        vm[i] = prmt(MAGIC15, px[i], 0x3240u);                                       // bytes (0x00, R, 0x40, 0x47)
#elif TXP_LAT_VMR == 1
        vm[i] = ((px[i] << 8) & 0xFF00u) | MAGIC15;
#else
        vm[i] = (px[i] & 0xFFu) * 256u + MAGIC15;
#endif
    }
    ok0 = alpha_fit_lattice(vm, VR, tab, r0);
    ok1 = true;
    if (FMT == BC5) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
#if TXP_LAT_LIFT
            vm[i] = __dp4a(px[i], 0x00008000u, MAGIC16);
#elif TXP_LAT_VMG == 0
            vm[i] = prmt(MAGIC15, px[i], 0x3250u);
#elif TXP_LAT_VMG == 1
            vm[i] = (px[i] & 0xFF00u) | MAGIC15;
#else
            uint32_t magic; asm volatile("mov.u32 %0, 0x47400000;" : "=r"(magic));
            vm[i] = (px[i] & 0xFF00u) | magic;
#endif
        }
        ok1 = alpha_fit_lattice(vm, VG, tab, r1);
    }
}

// ---- per-warp queues of irregular (block, channel) items -----------------------------------------------------------
// meta = block | RELOAD << 30 | channel << 31.  Fully valid blocks carry their 16 channel values as packed bytes, so
// no path needs a second trip to memory; partial (edge) blocks are re-gathered with their mask (RELOAD).
// Two levels, so that the main loop pays for one ballot per channel only:
//   queue A   <- every irregular item of a tile (ballot-compacted).  Drained 32 at a time: flat / narrow-range items
//                are finished on the spot in closed form (~80 instructions), the others are compacted once more into
//   queue Z/F <- "one-sided" items: zeros (or 255s) next to ordinary values -> semi-lattice path (alpha_semi_item);
//   queue L   <- everything else -> the 650-instruction literal search (alpha_literal_item).
// Every second-level queue is drained 32 items at a time, so each path always runs with full warps.
constexpr int QUEUE_A = 96;                       // <= 31 left over + 2 x 32 new items per tile
constexpr int QUEUE_B = 64;                       // <= 31 left over + 32 from one drain step
constexpr int QZ = 0, QF = 1, QL = 2;
constexpr uint32_t ITEM_RELOAD = 0x40000000u, ITEM_BLOCK = 0x3FFFFFFFu;
struct WarpQueue {
    uint4 a_vals[QUEUE_A];
    uint4 b_vals[3][QUEUE_B];
    uint32_t a_meta[QUEUE_A];
    uint32_t b_meta[3][QUEUE_B];
};

template <int FMT>
__device__ __forceinline__ void store_item(uint8_t* __restrict__ out, const uint32_t meta, const uint2 r) {
    const uint32_t b = meta & ITEM_BLOCK, ch = meta >> 31;
    reinterpret_cast<uint2*>(out)[FMT == BC4 ? (size_t)b : 2 * (size_t)b + ch] = r;
}

template <int FMT>
__device__ __noinline__ void alpha_literal_item(const BlockSource& src, uint8_t* __restrict__ out, const uint32_t meta, const uint4 V) {
#ifdef TXP_LAT_NODRAIN            // measurement only (tools/micro/alpha_ab.cu): cost of the literal path
    return;
#endif
    uint32_t v[16];
    uint2 r;
    if (meta & ITEM_RELOAD) {
        const uint32_t b = meta & ITEM_BLOCK, ch = meta >> 31;
        uint32_t px[16], mask;
        load_block_thread(src, b, px, mask);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (px[i] >> (8 * ch)) & 255u;    // lib.rs:200, :202-203
        r = alpha_fit_thread(v, mask);
    } else {
        const uint32_t w[4] = {V.x, V.y, V.z, V.w};
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (w[i >> 2] >> (8 * (i & 3))) & 255u;
        r = alpha_fit_full(v);
    }
    store_item<FMT>(out, meta, r);
}

// ---- semi-lattice path: zeros (ZSIDE) or 255s (!ZSIDE) next to ordinary values ---------------------------------------
// With m / M the smallest / largest value that is neither 0 nor 255 and M - m >= 7 (no fix_range, alpha.rs:70-77):
//  * 5-point book (alpha.rs:227-234) = lattice (m, M - m) + the codes 0 and 255, which are exact for the special pixels:
//    those pixels are moved to the centres of the unused slots 6 / 7, whose codebook bytes are 0 / 255 (indices 6 / 7).
//  * 7-point book (alpha.rs:237-242, quirk Q1) = the interpolants of the lattice (0, M) resp. (m, 255) plus E0 = m, E1 = M.
//    One end is regular; at the other the end code (0 resp. 255) is replaced by E0 resp. E1.  E0 wins exactly the values
//    <= B0 = floor((m + c_up) / 2), c_up = the smallest interpolant > m (E0 has index 0: ties go to it), and the zero
//    pixels too if m <= c_1; E1 wins the values >= A1 = ceil((M + c_dn) / 2), c_dn = the largest interpolant < M, and the
//    255 pixels if M >= c_6.  Those pixels are moved to the end of the lattice, whose codebook byte is patched to E0 / E1.
// Blocks that fail the side conditions (M - m < 7, m > c_1, M < c_6) return false -> literal queue.
// tests/test_alpha_lattice.py restates this in numpy and checks it against the oracle (random + swept corpora).
template <bool ZSIDE>
__device__ __forceinline__ bool alpha_fit_semi(const uint32_t w[4], const uint4* __restrict__ tab, uint2& out) {
    uint32_t vm[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) vm[i] = prmt(MAGIC15, w[i >> 2], 0x3240u + 0x10u * (i & 3));
    constexpr uint32_t K255 = MAGIC15 | 0xFF00u;
    // m / M over the ordinary values: the special value wraps to the top of the unsigned range and never wins the min
    uint32_t t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = ZSIDE ? vm[i] - (MAGIC15 + 0x100u) : (MAGIC15 | 0xFE00u) - vm[i];
    uint32_t tmin = __vimin3_u32(t[0], t[1], t[2]);
    uint32_t other = ZSIDE ? __vimax3_u32(vm[0], vm[1], vm[2]) : __vimin3_u32(vm[0], vm[1], vm[2]);
#pragma unroll
    for (int i = 3; i < 15; i += 2) {
        tmin = __vimin3_u32(tmin, t[i], t[i + 1]);
        other = ZSIDE ? __vimax3_u32(other, vm[i], vm[i + 1]) : __vimin3_u32(other, vm[i], vm[i + 1]);
    }
    tmin = min(tmin, t[15]);
    other = ZSIDE ? max(other, vm[15]) : min(other, vm[15]);
    const uint32_t m_vm = ZSIDE ? tmin + (MAGIC15 + 0x100u) : other, M_vm = ZSIDE ? other : (MAGIC15 | 0xFE00u) - tmin;
    const uint32_t m = (m_vm >> 8) & 255u, M = (M_vm >> 8) & 255u;
    const uint32_t r5 = M - m;
    const uint32_t lo7 = ZSIDE ? 0u : m, hi7 = ZSIDE ? M : 255u, r7 = hi7 - lo7;
    if (M_vm < m_vm + (7u << 8)) return false;
    if (ZSIDE ? (7u * m > M) : (M < m + ((6u * (255u - m) * 9363u) >> 16))) return false;      // m <= c_1  /  M >= c_6

    const uint4 e5 = tab[2 * r5], e7 = tab[2 * r7 + 1];
    uint32_t s5[4], s7[4], vx[16];
    // ---- 5-point book
    const float a5 = __uint_as_float(e5.x), origin5 = __fsub_rn(__uint_as_float(m_vm), __uint_as_float(e5.y));
    const float xs = __fadd_rn(origin5, __fdividef(ZSIDE ? 6.0f : 7.0f, a5));        // centre of slot 6 / 7
#pragma unroll
    for (int i = 0; i < 16; ++i) vx[i] = vm[i] == (ZSIDE ? MAGIC15 : K255) ? __float_as_uint(xs) : vm[i];
    const int err5 = lattice_book(vx, w, origin5, a5, e5.z + m * 0x01010101u, ((e5.w + m * 0x01010101u) & 0x0000FFFFu) | 0xFF000000u, s5);
    // ---- 7-point book, from the hi end
    uint32_t clo = hi7 * 0x01010101u - e7.z, chi = hi7 * 0x01010101u - e7.w;       // slot 0 = hi7, 1..6 = c_6..c_1, 7 = lo7
    if (ZSIDE) {
        const uint32_t c1 = (chi >> 16) & 255u, c2 = (chi >> 8) & 255u;
        const uint32_t b0_vm = MAGIC15 | (((m + (c1 > m ? c1 : c2)) >> 1) << 8);
        chi = (chi & 0x00FFFFFFu) | (m << 24);
#pragma unroll
        for (int i = 0; i < 16; ++i) vx[i] = vm[i] <= b0_vm ? MAGIC15 : vm[i];
    } else {
        const uint32_t c6 = (clo >> 8) & 255u, c5 = (clo >> 16) & 255u;
        const uint32_t a1_vm = MAGIC15 | (((M + (c6 < M ? c6 : c5) + 1u) >> 1) << 8);
        clo = (clo & 0xFFFFFF00u) | M;
#pragma unroll
        for (int i = 0; i < 16; ++i) vx[i] = vm[i] >= a1_vm ? K255 : vm[i];
    }
    const float origin7 = __fadd_rn(__uint_as_float(MAGIC15 | (hi7 << 8)), __uint_as_float(e7.y));
    const int err7 = lattice_book(vx, w, origin7, -__uint_as_float(e7.x), clo, chi, s7);
    const bool five = err5 <= err7;                                    // alpha.rs:251
    const uint32_t mhi = five ? 0x07060105u : 0x01070605u;
    uint32_t iw[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) iw[k] = prmt(0x04030200u, mhi, five ? s5[k] : s7[k]);
    out = pack_alpha_block(five ? m : hi7, five ? M : lo7, iw);
    return true;
}

template <int FMT, bool ZSIDE>
__device__ __noinline__ bool alpha_semi_item(uint8_t* __restrict__ out, const uint4* __restrict__ tab, const uint32_t meta, const uint4 V) {
    const uint32_t w[4] = {V.x, V.y, V.z, V.w};
    uint2 r;
    if (!alpha_fit_semi<ZSIDE>(w, tab, r)) return false;
    store_item<FMT>(out, meta, r);
    return true;
}

// compact the lanes with `take` into second-level queue `which`
__device__ __forceinline__ void queue_b_push(WarpQueue& q, uint32_t (&qb)[3], const int which, const uint32_t lane, const bool take,
                                             const uint32_t meta, const uint4 V) {
    const uint32_t mk = __ballot_sync(FULL, take);
    if (take) { const uint32_t i = qb[which] + __popc(mk & ((1u << lane) - 1u)); q.b_meta[which][i] = meta; q.b_vals[which][i] = V; }
    qb[which] += __popc(mk);
    __syncwarp();
}

template <int FMT>
__device__ __forceinline__ void drain_literal(WarpQueue& q, uint32_t (&qb)[3], const uint32_t lane, const BlockSource& src, uint8_t* __restrict__ out,
                                              const bool flush) {
#pragma unroll 1
    while (qb[QL] >= 32 || (flush && qb[QL] > 0)) {
        const uint32_t n = min(qb[QL], 32u);
        qb[QL] -= n;
        const bool has = lane < n;
        const uint32_t m = q.b_meta[QL][qb[QL] + (has ? lane : 0u)];
        const uint4 v = q.b_vals[QL][qb[QL] + (has ? lane : 0u)];
        __syncwarp();
        if (has) alpha_literal_item<FMT>(src, out, m, v);
    }
}

template <int FMT, bool ZSIDE>
__device__ __forceinline__ void drain_semi(WarpQueue& q, uint32_t (&qb)[3], const uint32_t lane, const BlockSource& src, uint8_t* __restrict__ out,
                                           const uint4* __restrict__ tab, const bool flush) {
    constexpr int W = ZSIDE ? QZ : QF;
#pragma unroll 1
    while (qb[W] >= 32 || (flush && qb[W] > 0)) {
        const uint32_t n = min(qb[W], 32u);
        qb[W] -= n;
        const bool has = lane < n;
        const uint32_t m = q.b_meta[W][qb[W] + (has ? lane : 0u)];
        const uint4 v = q.b_vals[W][qb[W] + (has ? lane : 0u)];
        __syncwarp();
        const bool done = has && alpha_semi_item<FMT, ZSIDE>(out, tab, m, v);
        queue_b_push(q, qb, QL, lane, has && !done, m, v);                // side conditions failed -> literal search
        drain_literal<FMT>(q, qb, lane, src, out, false);
    }
}

// one drain step of queue A for this lane's item (has == false: no item): closed form or hand-over to a second-level queue
template <int FMT>
__device__ __forceinline__ void drain_a_step(WarpQueue& q, uint32_t (&qb)[3], const uint32_t lane, const BlockSource& src, uint8_t* __restrict__ out,
                                             const uint4* __restrict__ tab, const bool has, const uint32_t meta, const uint4 V) {
    // min / max of the 16 packed bytes as two 16-bit lanes per word (VIMNMX3.U16x2)
    const uint32_t w[4] = {V.x, V.y, V.z, V.w};
    const uint32_t e0 = w[0] & 0x00FF00FFu, e1 = w[1] & 0x00FF00FFu, e2 = w[2] & 0x00FF00FFu, e3 = w[3] & 0x00FF00FFu;
    const uint32_t o0 = (w[0] >> 8) & 0x00FF00FFu, o1 = (w[1] >> 8) & 0x00FF00FFu, o2 = (w[2] >> 8) & 0x00FF00FFu, o3 = (w[3] >> 8) & 0x00FF00FFu;
    const uint32_t mn2 = __vimin3_u16x2(__vimin3_u16x2(e0, e1, e2), __vimin3_u16x2(e3, o0, o1), __vminu2(o2, o3));
    const uint32_t mx2 = __vimax3_u16x2(__vimax3_u16x2(e0, e1, e2), __vimax3_u16x2(e3, o0, o1), __vmaxu2(o2, o3));
    const uint32_t lo = min(mn2 & 0xFFFFu, mn2 >> 16), hi = max(mx2 & 0xFFFFu, mx2 >> 16);
    const bool full = has && !(meta & ITEM_RELOAD);
    const bool narrow = full && alpha_is_narrow(lo, hi);
    if (narrow) store_item<FMT>(out, meta, alpha_fit_narrow(lo, hi, w));
#if TXP_LAT_SEMI
    const bool zs = full && !narrow && lo == 0u && hi < 255u, fs = full && !narrow && hi == 255u && lo > 0u;
#else
    const bool zs = false, fs = false;
#endif
    queue_b_push(q, qb, QL, lane, has && !narrow && !zs && !fs, meta, V);
    drain_literal<FMT>(q, qb, lane, src, out, false);
#if TXP_LAT_SEMI
    queue_b_push(q, qb, QZ, lane, zs, meta, V);
    drain_semi<FMT, true>(q, qb, lane, src, out, tab, false);
    queue_b_push(q, qb, QF, lane, fs, meta, V);
    drain_semi<FMT, false>(q, qb, lane, src, out, tab, false);
#endif
}

// push this tile's irregular items, then drain queue A while it holds a full warp of items
template <int FMT>
__device__ __forceinline__ void queue_push_drain(WarpQueue& q, uint32_t& qa, uint32_t (&qb)[3], const uint32_t lane, const BlockSource& src,
                                                 uint8_t* __restrict__ out, const uint4* __restrict__ tab, const uint32_t b_reload,
                                                 const bool todo0, const uint32_t V0[4], const bool todo1, const uint32_t V1[4]) {
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t m0 = __ballot_sync(FULL, todo0);
    if (todo0) { const uint32_t i = qa + __popc(m0 & lt); q.a_meta[i] = b_reload; q.a_vals[i] = make_uint4(V0[0], V0[1], V0[2], V0[3]); }
    qa += __popc(m0);
    if (FMT == BC5) {
        const uint32_t m1 = __ballot_sync(FULL, todo1);
        if (todo1) { const uint32_t i = qa + __popc(m1 & lt); q.a_meta[i] = b_reload | 0x80000000u; q.a_vals[i] = make_uint4(V1[0], V1[1], V1[2], V1[3]); }
        qa += __popc(m1);
    }
    __syncwarp();
#pragma unroll 1
    while (qa >= 32) {
        qa -= 32;
        const uint32_t i = qa + lane;
        const uint32_t meta = q.a_meta[i];
        const uint4 V = q.a_vals[i];
        __syncwarp();
        drain_a_step<FMT>(q, qb, lane, src, out, tab, true, meta, V);
    }
}

template <int FMT>
__device__ __forceinline__ void queue_flush(WarpQueue& q, const uint32_t qa, uint32_t (&qb)[3], const uint32_t lane, const BlockSource& src,
                                            uint8_t* __restrict__ out, const uint4* __restrict__ tab) {
    const bool has = lane < qa;
    const uint32_t i = has ? lane : 0u;
    drain_a_step<FMT>(q, qb, lane, src, out, tab, has, q.a_meta[i], q.a_vals[i]);
#if TXP_LAT_SEMI
    drain_semi<FMT, true>(q, qb, lane, src, out, tab, true);
    drain_semi<FMT, false>(q, qb, lane, src, out, tab, true);
#endif
    drain_literal<FMT>(q, qb, lane, src, out, true);
}

// the per-block body shared by both kernels: px -> outputs for regular channels, queue items for the rest
template <int FMT>
__device__ __forceinline__ void alpha_block_body(const uint32_t px[16], const uint4* __restrict__ tab, uint8_t* __restrict__ out, const uint32_t b,
                                                 bool& todo0, uint32_t VR[4], bool& todo1, uint32_t VG[4]) {
    uint2 r0, r1;
    bool ok0, ok1;
    alpha_fit_block<FMT>(px, tab, ok0, r0, VR, ok1, r1, VG);
    todo0 = !ok0; todo1 = !ok1;
    uint2* o2 = reinterpret_cast<uint2*>(out) + (FMT == BC4 ? (size_t)b : 2 * (size_t)b);
    if (ok0) o2[0] = r0;
    if (FMT == BC5 && ok1) o2[1] = r1;
}

// ---- general kernel: list mode, mip chains, unaligned widths (direct loads) -----------------------------------------
template <int THREADS>
constexpr size_t lattice_smem() { return 512 * sizeof(uint4) + (THREADS / 32) * sizeof(WarpQueue); }

template <int FMT, int THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) alpha_lattice_kernel(const __grid_constant__ BlockSource src, uint8_t* __restrict__ out, const uint32_t ntiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint4* tab = reinterpret_cast<uint4*>(smem_raw);
    WarpQueue* queues = reinterpret_cast<WarpQueue*>(tab + 512);
    for (int i = threadIdx.x; i < 512; i += THREADS) tab[i] = g_alpha_lattice[i];
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpQueue& q = queues[warp];
    uint32_t qa = 0, qb[3] = {0, 0, 0};
    const uint32_t stride = gridDim.x * (THREADS / 32);
#pragma unroll 1
    for (uint32_t tile = blockIdx.x * (THREADS / 32) + warp; tile < ntiles; tile += stride) {
        const uint32_t b = tile * 32 + lane;
        bool todo0 = false, todo1 = false;
        uint32_t reload = 0, VR[4] = {0, 0, 0, 0}, VG[4] = {0, 0, 0, 0};
        if (b < src.nblocks) {
            uint32_t px[16], mask;
            load_block_thread(src, b, px, mask);
            if (mask == 0xFFFFu) {
                alpha_block_body<FMT>(px, tab, out, b, todo0, VR, todo1, VG);
            } else {
                todo0 = true; todo1 = FMT == BC5; reload = ITEM_RELOAD;
            }
        }
        queue_push_drain<FMT>(q, qa, qb, lane, src, out, tab, b | reload, todo0, VR, todo1, VG);
    }
    queue_flush<FMT>(q, qa, qb, lane, src, out, tab);
}

// ---- image-mode kernel: same algorithm, block rows staged through shared memory with cp.async ----------------------
// For a plain w x h image with 16-byte aligned rows (BlockSource::vec_ok, one level) each lane prefetches the four
// 16-byte row segments of its block STAGES-1 tiles ahead into its own 64 bytes of a per-warp ring buffer
// (LDGSTS, no registers held across the wait), and block coordinates advance incrementally (step_q/step_r =
// quotient / remainder of the per-iteration block stride by the blocks-per-row count, computed by the host)
// instead of a division per block.  A lane only ever reads the bytes it copied itself, so cp.async.wait_group
// is the only synchronisation.  Partial bottom rows (h % 4 != 0) go to the literal path through the queue.
__device__ __forceinline__ void cp_async16(const uint32_t smem_addr, const void* gptr) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int THREADS, int STAGES>
constexpr size_t lattice_image_smem() { return 512 * sizeof(uint4) + (size_t)THREADS * STAGES * 64 + (THREADS / 32) * sizeof(WarpQueue); }

template <int FMT, int THREADS, int MIN_CTAS, int STAGES>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) alpha_lattice_image_kernel(const __grid_constant__ BlockSource src, uint8_t* __restrict__ out,
                                                                               const uint32_t ntiles, const uint32_t step_q, const uint32_t step_r) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint4* tab = reinterpret_cast<uint4*>(smem_raw);
    uint4* ring = tab + 512;                                                          // [warp][stage][row][lane]
    WarpQueue* queues = reinterpret_cast<WarpQueue*>(ring + THREADS * STAGES * 4);
    for (int i = threadIdx.x; i < 512; i += THREADS) tab[i] = g_alpha_lattice[i];
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpQueue& q = queues[warp];
    uint32_t qa = 0, qb[3] = {0, 0, 0};
    const uint32_t stride = gridDim.x * (THREADS / 32);
    const uint32_t bw = src.bw, nblocks = (uint32_t)src.nblocks, full_rows = src.h >> 2;   // block rows with 4 pixel rows
    const uint32_t pitch = src.w * 4;                                                 // image bytes < 2^32 (checked by the host)
    uint4* my = ring + (warp * STAGES * 4) * 32 + lane;                               // + (stage * 4 + row) * 32
    const uint32_t my_s = (uint32_t)__cvta_generic_to_shared(my);
    const uint32_t stride32 = stride * 32;
    const uint32_t step_off = step_q * 4 * pitch + step_r * 16, wrap_off = 4 * pitch - bw * 16;

    // prefetch stream: block index, block coordinates and byte offset of the block STAGES-1 tiles ahead
    const uint32_t tile0 = blockIdx.x * (THREADS / 32) + warp;
    uint32_t pb = tile0 * 32 + lane, pby = pb / bw, pbx = pb - pby * bw, poff = pby * 4 * pitch + pbx * 16, pst = 0;
    uint32_t staged = 0;                                                              // bit k: the tile issued k prefetches ago was staged
    auto prefetch = [&]() {
        const bool ok = pb < nblocks && pby < full_rows;
        if (ok) {
            const uint8_t* g = src.rgba + poff;
            const uint32_t s = my_s + pst * (4 * 32 * 16);
#pragma unroll
            for (int r = 0; r < 4; ++r) cp_async16(s + r * (32 * 16), g + r * pitch);
        }
        cp_async_commit();
        staged = (staged << 1) | (ok ? 1u : 0u);
        pb += stride32; pbx += step_r; pby += step_q; poff += step_off;
        if (pbx >= bw) { pbx -= bw; ++pby; poff += wrap_off; }
        pst = pst + 1 == STAGES ? 0 : pst + 1;
    };
#pragma unroll
    for (int i = 0; i < STAGES - 1; ++i) prefetch();

    uint32_t st = 0;
#pragma unroll 1
    for (uint32_t tile = tile0; tile < ntiles; tile += stride) {
        prefetch();
        cp_async_wait<STAGES - 1>();
        const uint32_t b = pb - STAGES * stride32;                                    // this tile's block of this lane
        bool todo0 = false, todo1 = false;
        uint32_t reload = 0, VR[4] = {0, 0, 0, 0}, VG[4] = {0, 0, 0, 0};
        if ((staged >> (STAGES - 1)) & 1u) {
            uint32_t px[16];
            const uint4* sp = my + st * (4 * 32);
#pragma unroll
            for (int r = 0; r < 4; ++r) { const uint4 v = sp[r * 32]; px[4 * r] = v.x; px[4 * r + 1] = v.y; px[4 * r + 2] = v.z; px[4 * r + 3] = v.w; }
            alpha_block_body<FMT>(px, tab, out, b, todo0, VR, todo1, VG);
        } else if (b < nblocks) {                                                     // partial bottom row
            todo0 = true; todo1 = FMT == BC5; reload = ITEM_RELOAD;
        }
        queue_push_drain<FMT>(q, qa, qb, lane, src, out, tab, b | reload, todo0, VR, todo1, VG);
        st = st + 1 == STAGES ? 0 : st + 1;
    }
    cp_async_wait<0>();
    queue_flush<FMT>(q, qa, qb, lane, src, out, tab);
}

// ---- image-mode kernel with TMA staging --------------------------------------------------------------------------------
// Same algorithm and per-warp ring as alpha_lattice_image_kernel, but a tile's strip -- 32 blocks = 128 pixels x 4 rows --
// is ONE cp.async.bulk.tensor.2d issued by lane 0 (box {128 px, 4 rows} of a 2-D tensor map over the RGBA image, landing as
// [row][lane] x 16 B: the layout the cp.async ring has), completion through one mbarrier per (warp, stage).  The other 31
// lanes issue nothing for the copy and no lane computes global addresses.  Needs blocks-per-row % 32 == 0 (a tile never
// straddles two block rows; w % 128 == 0: every power-of-two texture from 128 up); other widths take the cp.async kernel.
// Rows past h (h % 4 != 0) are never staged: those tiles go to the literal path through the queue, as in the cp.async kernel.
__device__ __forceinline__ void mbar_init(const uint32_t bar, const uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(const uint32_t bar, const uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(const uint32_t bar, const uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TXP_MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TXP_MBAR_DONE_%=;\n"
        "bra TXP_MBAR_WAIT_%=;\n"
        "TXP_MBAR_DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const uint32_t smem_dst, const void* tmap, const uint32_t x, const uint32_t y, const uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_dst), "l"(tmap), "r"(x), "r"(y), "r"(bar) : "memory");
}

// exactly one lane of a converged warp (so that ptxas issues the uniform-datapath UTMALDG once, without a loop over active lanes)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

struct alignas(64) TmaDesc { unsigned char bytes[128]; };       // CUtensorMap (opaque, 64-byte aligned) without <cuda.h> in device code

template <int THREADS, int STAGES>
constexpr size_t lattice_tma_smem() { return lattice_image_smem<THREADS, STAGES>() + (size_t)(THREADS / 32) * STAGES * 8; }

template <int FMT, int THREADS, int MIN_CTAS, int STAGES>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) alpha_lattice_tma_kernel(const __grid_constant__ TmaDesc tmap, const __grid_constant__ BlockSource src,
                                                                             uint8_t* __restrict__ out, const uint32_t ntiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint4* tab = reinterpret_cast<uint4*>(smem_raw);
    uint4* ring = tab + 512;                                                          // [warp][stage][row][lane]
    WarpQueue* queues = reinterpret_cast<WarpQueue*>(ring + THREADS * STAGES * 4);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(queues + THREADS / 32);   // [warp][stage]
    for (int i = threadIdx.x; i < 512; i += THREADS) tab[i] = g_alpha_lattice[i];
    if (threadIdx.x < (THREADS / 32) * STAGES) mbar_init((uint32_t)__cvta_generic_to_shared(bars + threadIdx.x), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (threadIdx.x == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    __syncthreads();
    // the warp index through a shuffle: ptxas then knows that everything derived from it (ring / barrier addresses, tile
    // coordinates) is warp-uniform and keeps it in uniform registers, which is where UTMALDG takes its operands from
    const uint32_t warp = __shfl_sync(FULL, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    WarpQueue& q = queues[warp];
    uint32_t qa = 0, qb[3] = {0, 0, 0};
    const uint32_t stride = gridDim.x * (THREADS / 32);
    const uint32_t tiles_per_row = src.bw >> 5, nblocks = (uint32_t)src.nblocks, full_rows = src.h >> 2;   // bw % 32 == 0 (host)
    const uint4* my = ring + (warp * STAGES * 4) * 32 + lane;                         // + (stage * 4 + row) * 32
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring + (warp * STAGES * 4) * 32);
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bars + warp * STAGES);

    // prefetch stream (warp-uniform): tile index -> (block row, tile in row) STAGES-1 tiles ahead, advanced incrementally
    const uint32_t tile0 = blockIdx.x * (THREADS / 32) + warp;
    uint32_t pt = tile0, pty = pt / tiles_per_row, ptx = pt - pty * tiles_per_row, pst = 0;
    const uint32_t step_q = stride / tiles_per_row, step_r = stride - step_q * tiles_per_row;
    uint32_t staged = 0;                                                              // bit k: the tile issued k prefetches ago was staged
    auto prefetch = [&]() {
        const bool ok = pt < ntiles && pty < full_rows;                               // warp-uniform
        if (ok) {
            if (elect_one()) {
                const uint32_t bar = bar_s + pst * 8;
                mbar_expect_tx(bar, 4 * 32 * 16);
                tma_load_2d(ring_s + pst * (4 * 32 * 16), &tmap, ptx * 128, pty * 4, bar);
            }
        }
        staged = (staged << 1) | (ok ? 1u : 0u);
        pt += stride; ptx += step_r; pty += step_q;
        if (ptx >= tiles_per_row) { ptx -= tiles_per_row; ++pty; }
        pst = pst + 1 == STAGES ? 0 : pst + 1;
    };
#pragma unroll
    for (int i = 0; i < STAGES - 1; ++i) prefetch();

    // One parity bit for all stages: it flips when the stage index wraps.  A warp's tiles are visited in increasing order and a
    // tile that is not staged (partial bottom row, past the image) is followed only by tiles that are not staged either, so no
    // barrier is ever waited on after a round in which it was skipped.
    uint32_t st = 0, phase = 0;
#pragma unroll 1
    for (uint32_t tile = tile0; tile < ntiles; tile += stride) {
        prefetch();                                       // into the stage every lane finished reading last iteration (__syncwarp in queue_push_drain)
        const uint32_t b = tile * 32 + lane;
        bool todo0 = false, todo1 = false;
        uint32_t reload = 0, VR[4] = {0, 0, 0, 0}, VG[4] = {0, 0, 0, 0};
        if ((staged >> (STAGES - 1)) & 1u) {              // warp-uniform
            mbar_wait(bar_s + st * 8, phase);
            if (b < nblocks) {
                uint32_t px[16];
                const uint4* sp = my + st * (4 * 32);
#pragma unroll
                for (int r = 0; r < 4; ++r) { const uint4 v = sp[r * 32]; px[4 * r] = v.x; px[4 * r + 1] = v.y; px[4 * r + 2] = v.z; px[4 * r + 3] = v.w; }
                alpha_block_body<FMT>(px, tab, out, b, todo0, VR, todo1, VG);
            }
        } else if (b < nblocks) {                         // partial bottom row
            todo0 = true; todo1 = FMT == BC5; reload = ITEM_RELOAD;
        }
        queue_push_drain<FMT>(q, qa, qb, lane, src, out, tab, b | reload, todo0, VR, todo1, VG);
        if (++st == STAGES) { st = 0; phase ^= 1u; }
    }
    // copies issued for tiles past the end were never started (ok == false), so nothing is in flight here
    queue_flush<FMT>(q, qa, qb, lane, src, out, tab);
}

}  // namespace txp
