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

constexpr uint32_t MAGIC15 = 0x47400000u;         // 1.5 * 2^15 as fp32 bits
constexpr float MAGIC23 = 12582912.0f;            // 1.5 * 2^23

// PRMT with the hardware selector semantics (bit 3 of a nibble = sign replication).  __byte_perm() promises to ignore
// that bit, so nvcc masks every run-time selector with 0x7777 first; all selectors here have bit 3 clear by construction.
__device__ __forceinline__ uint32_t prmt(const uint32_t a, const uint32_t b, const uint32_t sel) {
    uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d;
}

// One book over 16 pixels.  vm: the pixels as 1.5*2^15 + v; V: the same values as packed bytes (4 per word);
// origin / a: d = vm - origin, slot = rint(d * a) (a < 0 for the 7-point book: x = hi - v);
// clo/chi: the codes of slots 0..7 as bytes.  Returns sum (v - code)^2; sel[k] holds the slots of pixels 4k..4k+3 as nibbles.
__device__ __forceinline__ uint32_t lattice_book(const uint32_t vm[16], const uint32_t V[4], const float origin, const float a,
                                                 const uint32_t clo, const uint32_t chi, uint32_t sel[4]) {
    const f32x2 o2 = pk(origin, origin), a2 = pk(a, a), m2 = pk(MAGIC23, MAGIC23);
    uint32_t err = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const f32x2 d01 = sub2(pk(__uint_as_float(vm[4 * k]), __uint_as_float(vm[4 * k + 1])), o2);
        const f32x2 d23 = sub2(pk(__uint_as_float(vm[4 * k + 2]), __uint_as_float(vm[4 * k + 3])), o2);
        float t0, t1, t2, t3;
        upk(fma2(d01, a2, m2), t0, t1);
        upk(fma2(d23, a2, m2), t2, t3);
        // low 16 bits: four slot nibbles (the magic's low 22 bits are zero, so nothing else reaches them)
        const uint32_t s = ((__float_as_uint(t3) * 16u + __float_as_uint(t2)) * 16u + __float_as_uint(t1)) * 16u + __float_as_uint(t0);
        sel[k] = s;
        const uint32_t c4 = prmt(clo, chi, s);
        const uint32_t e4 = __vabsdiffu4(c4, V[k]);
        err = __dp4a(e4, e4, err);
    }
    return err;
}

// One channel (byte CH of every RGBA word) of one fully valid block.  Returns false if the block is not regular.
template <int CH>
__device__ __forceinline__ bool alpha_fit_lattice(const uint32_t px[16], const uint4* __restrict__ tab, uint2& out) {
    uint32_t vm[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) vm[i] = prmt(MAGIC15, px[i], CH == 0 ? 0x3240u : 0x3250u);   // (0x00, v, 0x40, 0x47)
    // min / max on the raw bits (positive floats order like integers)
    uint32_t mn = __vimin3_u32(vm[0], vm[1], vm[2]), mx = __vimax3_u32(vm[0], vm[1], vm[2]);
#pragma unroll
    for (int i = 3; i < 15; i += 2) { mn = __vimin3_u32(mn, vm[i], vm[i + 1]); mx = __vimax3_u32(mx, vm[i], vm[i + 1]); }
    mn = min(mn, vm[15]); mx = max(mx, vm[15]);
    const uint32_t span = mx - mn;                                     // r << 8
    if (!(mn > MAGIC15 && mx < (MAGIC15 | 0xFF00u) && span >= (7u << 8))) return false;

    const uint4 e5 = tab[span >> 7], e7 = tab[(span >> 7) + 1];        // row r = two uint4
    const uint32_t lo = (mn >> 8) & 255u, hi = (mx >> 8) & 255u;
    uint32_t V[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t s2 = CH == 0 ? 0x0040u : 0x0051u;               // byte CH of both words
        V[k] = __byte_perm(__byte_perm(px[4 * k], px[4 * k + 1], s2), __byte_perm(px[4 * k + 2], px[4 * k + 3], s2), 0x5410u);
    }
    uint32_t s5[4], s7[4];
    // 5-point book from lo: x = v - lo, origin = lo - beta5, codes lo + offs
    const uint32_t err5 = lattice_book(vm, V, __fsub_rn(__uint_as_float(mn), __uint_as_float(e5.y)), __uint_as_float(e5.x),
                                       e5.z + lo * 0x01010101u, e5.w + lo * 0x01010101u, s5);
    // 7-point book from hi: x = hi - v = -(v - (hi + beta7)), codes hi - offs
    const uint32_t err7 = lattice_book(vm, V, __fadd_rn(__uint_as_float(mx), __uint_as_float(e7.y)), -__uint_as_float(e7.x),
                                       hi * 0x01010101u - e7.z, hi * 0x01010101u - e7.w, s7);
    const bool five = err5 <= err7;                                    // alpha.rs:251
    const uint32_t a0 = five ? lo : hi, a1 = five ? hi : lo;          // write_alpha_block5 as is / write_alpha_block7 swapped
    const uint32_t mhi = five ? 0x00000105u : 0x01070605u;            // slot -> index: 0 -> 0, N -> 1, s -> s + 1
    uint32_t z[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t w = prmt(0x04030200u, mhi, five ? s5[k] : s7[k]);   // four 3-bit indices, one per byte
        const uint32_t y = (w | (w >> 5)) & 0x003F003Fu;
        z[k] = (y | (y >> 10)) & 0xFFFu;
    }
    const uint32_t g0 = z[0] | (z[1] << 12), g1 = z[2] | (z[3] << 12);
    out = make_uint2(a0 | (a1 << 8) | (g0 << 16), (g0 >> 16) | (g1 << 8));
    return true;
}

// literal path for one queued (block, channel) item
template <int FMT>
__device__ __noinline__ void alpha_drain_item(const BlockSource& src, uint8_t* __restrict__ out, const uint32_t item) {
    const uint32_t b = item & 0x7FFFFFFFu, ch = item >> 31;
    uint32_t px[16], mask;
    load_block_thread(src, b, px, mask);
    uint32_t v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (px[i] >> (8 * ch)) & 255u;    // lib.rs:200, :202-203
    const uint2 r = mask == 0xFFFFu ? alpha_fit_full(v) : alpha_fit_thread(v, mask);
    reinterpret_cast<uint2*>(out)[FMT == BC4 ? (size_t)b : 2 * (size_t)b + ch] = r;
}

constexpr int LATTICE_QUEUE = 96;                 // <= 31 left over + 2 x 32 new items per tile

template <int FMT, int THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) alpha_lattice_kernel(const __grid_constant__ BlockSource src, uint8_t* __restrict__ out, const uint32_t ntiles) {
    __shared__ uint4 tab[512];
    __shared__ uint32_t queue[THREADS / 32][LATTICE_QUEUE];
    for (int i = threadIdx.x; i < 512; i += THREADS) tab[i] = g_alpha_lattice[i];
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t* q = queue[warp];
    uint32_t qn = 0;
    const uint32_t stride = gridDim.x * (THREADS / 32);
    for (uint32_t tile = blockIdx.x * (THREADS / 32) + warp; tile < ntiles; tile += stride) {
        const uint64_t b = (uint64_t)tile * 32 + lane;
        bool todo0 = false, todo1 = false;
        if (b < src.nblocks) {
            uint32_t px[16], mask;
            load_block_thread(src, b, px, mask);
            if (mask == 0xFFFFu) {
                uint2 r0, r1;
                const bool ok0 = alpha_fit_lattice<0>(px, tab, r0);
                todo0 = !ok0;
                if (FMT == BC4) {
                    if (ok0) reinterpret_cast<uint2*>(out)[b] = r0;
                } else {
                    const bool ok1 = alpha_fit_lattice<1>(px, tab, r1);
                    todo1 = !ok1;
                    if (ok0 && ok1) reinterpret_cast<uint4*>(out)[b] = make_uint4(r0.x, r0.y, r1.x, r1.y);
                    else if (ok0) reinterpret_cast<uint2*>(out)[2 * b] = r0;
                    else if (ok1) reinterpret_cast<uint2*>(out)[2 * b + 1] = r1;
                }
            } else {
                todo0 = true; todo1 = FMT == BC5;
            }
        }
        const uint32_t m0 = __ballot_sync(FULL, todo0);
        if (todo0) q[qn + __popc(m0 & lt)] = (uint32_t)b;
        qn += __popc(m0);
        if (FMT == BC5) {
            const uint32_t m1 = __ballot_sync(FULL, todo1);
            if (todo1) q[qn + __popc(m1 & lt)] = (uint32_t)b | 0x80000000u;
            qn += __popc(m1);
        }
        __syncwarp();
#pragma unroll 1
        while (qn >= 32) {
            qn -= 32;
            const uint32_t item = q[qn + lane];
            __syncwarp();
            alpha_drain_item<FMT>(src, out, item);
        }
    }
    if (lane < qn) alpha_drain_item<FMT>(src, out, q[lane]);
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
constexpr size_t lattice_image_smem() { return 512 * sizeof(uint4) + (size_t)THREADS * STAGES * 64 + (THREADS / 32) * LATTICE_QUEUE * sizeof(uint32_t); }

template <int FMT, int THREADS, int MIN_CTAS, int STAGES>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) alpha_lattice_image_kernel(const __grid_constant__ BlockSource src, uint8_t* __restrict__ out,
                                                                               const uint32_t ntiles, const uint32_t step_q, const uint32_t step_r) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint4* tab = reinterpret_cast<uint4*>(smem_raw);
    uint4* ring = tab + 512;                                                          // [warp][stage][row][lane]
    uint32_t* queue = reinterpret_cast<uint32_t*>(ring + THREADS * STAGES * 4);
    for (int i = threadIdx.x; i < 512; i += THREADS) tab[i] = g_alpha_lattice[i];
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t* q = queue + warp * LATTICE_QUEUE;
    uint32_t qn = 0;
    const uint32_t stride = gridDim.x * (THREADS / 32);
    const uint32_t bw = src.bw, nblocks = (uint32_t)src.nblocks, full_rows = src.h >> 2;   // block rows with 4 pixel rows
    const size_t pitch = (size_t)src.w * 4;
    uint4* my = ring + (warp * STAGES * 4) * 32 + lane;                               // + (stage * 4 + row) * 32
    const uint32_t my_s = (uint32_t)__cvta_generic_to_shared(my);

    const uint32_t tile0 = blockIdx.x * (THREADS / 32) + warp;
    // prefetch stream
    uint32_t pb = tile0 * 32 + lane, pby = pb / bw, pbx = pb - pby * bw, pst = 0;
    auto prefetch = [&]() {
        if (pb < nblocks && pby < full_rows) {
            const uint8_t* g = src.rgba + (size_t)pby * 4 * pitch + (size_t)pbx * 16;
            const uint32_t s = my_s + pst * (4 * 32 * 16);
#pragma unroll
            for (int r = 0; r < 4; ++r) cp_async16(s + r * (32 * 16), g + r * pitch);
        }
        cp_async_commit();
        pb += stride * 32; pbx += step_r; pby += step_q;
        if (pbx >= bw) { pbx -= bw; ++pby; }
        pst = pst + 1 == STAGES ? 0 : pst + 1;
    };
#pragma unroll
    for (int i = 0; i < STAGES - 1; ++i) prefetch();

    uint32_t b = tile0 * 32 + lane, by = b / bw, bx = b - by * bw, st = 0;
#pragma unroll 1
    for (uint32_t tile = tile0; tile < ntiles; tile += stride) {
        prefetch();
        cp_async_wait<STAGES - 1>();
        bool todo0 = false, todo1 = false;
        if (b < nblocks) {
            if (by < full_rows) {
                uint32_t px[16];
                const uint4* sp = my + st * (4 * 32);
#pragma unroll
                for (int r = 0; r < 4; ++r) { const uint4 v = sp[r * 32]; px[4 * r] = v.x; px[4 * r + 1] = v.y; px[4 * r + 2] = v.z; px[4 * r + 3] = v.w; }
                uint2 r0, r1;
                const bool ok0 = alpha_fit_lattice<0>(px, tab, r0);
                todo0 = !ok0;
                if (FMT == BC4) {
                    if (ok0) reinterpret_cast<uint2*>(out)[b] = r0;
                } else {
                    const bool ok1 = alpha_fit_lattice<1>(px, tab, r1);
                    todo1 = !ok1;
                    if (ok0 && ok1) reinterpret_cast<uint4*>(out)[b] = make_uint4(r0.x, r0.y, r1.x, r1.y);
                    else if (ok0) reinterpret_cast<uint2*>(out)[2 * (size_t)b] = r0;
                    else if (ok1) reinterpret_cast<uint2*>(out)[2 * (size_t)b + 1] = r1;
                }
            } else {
                todo0 = true; todo1 = FMT == BC5;
            }
        }
        const uint32_t m0 = __ballot_sync(FULL, todo0);
        if (todo0) q[qn + __popc(m0 & lt)] = b;
        qn += __popc(m0);
        if (FMT == BC5) {
            const uint32_t m1 = __ballot_sync(FULL, todo1);
            if (todo1) q[qn + __popc(m1 & lt)] = b | 0x80000000u;
            qn += __popc(m1);
        }
        __syncwarp();
#pragma unroll 1
        while (qn >= 32) {
            qn -= 32;
            const uint32_t item = q[qn + lane];
            __syncwarp();
            alpha_drain_item<FMT>(src, out, item);
        }
        b += stride * 32; bx += step_r; by += step_q;
        if (bx >= bw) { bx -= bw; ++by; }
        st = st + 1 == STAGES ? 0 : st + 1;
    }
    cp_async_wait<0>();
    if (lane < qn) alpha_drain_item<FMT>(src, out, q[lane]);
}

}  // namespace txp
