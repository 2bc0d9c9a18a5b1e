/*
 * texpresso_b200.h -- C ABI of the B200 (sm_100a) BC1..BC5 encoder / decoder.
 *
 * Drop-in boundary for the hot path of jansol/texpresso: the reference has no FFI of its own, its
 * boundary is the Rust public API in lib/src/lib.rs:38-336.  Every entry point below names the
 * reference item it replaces; INTEGRATION.md shows the `extern "C"` block a `texpresso-cuda` crate binds.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns and pre-sizes every buffer, nothing is retained
 *     after a call returns (reference: #![no_std] without alloc, lib.rs:25).
 *   - the reference's only error path is a panic (assert / slice bounds).  Here every function returns
 *     TXP_OK or a negative TXP_ERR_* code; a binding that wants the reference's contract asserts on != 0.
 *   - calls are synchronous and thread-safe.  Host-pointer entry points use the calling thread's current
 *     CUDA device (txp_set_device) and an internal per-device context (streams, pinned staging, device
 *     scratch).  There is NO CPU fallback: without a usable CUDA device they return TXP_ERR_CUDA.
 *   - pixel data is RGBA8, row-major, tightly packed (lib.rs:323); blocks are row-major, 8 bytes
 *     (BC1, BC4) or 16 bytes (BC2, BC3, BC5) each (lib.rs:159-168).
 */
#ifndef TEXPRESSO_B200_H
#define TEXPRESSO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define TXP_API __declspec(dllexport)
#else
#define TXP_API __attribute__((visibility("default")))
#endif

/* reference `enum Format`, lib.rs:39-46 (declaration order) */
enum { TXP_FORMAT_BC1 = 0, TXP_FORMAT_BC2 = 1, TXP_FORMAT_BC3 = 2, TXP_FORMAT_BC4 = 3, TXP_FORMAT_BC5 = 4 };

/* reference `enum Algorithm`, lib.rs:49-59; Default = ClusterFit (lib.rs:61-65) */
enum { TXP_ALGORITHM_RANGE_FIT = 0, TXP_ALGORITHM_CLUSTER_FIT = 1, TXP_ALGORITHM_ITERATIVE_CLUSTER_FIT = 2 };

enum {
    TXP_OK = 0,
    TXP_ERR_FORMAT = -1,           /* format / algorithm value out of range */
    TXP_ERR_DIMENSIONS = -2,       /* width == 0 (reference: chunks_mut(0) panics) or image too large */
    TXP_ERR_BUFFER_TOO_SMALL = -3, /* reference: assert!(output.len() >= compressed_size) lib.rs:295, slice panics :138/:324 */
    TXP_ERR_CUDA = -4,             /* no device / CUDA runtime failure; see txp_last_error() */
    TXP_ERR_ARGUMENT = -5          /* null pointer, trailing partial block, bad gpu count ... */
};

/* reference `struct Params`, lib.rs:76-90.  Default (lib.rs:92-100): ClusterFit, PERCEPTUAL, false. */
typedef struct txp_params {
    uint32_t algorithm;             /* TXP_ALGORITHM_* */
    float weights[3];               /* ColourWeights, lib.rs:68; UNIFORM {1,1,1} :71; PERCEPTUAL {0.2126,0.7152,0.0722} :74 */
    uint32_t weigh_colour_by_alpha; /* bool */
} txp_params;

/* ---- sizes -------------------------------------------------------------------------------------- */
TXP_API size_t txp_num_blocks(size_t size);                                   /* num_blocks, lib.rs:103-105 */
TXP_API size_t txp_block_size(int format);                                    /* Format::block_size, lib.rs:159-168; 0 if bad format */
TXP_API size_t txp_compressed_size(int format, size_t width, size_t height);  /* Format::compressed_size, lib.rs:175-179 */

/* ---- host-pointer entry points (H2D -> kernels -> D2H inside the call) ----------------------------- */

/* Format::compress, lib.rs:287-335.  Encodes output_len / block_size blocks in block-row-major order;
 * blocks lying past ceil(height/4) rows are encoded fully masked, as the reference does when `output`
 * is longer than compressed_size (its loop runs over output.chunks_mut).  rgba_len >= 4*width*height. */
TXP_API int txp_compress(int format, const uint8_t* rgba, size_t rgba_len, size_t width, size_t height,
                         const txp_params* params, uint8_t* output, size_t output_len);

/* Extension: Format::compress on an image that is still in its decoded file layout.  The reference's
 * CLI expands such images to RGBA8 on the host before compress (cli/src/image/png.rs:47-62, jpeg.rs:42-52):
 * L8 -> (l, l, l, 255), LA8 -> (l, l, l, a), RGB8 -> (r, g, b, 255).  Here the 1-3 byte pixels are copied to the device
 * as they are and expanded there (4x / 2x / 1.33x less host-to-device traffic: BC4 / BC5 are PCIe-bound through the host API).
 * The result is byte-identical to expanding on the host and calling txp_compress.  Host pointers only. */
#define TXP_PIXELS_L8 1
#define TXP_PIXELS_LA8 2
#define TXP_PIXELS_RGB8 3
#define TXP_PIXELS_RGBA8 4
#define TXP_PIXELS_RG8 5    /* no counterpart in the reference: (r, g) -> (r, g, 0, 255), two-channel normal maps for BC5 */
TXP_API int txp_compress_pixels(int format, const uint8_t* pixels, size_t pixels_len, int layout, size_t width, size_t height,
                                const txp_params* params, uint8_t* output, size_t output_len);

/* Format::decompress, lib.rs:124-156.  data_len >= compressed_size, output_len >= 4*width*height. */
TXP_API int txp_decompress(int format, const uint8_t* data, size_t data_len, size_t width, size_t height,
                           uint8_t* output, size_t output_len);

/* Format::compress_block_masked, lib.rs:188-234.  mask bit i = pixel i (i = 4*py+px) is valid. */
TXP_API int txp_compress_block_masked(int format, const uint8_t rgba[64], uint32_t mask,
                                      const txp_params* params, uint8_t* output, size_t output_len);

/* Format::decompress_block, lib.rs:240-277. */
TXP_API int txp_decompress_block(int format, const uint8_t* block, size_t block_len, uint8_t output[64]);

/* n independent compress_block_masked / decompress_block calls in one launch
 * (rgba_blocks: n x 64 bytes, masks: n words, output: n x block_size bytes). */
TXP_API int txp_compress_blocks(int format, const uint8_t* rgba_blocks, const uint32_t* masks, size_t n,
                                const txp_params* params, uint8_t* output);
TXP_API int txp_decompress_blocks(int format, const uint8_t* blocks, size_t n, uint8_t* rgba_blocks);

/* ---- device-pointer entry points (data resident in HBM; asynchronous on `cuda_stream`) ---------------- */
/* Same semantics as txp_compress / txp_decompress with device pointers on the current device.
 * cuda_stream is a cudaStream_t (NULL = default stream).  No synchronisation is performed. */
TXP_API int txp_compress_device(int format, const void* d_rgba, size_t width, size_t height,
                                const txp_params* params, void* d_output, size_t output_len, void* cuda_stream);
TXP_API int txp_decompress_device(int format, const void* d_data, size_t width, size_t height,
                                  void* d_output, size_t output_len, void* cuda_stream);

/* ---- sharding (reference: rayon par_chunks_mut over block rows, lib.rs:300-305 / :128-134) -------------- */
/* Balanced block-row range [*row_begin, *row_end) of shard `rank` out of `world` for an image of `height`. */
TXP_API void txp_shard_rows(size_t height, int rank, int world, size_t* row_begin, size_t* row_end);

/* One process, n_gpus devices (0..n_gpus-1; n_gpus == 1 means the calling thread's current device, so that one process per GPU
 * can use the multi / batch entry points as they are): block rows are split with txp_shard_rows, one host worker
 * per device, each device writes its own slice of `output`.  No collectives. */
TXP_API int txp_compress_multi(int format, const uint8_t* rgba, size_t rgba_len, size_t width, size_t height,
                               const txp_params* params, uint8_t* output, size_t output_len, int n_gpus);

/* Batch of independent textures (e.g. mip levels), texture t -> device t % n_gpus. */
TXP_API int txp_compress_batch(int format, const uint8_t* const* rgba, const size_t* widths, const size_t* heights,
                               size_t n_textures, const txp_params* params, uint8_t* const* outputs, int n_gpus);

/* Format::decompress (lib.rs:124-156) sharded like the reference's rayon loop over block rows (lib.rs:128-134): block rows
 * are split with txp_shard_rows over devices 0..n_gpus-1, each device decodes its slice of `data` into its own pixel rows
 * of `output`.  No collectives. */
TXP_API int txp_decompress_multi(int format, const uint8_t* data, size_t data_len, size_t width, size_t height,
                                 uint8_t* output, size_t output_len, int n_gpus);

/* Batch of independent compressed textures, texture t -> device t % n_gpus; data[t] holds compressed_size(format, widths[t],
 * heights[t]) bytes, outputs[t] receives 4*widths[t]*heights[t] bytes. */
TXP_API int txp_decompress_batch(int format, const uint8_t* const* data, const size_t* widths, const size_t* heights,
                                 size_t n_textures, uint8_t* const* outputs, int n_gpus);

/* ---- mip chains (extension: the reference generates no mips, cli/src/main.rs:153) ------------------------------------ */
/* Levels: (w,h), (max(1,w/2), max(1,h/2)), ... down to 1x1; each level is the 2x2 box filter (a+b+c+d+2)>>2 of the
 * previous one (edge-clamped), generated on the device.  Output: the levels' blocks concatenated, level 0 first. */
TXP_API int txp_mip_levels(size_t width, size_t height);
TXP_API size_t txp_mipchain_compressed_size(int format, size_t width, size_t height);
TXP_API int txp_compress_mipchain(int format, const uint8_t* rgba, size_t rgba_len, size_t width, size_t height,
                                  const txp_params* params, uint8_t* output, size_t output_len);
/* Batch of textures, each encoded with its full mip chain; texture t -> device t % n_gpus, copies overlapped with kernels. */
TXP_API int txp_compress_batch_mips(int format, const uint8_t* const* rgba, const size_t* widths, const size_t* heights,
                                    size_t n_textures, const txp_params* params, uint8_t* const* outputs, int n_gpus);

/* ---- runtime ------------------------------------------------------------------------------------------ */
TXP_API int txp_device_count(void);          /* number of CUDA devices, or TXP_ERR_CUDA */
TXP_API int txp_set_device(int device);      /* cudaSetDevice for the calling thread */
TXP_API const char* txp_last_error(void);    /* thread-local description of the last failure */
TXP_API uint64_t txp_kernel_launches(void);  /* kernels launched by this library since load */
TXP_API const char* txp_version(void);
TXP_API int txp_debug_get(int key, uint64_t* value);   /* key 0 / 1: the knobs above; 2 / 3 / 4: ClusterFit launches so far that took the lane-per-block search, the lane-per-block iterative search, the warp-per-block search (tests assert which structure an entry point reached) */
/* Measurement helper for bench.py: rate of independent rounded fp32 products (FMUL, one lane-operation each) on the current
 * device, in lane-operations per second -- the measured counterpart of SMs x 128 lanes x clock, the roof of the ClusterFit
 * search under the reference's no-FMA-contraction contract. */
TXP_API int txp_measure_fp32_issue(double* lane_ops_per_second);
TXP_API int txp_debug_plan(int format, const txp_params* params, size_t width, size_t rows, int sm_count, size_t* chunk_rows, size_t max_chunks, size_t* n_chunks, uint64_t* lane_chunks);   /* the host pipeline's chunk plan for `rows` block rows of a `width`-pixel-wide image on a device with sm_count SMs: block rows per chunk in order, and a bit mask of the chunks that take the lane-per-block search below the lone-launch threshold; host-only, testable without a GPU */
TXP_API int txp_debug_host_copy(void* dst, const void* src, size_t n);   /* the multi-threaded staging copy the library uses for pageable caller buffers (CopyPool, TXP_COPY_THREADS), exposed so that it can be tested without a GPU */
TXP_API int txp_debug_set(int key, int value);   /* tuning knobs for A/B measurements; key 0: ClusterFit kernel structure 0 auto, 1 fused, 2 warp per block, 3 lane per block; key 1: smallest launch (blocks) that takes the lane-per-block search in auto mode; key 0 value 4: full lane rounds + warp-per-block tail; key 2: tail threshold of that split in percent of a round (0 = off, default); key 3: host pipeline chunk size override in MiB (0 = automatic); key 4: growth factor of the geometric chunk plan of small ClusterFit shards; key 5: smallest shard, in rounds of the lane-per-block search, that takes the round-aligned chunk plan (0 = off, default 4); key 6: the largest such shard in rounds (default 40); key 7: rounds per lane chunk of that plan (default 2) */

#ifdef __cplusplus
}
#endif
#endif /* TEXPRESSO_B200_H */
