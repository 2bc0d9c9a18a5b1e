// texpresso.hpp -- C++ host-side mirror of the reference's public API (jansol/texpresso, lib/src/lib.rs:38-336)
// over the C ABI of libtexpresso_b200 (include/texpresso_b200.h).  Header only.
//
//   texpresso::Format::Bc1 .. Bc5, texpresso::Algorithm, texpresso::Params, COLOUR_WEIGHTS_UNIFORM / _PERCEPTUAL,
//   num_blocks, block_size, compressed_size, compress, decompress, compress_block_masked, decompress_block
//
// Same names, argument order and meaning as the Rust items.  Where the reference panics (assert! lib.rs:295,
// slice bounds :138/:324) these throw texpresso::Error.  All arithmetic runs in the CUDA kernels; there is no
// CPU fallback.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "texpresso_b200.h"

namespace texpresso {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("texpresso_b200 error " + std::to_string(c) + ": " + m), code(c) {}
};

inline void check(int rc) { if (rc != TXP_OK) throw Error(rc, txp_last_error()); }

/// lib.rs:49-65
enum class Algorithm : uint32_t { RangeFit = 0, ClusterFit = 1, IterativeClusterFit = 2 };

/// lib.rs:68-74
using ColourWeights = std::array<float, 3>;
constexpr ColourWeights COLOUR_WEIGHTS_UNIFORM{1.0f, 1.0f, 1.0f};
constexpr ColourWeights COLOUR_WEIGHTS_PERCEPTUAL{0.2126f, 0.7152f, 0.0722f};

/// lib.rs:76-100 (Default: ClusterFit, perceptual weights, no alpha weighting)
struct Params {
    Algorithm algorithm = Algorithm::ClusterFit;
    ColourWeights weights = COLOUR_WEIGHTS_PERCEPTUAL;
    bool weigh_colour_by_alpha = false;

    txp_params c() const {
        return txp_params{static_cast<uint32_t>(algorithm), {weights[0], weights[1], weights[2]}, weigh_colour_by_alpha ? 1u : 0u};
    }
};

/// lib.rs:103-105
inline std::size_t num_blocks(std::size_t size) { return txp_num_blocks(size); }

/// lib.rs:39-46 with the methods of `impl Format` (lib.rs:117-336)
class Format {
public:
    enum Value : int { Bc1 = 0, Bc2 = 1, Bc3 = 2, Bc4 = 3, Bc5 = 4 };
    constexpr Format(Value v) : v_(v) {}
    constexpr operator Value() const { return v_; }

    /// lib.rs:159-168
    std::size_t block_size() const { return txp_block_size(v_); }
    /// lib.rs:175-179
    std::size_t compressed_size(std::size_t width, std::size_t height) const { return txp_compressed_size(v_, width, height); }

    /// lib.rs:287-335
    void compress(const uint8_t* rgba, std::size_t rgba_len, std::size_t width, std::size_t height, const Params& params,
                  uint8_t* output, std::size_t output_len) const {
        const txp_params p = params.c();
        check(txp_compress(v_, rgba, rgba_len, width, height, &p, output, output_len));
    }
    void compress(const std::vector<uint8_t>& rgba, std::size_t width, std::size_t height, const Params& params,
                  std::vector<uint8_t>& output) const {
        compress(rgba.data(), rgba.size(), width, height, params, output.data(), output.size());
    }

    /// Extension: compress() on an image in its decoded file layout (TXP_PIXELS_L8 / LA8 / RGB8 / RGBA8 / RG8); the expansion
    /// the reference's CLI does on the host (cli/src/image/png.rs:47-62) runs on the device.
    void compress_pixels(const uint8_t* pixels, std::size_t pixels_len, int layout, std::size_t width, std::size_t height,
                         const Params& params, uint8_t* output, std::size_t output_len) const {
        const txp_params p = params.c();
        check(txp_compress_pixels(v_, pixels, pixels_len, layout, width, height, &p, output, output_len));
    }

    /// lib.rs:124-156
    void decompress(const uint8_t* data, std::size_t data_len, std::size_t width, std::size_t height, uint8_t* output,
                    std::size_t output_len) const {
        check(txp_decompress(v_, data, data_len, width, height, output, output_len));
    }
    void decompress(const std::vector<uint8_t>& data, std::size_t width, std::size_t height, std::vector<uint8_t>& output) const {
        decompress(data.data(), data.size(), width, height, output.data(), output.size());
    }

    /// lib.rs:188-234
    void compress_block_masked(const std::array<std::array<uint8_t, 4>, 16>& rgba, uint32_t mask, const Params& params,
                               uint8_t* output, std::size_t output_len) const {
        const txp_params p = params.c();
        check(txp_compress_block_masked(v_, &rgba[0][0], mask, &p, output, output_len));
    }

    /// lib.rs:240-277
    std::array<std::array<uint8_t, 4>, 16> decompress_block(const uint8_t* block, std::size_t block_len) const {
        std::array<std::array<uint8_t, 4>, 16> out{};
        check(txp_decompress_block(v_, block, block_len, &out[0][0]));
        return out;
    }

private:
    Value v_;
};

}  // namespace texpresso
