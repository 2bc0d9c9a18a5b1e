// The reference's own unit tests (lib/src/lib.rs:345-505) restated against the C++ mirror of its API
// (include/texpresso.hpp).  Vectors: lib/src/test_data.rs.  Built and run by tests/test_cpp_mirror.py on the GPU box.
#include <cstdio>
#include <cstring>
#include <vector>
#include "texpresso.hpp"

using namespace texpresso;

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)

struct TestDataSet { std::vector<uint8_t> encoded, decoded; };

static std::vector<uint8_t> expand_single_to_rgb(const uint8_t (&in)[16]) {           // test_data.rs:137-148
    std::vector<uint8_t> out(48);
    for (int i = 0; i < 16; ++i) out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = in[i];
    return out;
}
static std::vector<uint8_t> add_alpha_to_rgb(const std::vector<uint8_t>& rgb, const uint8_t (&alpha)[16]) {   // :154-166
    std::vector<uint8_t> out(64);
    for (int i = 0; i < 16; ++i) { out[4 * i] = rgb[3 * i]; out[4 * i + 1] = rgb[3 * i + 1]; out[4 * i + 2] = rgb[3 * i + 2]; out[4 * i + 3] = alpha[i]; }
    return out;
}

static const uint8_t FF[16] = {255,255,255,255,255,255,255,255,255,255,255,255,255,255,255,255};
static const uint8_t LINEAR_RAMP[16] = {0x00,0x11,0x22,0x33,0x44,0x55,0x66,0x77,0x88,0x99,0xAA,0xBB,0xCC,0xDD,0xEE,0xFF};
static const uint8_t BC3_ALPHA_DECODED[16] = {0x00,0x24,0x48,0x6D,0x91,0xB6,0xDB,0xFF,0x00,0x24,0x48,0x6D,0x91,0xB6,0xDB,0xFF};
static const uint8_t GRAY_7F[16] = {0xFF,0x00,0xFF,0x00,0x00,0x7F,0x7F,0xFF,0xFF,0x7F,0x7F,0x00,0x00,0xFF,0x00,0xFF};
static const uint8_t GRAY_BLOCK_LUMA[16] = {0xFF,0x00,0xFF,0x00,0x00,0x55,0x55,0xFF,0xFF,0x55,0x55,0x00,0x00,0xFF,0x00,0xFF};
static std::vector<uint8_t> colour_block_rgb() {
    std::vector<uint8_t> v;
    const uint8_t rows[4][3] = {{0xFF,0x96,0x4A},{0xFF,0x78,0x34},{0xFF,0x69,0x29},{0xFF,0x69,0x29}};
    for (auto& r : rows) for (int i = 0; i < 4; ++i) v.insert(v.end(), r, r + 3);
    return v;
}

static void execute_decompression_test(Format format, const TestDataSet& data) {       // lib.rs:363-367
    std::vector<uint8_t> out(64);
    format.decompress(data.encoded, 4, 4, out);
    CHECK(out == data.decoded);
}

static void execute_compression_test(Format format, const TestDataSet& data) {         // lib.rs:369-393
    for (Algorithm algorithm : {Algorithm::ClusterFit, Algorithm::RangeFit, Algorithm::IterativeClusterFit}) {
        std::vector<uint8_t> out(format.block_size());
        format.compress(data.decoded, 4, 4, Params{algorithm, COLOUR_WEIGHTS_UNIFORM, false}, out);
        CHECK(out == data.encoded);
    }
}

int main() {
    // test_storage_requirements, lib.rs:350-361
    CHECK(Format(Format::Bc1).compressed_size(16, 32) == 256); CHECK(Format(Format::Bc1).compressed_size(15, 32) == 256);
    CHECK(Format(Format::Bc2).compressed_size(16, 32) == 512); CHECK(Format(Format::Bc2).compressed_size(15, 32) == 512);
    CHECK(Format(Format::Bc3).compressed_size(16, 32) == 512); CHECK(Format(Format::Bc3).compressed_size(15, 32) == 512);
    CHECK(Format(Format::Bc4).compressed_size(16, 32) == 256); CHECK(Format(Format::Bc4).compressed_size(15, 32) == 256);
    CHECK(Format(Format::Bc5).compressed_size(16, 32) == 512); CHECK(Format(Format::Bc5).compressed_size(15, 32) == 512);

    const TestDataSet BC1_GRAY{{0x00,0x00,0xFF,0xFF,0x11,0x68,0x29,0x44}, add_alpha_to_rgb(expand_single_to_rgb(GRAY_7F), FF)};
    const TestDataSet BC1_COLOUR{{0xA9,0xFC,0x45,0xFB,0x00,0xFF,0x55,0x55}, add_alpha_to_rgb(colour_block_rgb(), FF)};
    const TestDataSet BC2_GRAY{{0x10,0x32,0x54,0x76,0x98,0xBA,0xDC,0xFE,0xFF,0xFF,0x00,0x00,0x44,0x3D,0x7C,0x11},
                               add_alpha_to_rgb(expand_single_to_rgb(GRAY_BLOCK_LUMA), LINEAR_RAMP)};
    const TestDataSet BC2_COLOUR{{0x10,0x32,0x54,0x76,0x98,0xBA,0xDC,0xFE,0xA9,0xFC,0x45,0xFB,0x00,0xFF,0x55,0x55},
                                 add_alpha_to_rgb(colour_block_rgb(), LINEAR_RAMP)};
    const TestDataSet BC3_GRAY{{0x24,0xDB,0x86,0xC6,0xE6,0x86,0xC6,0xE6,0xFF,0xFF,0x00,0x00,0x44,0x3D,0x7C,0x11},
                               add_alpha_to_rgb(expand_single_to_rgb(GRAY_BLOCK_LUMA), BC3_ALPHA_DECODED)};
    const TestDataSet BC3_COLOUR{{0x24,0xDB,0x86,0xC6,0xE6,0x86,0xC6,0xE6,0xA9,0xFC,0x45,0xFB,0x00,0xFF,0x55,0x55},
                                 add_alpha_to_rgb(colour_block_rgb(), BC3_ALPHA_DECODED)};
    const TestDataSet BC4_GRAY{{0x7F,0x84,0xF7,0x6D,0xE0,0x07,0xEC,0xFB}, add_alpha_to_rgb(expand_single_to_rgb(GRAY_7F), FF)};
    std::vector<uint8_t> bc5_rgb = {
        0xFF,0x00,0x00, 0x00,0xFF,0x00, 0xFF,0x00,0x00, 0x00,0xFF,0x00,
        0x00,0xFF,0x00, 0x7F,0x7F,0x00, 0x7F,0x7F,0x00, 0xFF,0x00,0x00,
        0xFF,0x00,0x00, 0x7F,0x7F,0x00, 0x7F,0x7F,0x00, 0x00,0xFF,0x00,
        0x00,0xFF,0x00, 0xFF,0x00,0x00, 0x00,0xFF,0x00, 0xFF,0x00,0x00};
    const TestDataSet BC5_GRAY{{0x7F,0x84,0xF7,0x6D,0xE0,0x07,0xEC,0xFB,0x7F,0x84,0xBE,0x7F,0xC0,0x06,0x7E,0xDF}, add_alpha_to_rgb(bc5_rgb, FF)};

    execute_decompression_test(Format::Bc1, BC1_GRAY);   execute_compression_test(Format::Bc1, BC1_GRAY);       // lib.rs:395-403
    execute_decompression_test(Format::Bc1, BC1_COLOUR); execute_compression_test(Format::Bc1, BC1_COLOUR);     // :405-413
    execute_decompression_test(Format::Bc2, BC2_GRAY);   execute_compression_test(Format::Bc2, BC2_GRAY);       // :446-454
    execute_decompression_test(Format::Bc2, BC2_COLOUR); execute_compression_test(Format::Bc2, BC2_COLOUR);     // :456-464
    execute_decompression_test(Format::Bc3, BC3_GRAY);   execute_compression_test(Format::Bc3, BC3_GRAY);       // :466-474
    execute_decompression_test(Format::Bc3, BC3_COLOUR); execute_compression_test(Format::Bc3, BC3_COLOUR);     // :476-484
    execute_decompression_test(Format::Bc4, BC4_GRAY);   execute_compression_test(Format::Bc4, BC4_GRAY);       // :486-494
    execute_decompression_test(Format::Bc5, BC5_GRAY);   execute_compression_test(Format::Bc5, BC5_GRAY);       // :496-504

    {   // test_bc1_decompression_height_not_multiple_of_4, lib.rs:415-444
        const std::vector<uint8_t> encoded = {0x8E,0x73,0x71,0x8C,0xAA,0xAA,0xAA,0xAA,0x8E,0x73,0x71,0x8C,0xAA,0xAA,0xFF,0xFF};
        std::vector<uint8_t> output(4 * 4 * 6);
        Format(Format::Bc1).decompress(encoded, 4, 6, output);
        const uint8_t REFERENCE[4] = {0x7F, 0x7F, 0x7F, 0xFF};
        for (size_t px = 0; px < output.size() / 4; ++px) CHECK(std::memcmp(&output[4 * px], REFERENCE, 4) == 0);
    }
    {   // the panic contract (lib.rs:295): a too-short output buffer must be rejected
        std::vector<uint8_t> rgba(64), out(4);
        bool threw = false;
        try { Format(Format::Bc1).compress(rgba, 4, 4, Params{}, out); } catch (const Error& e) { threw = e.code == TXP_ERR_BUFFER_TOO_SMALL; }
        CHECK(threw);
    }
    {   // block entry points, lib.rs:188-234 / :240-277
        std::array<std::array<uint8_t, 4>, 16> px{};
        for (int i = 0; i < 16; ++i) px[i] = {BC1_COLOUR.decoded[4 * i], BC1_COLOUR.decoded[4 * i + 1], BC1_COLOUR.decoded[4 * i + 2], 255};
        uint8_t blk[8];
        Format(Format::Bc1).compress_block_masked(px, 0xFFFF, Params{Algorithm::ClusterFit, COLOUR_WEIGHTS_UNIFORM, false}, blk, 8);
        CHECK(std::memcmp(blk, BC1_COLOUR.encoded.data(), 8) == 0);
        const auto dec = Format(Format::Bc1).decompress_block(blk, 8);
        CHECK(std::memcmp(&dec[0][0], BC1_COLOUR.decoded.data(), 64) == 0);
    }
    std::printf(failures ? "%d FAILURES\n" : "all reference unit tests passed (%d failures)\n", failures);
    return failures ? 1 : 0;
}
