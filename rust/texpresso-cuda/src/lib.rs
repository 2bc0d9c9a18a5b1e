// Source-only (not compiled in this repository: no Rust toolchain in the build image).  See INTEGRATION.md.
//! Drop-in for `texpresso::Format::{compress, decompress}` backed by libtexpresso_b200.so (sm_100a).
use std::os::raw::c_int;

pub use texpresso::{Algorithm, ColourWeights, Params, COLOUR_WEIGHTS_PERCEPTUAL, COLOUR_WEIGHTS_UNIFORM};

#[repr(C)]
struct TxpParams { algorithm: u32, weights: [f32; 3], weigh_colour_by_alpha: u32 }

extern "C" {
    fn txp_block_size(format: c_int) -> usize;
    fn txp_compressed_size(format: c_int, width: usize, height: usize) -> usize;
    fn txp_compress(format: c_int, rgba: *const u8, rgba_len: usize, width: usize, height: usize,
                    params: *const TxpParams, output: *mut u8, output_len: usize) -> c_int;
    fn txp_compress_pixels(format: c_int, pixels: *const u8, pixels_len: usize, layout: c_int, width: usize, height: usize,
                           params: *const TxpParams, output: *mut u8, output_len: usize) -> c_int;
    fn txp_decompress(format: c_int, data: *const u8, data_len: usize, width: usize, height: usize,
                      output: *mut u8, output_len: usize) -> c_int;
    fn txp_compress_block_masked(format: c_int, rgba: *const u8, mask: u32, params: *const TxpParams,
                                 output: *mut u8, output_len: usize) -> c_int;
    fn txp_decompress_block(format: c_int, block: *const u8, block_len: usize, output: *mut u8) -> c_int;
    fn txp_last_error() -> *const std::os::raw::c_char;
}

#[derive(Clone, Copy, Debug, Eq, PartialEq)]
pub enum Format { Bc1, Bc2, Bc3, Bc4, Bc5 }           // same order as texpresso::Format (lib.rs:40-46)

/// Extension: the decoded file layouts the reference's CLI expands to RGBA8 on the host (cli/src/image/png.rs:47-62);
/// `compress_pixels` expands them on the device instead.  `Rg8` = (r, g, 0, 255) has no counterpart in the reference.
#[derive(Clone, Copy, Debug, Eq, PartialEq)]
pub enum PixelLayout { L8 = 1, La8 = 2, Rgb8 = 3, Rgba8 = 4, Rg8 = 5 }

fn c_params(p: Params) -> TxpParams {
    TxpParams {
        algorithm: match p.algorithm { Algorithm::RangeFit => 0, Algorithm::ClusterFit => 1, Algorithm::IterativeClusterFit => 2 },
        weights: p.weights,
        weigh_colour_by_alpha: p.weigh_colour_by_alpha as u32,
    }
}

fn check(rc: c_int) {
    // the reference's only error path is a panic (lib.rs:295 and slice bounds); keep that contract
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(txp_last_error()) }.to_string_lossy().into_owned();
        panic!("texpresso-cuda: error {rc}: {msg}");
    }
}

impl Format {
    fn id(self) -> c_int { self as c_int }
    pub fn block_size(self) -> usize { unsafe { txp_block_size(self.id()) } }
    pub fn compressed_size(self, width: usize, height: usize) -> usize { unsafe { txp_compressed_size(self.id(), width, height) } }

    pub fn compress(self, rgba: &[u8], width: usize, height: usize, params: Params, output: &mut [u8]) {
        let p = c_params(params);
        check(unsafe { txp_compress(self.id(), rgba.as_ptr(), rgba.len(), width, height, &p, output.as_mut_ptr(), output.len()) });
    }

    /// `compress` on an image that is still in its decoded file layout (1-4 bytes per pixel).
    pub fn compress_pixels(self, pixels: &[u8], layout: PixelLayout, width: usize, height: usize, params: Params, output: &mut [u8]) {
        let p = c_params(params);
        check(unsafe { txp_compress_pixels(self.id(), pixels.as_ptr(), pixels.len(), layout as c_int, width, height, &p, output.as_mut_ptr(), output.len()) });
    }

    pub fn decompress(self, data: &[u8], width: usize, height: usize, output: &mut [u8]) {
        check(unsafe { txp_decompress(self.id(), data.as_ptr(), data.len(), width, height, output.as_mut_ptr(), output.len()) });
    }

    pub fn compress_block_masked(self, rgba: [[u8; 4]; 16], mask: u32, params: Params, output: &mut [u8]) {
        let p = c_params(params);
        check(unsafe { txp_compress_block_masked(self.id(), rgba.as_ptr() as *const u8, mask, &p, output.as_mut_ptr(), output.len()) });
    }

    pub fn decompress_block(self, block: &[u8]) -> [[u8; 4]; 16] {
        let mut out = [[0u8; 4]; 16];
        check(unsafe { txp_decompress_block(self.id(), block.as_ptr(), block.len(), out.as_mut_ptr() as *mut u8) });
        out
    }
}
