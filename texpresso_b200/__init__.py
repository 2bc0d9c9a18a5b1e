"""texpresso_b200 -- B200 (sm_100a) BC1..BC5 encoder/decoder with the library API of jansol/texpresso.

Host-side mirror of the reference's public interface (reference lib/src/lib.rs:38-336), same names and
argument meaning, over the C ABI in include/texpresso_b200.h:

    Format.Bc1 .. Format.Bc5   .compress / .decompress / .compressed_size / .block_size /
                               .compress_block_masked / .decompress_block
    Algorithm.RangeFit | ClusterFit | IterativeClusterFit,  Params(algorithm, weights, weigh_colour_by_alpha)
    COLOUR_WEIGHTS_UNIFORM, COLOUR_WEIGHTS_PERCEPTUAL, num_blocks

Where the reference panics (short buffers, zero width) these raise TexpressoError.  All arithmetic runs in
the CUDA kernels of csrc/; there is no CPU path.
"""
import ctypes
import enum
from dataclasses import dataclass, field

import numpy as np

from ._lib import CParams, TexpressoError, check, load

__all__ = ["Format", "Algorithm", "Params", "ColourWeights", "COLOUR_WEIGHTS_UNIFORM", "COLOUR_WEIGHTS_PERCEPTUAL",
           "num_blocks", "TexpressoError", "shard_rows", "compress_multi", "compress_batch", "decompress_multi", "decompress_batch", "compress_blocks",
           "decompress_blocks", "mip_levels", "generate_mips", "compress_mipchain", "compress_batch_mips", "device_count", "set_device", "kernel_launches", "version"]

ColourWeights = tuple
COLOUR_WEIGHTS_UNIFORM = (1.0, 1.0, 1.0)                 # lib.rs:71
COLOUR_WEIGHTS_PERCEPTUAL = (0.2126, 0.7152, 0.0722)     # lib.rs:74


class Algorithm(enum.IntEnum):                           # lib.rs:49-59
    RangeFit = 0
    ClusterFit = 1
    IterativeClusterFit = 2

    @classmethod
    def default(cls):                                    # lib.rs:61-65
        return cls.ClusterFit


@dataclass(frozen=True)
class Params:                                            # lib.rs:76-100
    algorithm: Algorithm = Algorithm.ClusterFit
    weights: tuple = COLOUR_WEIGHTS_PERCEPTUAL
    weigh_colour_by_alpha: bool = False

    def _c(self):
        return CParams(int(self.algorithm), (ctypes.c_float * 3)(*[float(x) for x in self.weights]),
                       1 if self.weigh_colour_by_alpha else 0)


def num_blocks(size):                                    # lib.rs:103-105
    return (int(size) + 3) // 4


def _u8(a, name):
    a = np.asarray(a)
    if a.dtype != np.uint8:
        raise TypeError(f"{name} must be uint8")
    return np.ascontiguousarray(a).reshape(-1)


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def _out(output, name="output"):
    """A caller-supplied output must be written in place: contiguous uint8, no temporary copy."""
    out = _u8(output, name)
    if out.size and not np.shares_memory(out, output):           # (an empty array shares no memory with anything)
        raise ValueError(f"{name} must be a contiguous uint8 array")
    return out


class Format(enum.IntEnum):                              # lib.rs:39-46
    Bc1 = 0
    Bc2 = 1
    Bc3 = 2
    Bc4 = 3
    Bc5 = 4

    def block_size(self):                                # lib.rs:159-168
        return load().txp_block_size(int(self))

    def compressed_size(self, width, height):            # lib.rs:175-179
        return load().txp_compressed_size(int(self), width, height)

    def compress(self, rgba, width, height, params=None, output=None):
        """Format::compress (lib.rs:287-335).  `output` (uint8 array) is filled in place and returned; if
        omitted an array of compressed_size bytes is allocated."""
        rgba = _u8(rgba, "rgba")
        params = params or Params()
        if output is None:
            output = np.empty(self.compressed_size(width, height), dtype=np.uint8)
        out = _out(output)
        cp = params._c()
        check(load().txp_compress(int(self), _ptr(rgba), rgba.size, width, height, ctypes.byref(cp), _ptr(out), out.size))
        return output

    def decompress(self, data, width, height, output=None):
        """Format::decompress (lib.rs:124-156)."""
        data = _u8(data, "data")
        if output is None:
            output = np.empty(width * height * 4, dtype=np.uint8)
        out = _out(output)
        check(load().txp_decompress(int(self), _ptr(data), data.size, width, height, _ptr(out), out.size))
        return output

    def compress_block_masked(self, rgba, mask, params=None, output=None):
        """Format::compress_block_masked (lib.rs:188-234): rgba is 16x4 bytes, mask bit i = pixel i valid."""
        rgba = _u8(rgba, "rgba")
        if rgba.size != 64:
            raise ValueError("rgba must hold 16 RGBA pixels")
        params = params or Params()
        if output is None:
            output = np.empty(self.block_size(), dtype=np.uint8)
        out = _out(output)
        cp = params._c()
        check(load().txp_compress_block_masked(int(self), _ptr(rgba), int(mask) & 0xFFFFFFFF, ctypes.byref(cp), _ptr(out), out.size))
        return output

    def decompress_block(self, block):
        """Format::decompress_block (lib.rs:240-277) -> (16, 4) uint8."""
        block = _u8(block, "block")
        out = np.empty(64, dtype=np.uint8)
        check(load().txp_decompress_block(int(self), _ptr(block), block.size, _ptr(out)))
        return out.reshape(16, 4)


# ---- extensions beyond the reference surface (sharding / batching / bulk block calls) -------------------
def compress_blocks(fmt, rgba_blocks, masks, params=None):
    """n independent compress_block_masked calls in one launch: rgba_blocks (n,64) uint8, masks (n,) uint32."""
    rgba_blocks = _u8(rgba_blocks, "rgba_blocks")
    masks = np.ascontiguousarray(masks, dtype=np.uint32).reshape(-1)
    n = masks.size
    if rgba_blocks.size != 64 * n:
        raise ValueError("rgba_blocks must be n x 64 bytes")
    params = params or Params()
    out = np.empty(n * Format(fmt).block_size(), dtype=np.uint8)
    cp = params._c()
    check(load().txp_compress_blocks(int(fmt), _ptr(rgba_blocks), _ptr(masks), n, ctypes.byref(cp), _ptr(out)))
    return out.reshape(n, -1)


def decompress_blocks(fmt, blocks):
    blocks = _u8(blocks, "blocks")
    bs = Format(fmt).block_size()
    n = blocks.size // bs
    out = np.empty(n * 64, dtype=np.uint8)
    check(load().txp_decompress_blocks(int(fmt), _ptr(blocks), n, _ptr(out)))
    return out.reshape(n, 16, 4)


PIXELS_L8, PIXELS_LA8, PIXELS_RGB8, PIXELS_RGBA8, PIXELS_RG8 = 1, 2, 3, 4, 5


def compress_pixels(fmt, pixels, width, height, params=None, output=None, layout=None):
    """Format.compress on an image in its decoded file layout: pixels is (h, w), (h, w, 1..4) or flat with 1-4 bytes per
    pixel (L8, LA8, RGB8, RGBA8; layout=PIXELS_RG8 reads 2-byte pixels as (r, g, 0, 255) instead of gray + alpha).
    Expansion to RGBA8 (cli/src/image/png.rs:47-62) happens on the device."""
    pixels = _u8(pixels, "pixels")
    npix = int(width) * int(height)
    if npix == 0 or pixels.size % npix or not 1 <= pixels.size // npix <= 4:
        raise ValueError("pixels must hold 1, 2, 3 or 4 bytes per pixel")
    layout = pixels.size // npix if layout is None else int(layout)
    params = params or Params()
    if output is None:
        output = np.empty(Format(fmt).compressed_size(width, height), dtype=np.uint8)
    out = _out(output)
    cp = params._c()
    check(load().txp_compress_pixels(int(fmt), _ptr(pixels), pixels.size, layout, width, height, ctypes.byref(cp), _ptr(out), out.size))
    return output


def expand_pixels(pixels, width, height, layout=None):
    """Host (numpy) statement of the expansion, for tests and tools (cli/src/image/png.rs:47-62)."""
    pixels = np.asarray(pixels, dtype=np.uint8).reshape(height, width, -1)
    c = pixels.shape[2]
    out = np.empty((height, width, 4), np.uint8)
    if c == 1: out[..., :3] = pixels; out[..., 3] = 255
    elif c == 2 and layout == PIXELS_RG8: out[..., :2] = pixels; out[..., 2] = 0; out[..., 3] = 255
    elif c == 2: out[..., :3] = pixels[..., :1]; out[..., 3] = pixels[..., 1]
    elif c == 3: out[..., :3] = pixels; out[..., 3] = 255
    else: out[...] = pixels
    return out


def _cuda_u8(t, name, need):
    """A torch CUDA tensor as (pointer, bytes): uint8, contiguous, at least `need` bytes.  (torch is plumbing for device memory only.)"""
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.uint8 or not t.is_contiguous():
        raise TypeError(f"{name} must be a contiguous uint8 CUDA tensor")
    if t.numel() < need:
        raise ValueError(f"{name}: {t.numel()} bytes, {need} needed")
    return ctypes.c_void_p(t.data_ptr()), t.numel()


def compress_device(fmt, rgba, width, height, params=None, output=None, stream=None):
    """txp_compress_device: Format::compress (lib.rs:287-335) for callers whose RGBA8 image already lives in HBM.  rgba / output are torch
    uint8 CUDA tensors on the current device; the kernels are enqueued on `stream` (a torch stream; default: torch's current stream) and the
    call returns without waiting -- synchronise the stream before reading `output` on the host."""
    import torch
    F = Format(fmt)
    params = params or Params()
    need = F.compressed_size(width, height)
    if output is None:
        output = torch.empty(need, dtype=torch.uint8, device=rgba.device)
    src, _ = _cuda_u8(rgba, "rgba", int(width) * int(height) * 4)
    dst, out_len = _cuda_u8(output, "output", need)
    st = (stream or torch.cuda.current_stream(rgba.device)).cuda_stream
    cp = params._c()
    check(load().txp_compress_device(int(F), src, width, height, ctypes.byref(cp), dst, out_len, ctypes.c_void_p(st)))
    return output


def decompress_device(fmt, data, width, height, output=None, stream=None):
    """txp_decompress_device: Format::decompress (lib.rs:124-156) on torch uint8 CUDA tensors; asynchronous like compress_device."""
    import torch
    F = Format(fmt)
    need = int(width) * int(height) * 4
    if output is None:
        output = torch.empty(need, dtype=torch.uint8, device=data.device)
    src, _ = _cuda_u8(data, "data", F.compressed_size(width, height))
    dst, out_len = _cuda_u8(output, "output", need)
    st = (stream or torch.cuda.current_stream(data.device)).cuda_stream
    check(load().txp_decompress_device(int(F), src, width, height, dst, out_len, ctypes.c_void_p(st)))
    return output


def shard_rows(height, rank, world):
    """Block-row range [begin, end) owned by `rank` of `world` (reference grain: one block row, lib.rs:300-305)."""
    a, b = ctypes.c_size_t(), ctypes.c_size_t()
    load().txp_shard_rows(height, rank, world, ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


def compress_multi(fmt, rgba, width, height, params=None, n_gpus=1, output=None):
    rgba = _u8(rgba, "rgba")
    params = params or Params()
    if output is None:
        output = np.empty(Format(fmt).compressed_size(width, height), dtype=np.uint8)
    out = _out(output)
    cp = params._c()
    check(load().txp_compress_multi(int(fmt), _ptr(rgba), rgba.size, width, height, ctypes.byref(cp), _ptr(out), out.size, n_gpus))
    return output


def decompress_multi(fmt, data, width, height, n_gpus=1, output=None):
    """Format.decompress with the block rows sharded over n_gpus devices (reference grain: lib.rs:128-134)."""
    data = _u8(data, "data")
    if output is None:
        output = np.empty(int(width) * int(height) * 4, dtype=np.uint8)
    out = _out(output)
    check(load().txp_decompress_multi(int(fmt), _ptr(data), data.size, width, height, _ptr(out), out.size, n_gpus))
    return output


def _batch_args(textures, need_in, need_out, outputs):
    """Checks every (array, width, height) of a batch against the sizes the C ABI will read / write (it takes no lengths)."""
    n = len(textures)
    arrs = []
    for t, (a, w, h) in enumerate(textures):
        a = _u8(a, f"textures[{t}]")
        if int(w) <= 0 or int(h) < 0:
            raise ValueError(f"textures[{t}]: bad dimensions {w}x{h}")
        if a.size < need_in(int(w), int(h)):
            raise ValueError(f"textures[{t}]: {a.size} bytes, {need_in(int(w), int(h))} needed for {w}x{h}")
        arrs.append(a)
    if outputs is None:
        outputs = [np.empty(need_out(int(t[1]), int(t[2])), dtype=np.uint8) for t in textures]
    if len(outputs) != n:
        raise ValueError("outputs must hold one array per texture")
    outs = []
    for t, o in enumerate(outputs):
        oo = _out(o, f"outputs[{t}]")
        if oo.size < need_out(int(textures[t][1]), int(textures[t][2])):
            raise ValueError(f"outputs[{t}]: {oo.size} bytes, {need_out(int(textures[t][1]), int(textures[t][2]))} needed")
        outs.append(oo)
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    ins = (vp * n)(*[a.ctypes.data for a in arrs])
    ous = (vp * n)(*[o.ctypes.data for o in outs])
    ws = (sz * n)(*[int(t[1]) for t in textures])
    hs = (sz * n)(*[int(t[2]) for t in textures])
    return n, arrs, outputs, ins, ous, ws, hs


def decompress_batch(fmt, textures, n_gpus=1, outputs=None):
    """textures: list of (compressed uint8 array, width, height); texture t is decoded on device t % n_gpus."""
    f = Format(fmt)
    n, _keep, outputs, ins, ous, ws, hs = _batch_args(textures, lambda w, h: f.compressed_size(w, h), lambda w, h: 4 * w * h, outputs)
    check(load().txp_decompress_batch(int(fmt), ins, ws, hs, n, ous, n_gpus))
    return outputs


def compress_batch(fmt, textures, params=None, n_gpus=1, outputs=None):
    """textures: list of (rgba uint8 array, width, height); texture t is encoded on device t % n_gpus.  Textures of up to
    32 MiB are pipelined through the device's slots (copies and kernels of neighbouring textures overlap)."""
    params = params or Params()
    f = Format(fmt)
    n, _keep, outputs, ins, ous, ws, hs = _batch_args(textures, lambda w, h: 4 * w * h, lambda w, h: f.compressed_size(w, h), outputs)
    cp = params._c()
    check(load().txp_compress_batch(int(fmt), ins, ws, hs, n, ctypes.byref(cp), ous, n_gpus))
    return outputs


def mip_levels(width, height):
    """[(w, h), ...] of the mip chain down to 1x1 (each dimension halves, floor, min 1)."""
    out, w, h = [], int(width), int(height)
    while True:
        out.append((w, h))
        if w == 1 and h == 1:
            return out
        w, h = max(1, w // 2), max(1, h // 2)


def generate_mips(rgba, width, height):
    """Host (numpy) statement of the device mip filter, for tests and tools: 2x2 box, (a+b+c+d+2)>>2, edge-clamped."""
    lv = [np.asarray(rgba, dtype=np.uint8).reshape(height, width, 4)]
    for (w, h) in mip_levels(width, height)[1:]:
        s = lv[-1].astype(np.uint16)
        sh, sw = s.shape[:2]
        ys0 = np.minimum(2 * np.arange(h), sh - 1); ys1 = np.minimum(2 * np.arange(h) + 1, sh - 1)
        xs0 = np.minimum(2 * np.arange(w), sw - 1); xs1 = np.minimum(2 * np.arange(w) + 1, sw - 1)
        acc = s[ys0][:, xs0] + s[ys0][:, xs1] + s[ys1][:, xs0] + s[ys1][:, xs1] + 2
        lv.append((acc >> 2).astype(np.uint8))
    return lv


def compress_mipchain(fmt, rgba, width, height, params=None):
    """Encodes a texture and its device-generated mip chain; returns the concatenated blocks (level 0 first)."""
    rgba = _u8(rgba, "rgba")
    params = params or Params()
    out = np.empty(load().txp_mipchain_compressed_size(int(fmt), width, height), dtype=np.uint8)
    cp = params._c()
    check(load().txp_compress_mipchain(int(fmt), _ptr(rgba), rgba.size, width, height, ctypes.byref(cp), _ptr(out), out.size))
    return out


def compress_batch_mips(fmt, textures, params=None, n_gpus=1, outputs=None):
    """textures: list of (rgba uint8 array, width, height); each is encoded with its full mip chain on device t % n_gpus."""
    params = params or Params()
    size = lambda w, h: load().txp_mipchain_compressed_size(int(fmt), w, h)
    for t, (_a, w, h) in enumerate(textures):
        if int(w) <= 0 or int(h) <= 0:
            raise ValueError(f"textures[{t}]: bad dimensions {w}x{h}")
    n, _keep, outputs, ins, ous, ws, hs = _batch_args(textures, lambda w, h: 4 * w * h, size, outputs)
    cp = params._c()
    check(load().txp_compress_batch_mips(int(fmt), ins, ws, hs, n, ctypes.byref(cp), ous, n_gpus))
    return outputs


def device_count():
    n = load().txp_device_count()
    if n < 0:
        check(n)
    return n


def set_device(device):
    check(load().txp_set_device(device))


def kernel_launches():
    return load().txp_kernel_launches()


def version():
    return load().txp_version().decode()
