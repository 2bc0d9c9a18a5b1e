"""ctypes binding of the CPU oracle (oracle/txp_oracle.c).  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
import ctypes, pathlib, subprocess
import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent.parent
SO = ROOT / "oracle" / "_build" / "libtxp_oracle.so"

BC1, BC2, BC3, BC4, BC5 = range(5)
RANGE_FIT, CLUSTER_FIT, ITERATIVE_CLUSTER_FIT = range(3)
UNIFORM = (1.0, 1.0, 1.0)
PERCEPTUAL = (0.2126, 0.7152, 0.0722)


class Params(ctypes.Structure):
    _fields_ = [("algorithm", ctypes.c_uint32), ("weights", ctypes.c_float * 3), ("weigh_colour_by_alpha", ctypes.c_uint32)]


class Stats(ctypes.Structure):
    _fields_ = [("blocks", ctypes.c_uint64), ("single_blocks", ctypes.c_uint64), ("range_blocks", ctypes.c_uint64),
                ("cluster_blocks", ctypes.c_uint64), ("cand3", ctypes.c_uint64), ("cand4", ctypes.c_uint64),
                ("orderings3", ctypes.c_uint64), ("orderings4", ctypes.c_uint64), ("count_hist", ctypes.c_uint64 * 17)]


def make_params(algorithm=CLUSTER_FIT, weights=PERCEPTUAL, weigh_colour_by_alpha=False):
    return Params(int(algorithm), (ctypes.c_float * 3)(*weights), 1 if weigh_colour_by_alpha else 0)


_lib = None


def build():
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle")], check=True)


def lib():
    global _lib
    if _lib is None:
        if not SO.exists():
            build()
        L = ctypes.CDLL(str(SO))
        u8p = ctypes.c_void_p
        L.txo_block_size.restype = ctypes.c_size_t
        L.txo_block_size.argtypes = [ctypes.c_int]
        L.txo_compressed_size.restype = ctypes.c_size_t
        L.txo_compressed_size.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t]
        L.txo_compress.restype = ctypes.c_int
        L.txo_compress.argtypes = [ctypes.c_int, u8p, ctypes.c_size_t, ctypes.c_size_t, ctypes.POINTER(Params), u8p,
                                   ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(Stats)]
        L.txo_decompress.restype = ctypes.c_int
        L.txo_decompress.argtypes = [ctypes.c_int, u8p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, u8p, ctypes.c_size_t]
        L.txo_compress_block_masked.restype = None
        L.txo_compress_block_masked.argtypes = [ctypes.c_int, u8p, ctypes.c_uint32, ctypes.POINTER(Params), u8p]
        L.txo_decompress_block.restype = None
        L.txo_decompress_block.argtypes = [ctypes.c_int, u8p, u8p]
        L.txo_colour_block_error.restype = ctypes.c_double
        L.txo_colour_block_error.argtypes = [ctypes.c_int, u8p, ctypes.c_uint32, ctypes.POINTER(Params), u8p]
        L.txo_compress_blocks.restype = None
        L.txo_compress_blocks.argtypes = [ctypes.c_int, u8p, u8p, ctypes.c_size_t, ctypes.POINTER(Params), u8p]
        L.txo_decompress_blocks.restype = None
        L.txo_decompress_blocks.argtypes = [ctypes.c_int, u8p, ctypes.c_size_t, u8p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def compressed_size(fmt, w, h):
    return lib().txo_compressed_size(fmt, w, h)


def compress(fmt, rgba, w, h, params=None, out_len=None, threads=1, want_stats=False):
    rgba = np.ascontiguousarray(rgba, dtype=np.uint8).reshape(-1)
    assert rgba.size >= w * h * 4
    params = params or make_params()
    n = compressed_size(fmt, w, h) if out_len is None else out_len
    out = np.zeros(n, dtype=np.uint8)
    st = Stats()
    rc = lib().txo_compress(fmt, _ptr(rgba), w, h, ctypes.byref(params), _ptr(out), n, threads, ctypes.byref(st) if want_stats else None)
    if rc != 0:
        raise ValueError("oracle compress rejected the call (the reference would panic)")
    return (out, st) if want_stats else out


def decompress(fmt, data, w, h, out_len=None):
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    n = w * h * 4 if out_len is None else out_len
    out = np.zeros(n, dtype=np.uint8)
    rc = lib().txo_decompress(fmt, _ptr(data), data.size, w, h, _ptr(out), n)
    if rc != 0:
        raise ValueError("oracle decompress rejected the call (the reference would panic)")
    return out


def compress_block_masked(fmt, rgba64, mask, params=None):
    rgba64 = np.ascontiguousarray(rgba64, dtype=np.uint8).reshape(-1)
    assert rgba64.size == 64
    params = params or make_params()
    out = np.zeros(lib().txo_block_size(fmt), dtype=np.uint8)
    lib().txo_compress_block_masked(fmt, _ptr(rgba64), mask, ctypes.byref(params), _ptr(out))
    return out


def decompress_block(fmt, block):
    block = np.ascontiguousarray(block, dtype=np.uint8).reshape(-1)
    out = np.zeros(64, dtype=np.uint8)
    lib().txo_decompress_block(fmt, _ptr(block), _ptr(out))
    return out


def colour_block_error(fmt, rgba64, mask, params, block8):
    rgba64 = np.ascontiguousarray(rgba64, dtype=np.uint8).reshape(-1)
    block8 = np.ascontiguousarray(block8, dtype=np.uint8).reshape(-1)
    return lib().txo_colour_block_error(fmt, _ptr(rgba64), mask, ctypes.byref(params), _ptr(block8))


def compress_blocks(fmt, rgba_blocks, masks, params=None):
    rgba_blocks = np.ascontiguousarray(rgba_blocks, dtype=np.uint8).reshape(-1)
    masks = np.ascontiguousarray(masks, dtype=np.uint32).reshape(-1)
    n = masks.size
    assert rgba_blocks.size == 64 * n
    params = params or make_params()
    out = np.zeros(n * lib().txo_block_size(fmt), dtype=np.uint8)
    lib().txo_compress_blocks(fmt, _ptr(rgba_blocks), _ptr(masks), n, ctypes.byref(params), _ptr(out))
    return out.reshape(n, -1)


def decompress_blocks(fmt, blocks):
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1)
    n = blocks.size // lib().txo_block_size(fmt)
    out = np.zeros(n * 64, dtype=np.uint8)
    lib().txo_decompress_blocks(fmt, _ptr(blocks), n, _ptr(out))
    return out.reshape(n, 16, 4)
