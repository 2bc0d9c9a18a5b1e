"""ctypes loader for libtexpresso_b200.so (the C ABI in include/texpresso_b200.h).

There is no fallback: if the shared library is missing the import fails, and if no CUDA device is
usable every data call raises TexpressoError."""
import ctypes, os, pathlib

HERE = pathlib.Path(__file__).resolve().parent
SO_PATH = pathlib.Path(os.environ.get("TEXPRESSO_B200_LIB", HERE / "libtexpresso_b200.so"))

# every symbol include/texpresso_b200.h declares
EXPORTS = [
    "txp_num_blocks", "txp_block_size", "txp_compressed_size", "txp_compress", "txp_compress_pixels", "txp_decompress",
    "txp_compress_block_masked", "txp_decompress_block", "txp_compress_blocks", "txp_decompress_blocks",
    "txp_compress_device", "txp_decompress_device", "txp_shard_rows", "txp_compress_multi", "txp_compress_batch", "txp_decompress_multi", "txp_decompress_batch",
    "txp_mip_levels", "txp_mipchain_compressed_size", "txp_compress_mipchain", "txp_compress_batch_mips",
    "txp_device_count", "txp_set_device", "txp_last_error", "txp_kernel_launches", "txp_version", "txp_debug_set", "txp_debug_get", "txp_measure_fp32_issue", "txp_debug_host_copy", "txp_debug_plan",
]


class TexpressoError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"texpresso_b200 error {code}: {message}")
        self.code = code


class CParams(ctypes.Structure):
    _fields_ = [("algorithm", ctypes.c_uint32), ("weights", ctypes.c_float * 3), ("weigh_colour_by_alpha", ctypes.c_uint32)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not SO_PATH.exists():
        raise ImportError(f"{SO_PATH} is missing: build it with `python -m texpresso_b200.build` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = ctypes.CDLL(str(SO_PATH))
    sz, vp, ci, u32 = ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32
    pp = ctypes.POINTER(CParams)
    sig = {
        "txp_num_blocks": (sz, [sz]),
        "txp_block_size": (sz, [ci]),
        "txp_compressed_size": (sz, [ci, sz, sz]),
        "txp_compress": (ci, [ci, vp, sz, sz, sz, pp, vp, sz]),
        "txp_compress_pixels": (ci, [ci, vp, sz, ci, sz, sz, pp, vp, sz]),
        "txp_decompress": (ci, [ci, vp, sz, sz, sz, vp, sz]),
        "txp_compress_block_masked": (ci, [ci, vp, u32, pp, vp, sz]),
        "txp_decompress_block": (ci, [ci, vp, sz, vp]),
        "txp_compress_blocks": (ci, [ci, vp, vp, sz, pp, vp]),
        "txp_decompress_blocks": (ci, [ci, vp, sz, vp]),
        "txp_compress_device": (ci, [ci, vp, sz, sz, pp, vp, sz, vp]),
        "txp_decompress_device": (ci, [ci, vp, sz, sz, vp, sz, vp]),
        "txp_shard_rows": (None, [sz, ci, ci, ctypes.POINTER(sz), ctypes.POINTER(sz)]),
        "txp_compress_multi": (ci, [ci, vp, sz, sz, sz, pp, vp, sz, ci]),
        "txp_compress_batch": (ci, [ci, ctypes.POINTER(vp), ctypes.POINTER(sz), ctypes.POINTER(sz), sz, pp, ctypes.POINTER(vp), ci]),
        "txp_decompress_multi": (ci, [ci, vp, sz, sz, sz, vp, sz, ci]),
        "txp_decompress_batch": (ci, [ci, ctypes.POINTER(vp), ctypes.POINTER(sz), ctypes.POINTER(sz), sz, ctypes.POINTER(vp), ci]),
        "txp_mip_levels": (ci, [sz, sz]),
        "txp_mipchain_compressed_size": (sz, [ci, sz, sz]),
        "txp_compress_mipchain": (ci, [ci, vp, sz, sz, sz, pp, vp, sz]),
        "txp_compress_batch_mips": (ci, [ci, ctypes.POINTER(vp), ctypes.POINTER(sz), ctypes.POINTER(sz), sz, pp, ctypes.POINTER(vp), ci]),
        "txp_device_count": (ci, []),
        "txp_set_device": (ci, [ci]),
        "txp_last_error": (ctypes.c_char_p, []),
        "txp_kernel_launches": (ctypes.c_uint64, []),
        "txp_version": (ctypes.c_char_p, []),
        "txp_debug_set": (ci, [ci, ci]),
        "txp_debug_get": (ci, [ci, ctypes.POINTER(ctypes.c_uint64)]),
        "txp_measure_fp32_issue": (ci, [ctypes.POINTER(ctypes.c_double)]),
        "txp_debug_host_copy": (ci, [vp, vp, sz]),
        "txp_debug_plan": (ci, [ci, pp, sz, sz, ci, ctypes.POINTER(sz), sz, ctypes.POINTER(sz), ctypes.POINTER(ctypes.c_uint64)]),
    }
    for name in EXPORTS:
        fn = getattr(L, name)          # AttributeError if the library does not export it
        fn.restype, fn.argtypes = sig[name]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise TexpressoError(rc, load().txp_last_error().decode("utf-8", "replace"))
