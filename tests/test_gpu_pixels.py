"""txp_compress_pixels: L8 / LA8 / RGB8 images expanded to RGBA8 on the device exactly as the
reference's CLI expands them on the host (cli/src/image/png.rs:47-62, jpeg.rs:42-52), then Format::compress.
Bit-exact against the oracle run on the host-expanded image."""
import numpy as np
import pytest

from tests import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h", [(133, 71), (1024, 512), (4, 4), (3, 9)])
@pytest.mark.parametrize("fmt,alg,layout", [(3, 1, 1), (4, 1, 2), (4, 1, 5), (0, 0, 3), (2, 1, 2), (0, 1, 1), (1, 0, 3), (2, 1, 4), (0, 1, 5)])
def test_compress_pixels_matches_host_expansion(fmt, alg, layout, w, h):
    import texpresso_b200 as T
    rng = np.random.default_rng(100 * fmt + 10 * layout + w)
    pix = rng.integers(0, 256, size=(h, w, 2 if layout == 5 else layout), dtype=np.uint8)
    if layout == 2:
        pix[..., 1] = np.where(rng.random((h, w)) < 0.5, 255, pix[..., 1])
    tp = T.Params(T.Algorithm(alg), tuple(O.PERCEPTUAL), False)
    got = T.compress_pixels(fmt, pix, w, h, tp, layout=layout)
    rgba = T.expand_pixels(pix, w, h, layout=layout)
    if layout == 5:
        assert np.array_equal(rgba[..., :2], pix) and (rgba[..., 2] == 0).all() and (rgba[..., 3] == 255).all()
    if layout == 1:
        assert (rgba[..., 0] == rgba[..., 2]).all() and (rgba[..., 3] == 255).all()
    want = O.compress(fmt, rgba, w, h, O.make_params(alg, O.PERCEPTUAL, False), threads=8)
    assert np.array_equal(got, want)
    assert np.array_equal(got, T.Format(fmt).compress(rgba, w, h, tp))


def test_compress_pixels_argument_errors():
    import texpresso_b200 as T
    from texpresso_b200 import _lib
    import ctypes
    L = _lib.load()
    pix = np.zeros(16, np.uint8)
    out = np.zeros(8, np.uint8)
    cp = T.Params()._c()
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    assert L.txp_compress_pixels(3, p(pix), 15, 1, 4, 4, ctypes.byref(cp), p(out), 8) != 0      # pixels too short
    assert L.txp_compress_pixels(3, p(pix), 16, 6, 4, 4, ctypes.byref(cp), p(out), 8) != 0      # bad layout
    assert L.txp_compress_pixels(3, p(pix), 16, 0, 4, 4, ctypes.byref(cp), p(out), 8) != 0
    assert L.txp_compress_pixels(3, p(pix), 16, 1, 4, 4, ctypes.byref(cp), p(out), 7) != 0      # output too short
    assert L.txp_compress_pixels(3, p(pix), 16, 1, 4, 4, ctypes.byref(cp), p(out), 8) == 0
