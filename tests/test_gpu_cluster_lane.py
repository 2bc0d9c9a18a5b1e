"""GPU parity tests for the lane-per-block ClusterFit search kernel (texpresso_b200/csrc/txp_cluster_lane.cuh).

In automatic mode that kernel is only taken by launches of >= 131072 blocks, so these tests force it
(txp_debug_set(0, 3)) on the small corpora the oracle finishes quickly: every dispatch class and edge case of
tests/blockgen.py, the random fuzz corpus, ragged images (edge masks), a smooth image (tiles with mixed point counts:
exercises the per-tile counting sort) and a mip chain.  Bit-exact, and identical to the warp-per-block kernel."""
import numpy as np
import pytest

from tests import oracle_lib as O
from tests import blockgen
from tests.test_gpu_fuzz import _corpus

pytestmark = pytest.mark.gpu
WEIGHTS = {"uniform": O.UNIFORM, "perceptual": O.PERCEPTUAL, "odd": (0.3, 1.7, 0.05)}


@pytest.fixture()
def TL():
    import texpresso_b200 as T
    from texpresso_b200 import _lib
    L = _lib.load()
    _lib.check(L.txp_debug_set(0, 3))          # lane-per-block search for every ClusterFit launch
    yield T
    _lib.check(L.txp_debug_set(0, 0))


def _set_variant(v):
    from texpresso_b200 import _lib
    _lib.check(_lib.load().txp_debug_set(0, v))


@pytest.mark.parametrize("awa", [False, True])
@pytest.mark.parametrize("wname", ["uniform", "perceptual", "odd"])
@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_lane_clusterfit_blocks_bit_exact(TL, fmt, wname, awa):
    T = TL
    blocks, masks, tags = blockgen.colour_cases()
    tp = T.Params(T.Algorithm(1), tuple(WEIGHTS[wname]), awa)
    got = T.compress_blocks(fmt, blocks, masks, tp)
    want = O.compress_blocks(fmt, blocks, masks, O.make_params(1, WEIGHTS[wname], awa))
    diff = np.nonzero((got != want).any(axis=1))[0]
    assert diff.size == 0, (diff.size, [(int(i), tags[i], hex(int(masks[i])), bytes(got[i]).hex(), bytes(want[i]).hex()) for i in diff[:6]])


@pytest.mark.parametrize("fmt,weights,awa", [(0, O.PERCEPTUAL, False), (0, O.UNIFORM, True), (1, O.PERCEPTUAL, False), (2, O.PERCEPTUAL, True)])
def test_lane_random_blocks_bit_exact(TL, fmt, weights, awa):
    T = TL
    blocks, masks = _corpus(7000 + fmt)
    tp = T.Params(T.Algorithm(1), tuple(weights), awa)
    got = T.compress_blocks(fmt, blocks, masks, tp)
    want = O.compress_blocks(fmt, blocks, masks, O.make_params(1, weights, awa))
    diff = np.nonzero((got != want).any(axis=1))[0]
    assert diff.size == 0, (diff.size, [(int(i), hex(int(masks[i])), bytes(blocks[i].reshape(-1)).hex(), bytes(got[i]).hex(), bytes(want[i]).hex()) for i in diff[:3]])


@pytest.mark.parametrize("kind,w,h", [("smooth", 260, 131), ("noise_alpha", 129, 67), ("smooth", 1024, 512), ("noise_opaque", 5, 3)])
@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_lane_images_bit_exact(TL, fmt, kind, w, h):
    T = TL
    from texpresso_b200 import synth
    img = synth.generate(kind, w, h, seed=11)
    tp = T.Params(T.Algorithm(1), tuple(O.PERCEPTUAL), False)
    got = T.Format(fmt).compress(img, w, h, tp)
    want = O.compress(fmt, img, w, h, O.make_params(1, O.PERCEPTUAL, False), threads=8)
    bs = 8 if fmt == 0 else 16
    nd = int((got.reshape(-1, bs) != want.reshape(-1, bs)).any(axis=1).sum())
    assert nd == 0, (fmt, kind, nd)


def test_lane_equals_warp_kernel_and_auto_threshold():
    """The three kernel structures give the same bytes; automatic mode switches at the block-count threshold."""
    import texpresso_b200 as T
    from texpresso_b200 import synth, _lib
    L = _lib.load()
    w, h = 2048, 1024                                # 131072 blocks: the automatic threshold
    for kind in ("smooth", "noise_alpha"):
        img = synth.generate(kind, w, h, seed=21)
        for fmt in (0, 2):
            outs = {}
            for name, v in (("auto", 0), ("fused", 1), ("warp", 2), ("lane", 3)):
                _lib.check(L.txp_debug_set(0, v))
                outs[name] = T.Format(fmt).compress(img, w, h, T.Params())
            _lib.check(L.txp_debug_set(0, 0))
            for name in ("fused", "warp", "lane"):
                assert np.array_equal(outs["auto"], outs[name]), (kind, fmt, name)


@pytest.mark.parametrize("fmt", [0, 2])
def test_lane_mipchain(TL, fmt):
    T = TL
    from texpresso_b200 import synth
    w, h = 200, 120
    img = synth.generate("smooth", w, h, seed=31)
    tp = T.Params(T.Algorithm(1), tuple(O.PERCEPTUAL), False)
    got = T.compress_mipchain(fmt, img, w, h, tp)
    op = O.make_params(1, O.PERCEPTUAL, False)
    want = np.concatenate([O.compress(fmt, lv, lv.shape[1], lv.shape[0], op) for lv in T.generate_mips(img, w, h)])
    assert np.array_equal(got, want)
