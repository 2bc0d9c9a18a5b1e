"""GPU parity tests for the lane-per-block ClusterFit / IterativeClusterFit search kernels (texpresso_b200/csrc/txp_cluster_lane.cuh).

In automatic mode those kernels are only taken by launches of >= 262144 blocks, so these tests force them
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
@pytest.mark.parametrize("alg", [1, 2])
@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_lane_clusterfit_blocks_bit_exact(TL, fmt, alg, wname, awa):
    T = TL
    blocks, masks, tags = blockgen.colour_cases()
    tp = T.Params(T.Algorithm(alg), tuple(WEIGHTS[wname]), awa)
    got = T.compress_blocks(fmt, blocks, masks, tp)
    want = O.compress_blocks(fmt, blocks, masks, O.make_params(alg, WEIGHTS[wname], awa))
    diff = np.nonzero((got != want).any(axis=1))[0]
    assert diff.size == 0, (diff.size, [(int(i), tags[i], hex(int(masks[i])), bytes(got[i]).hex(), bytes(want[i]).hex()) for i in diff[:6]])


@pytest.mark.parametrize("alg", [1, 2])
@pytest.mark.parametrize("fmt,weights,awa", [(0, O.PERCEPTUAL, False), (0, O.UNIFORM, True), (1, O.PERCEPTUAL, False), (2, O.PERCEPTUAL, True)])
def test_lane_random_blocks_bit_exact(TL, fmt, weights, awa, alg):
    T = TL
    blocks, masks = _corpus(7000 + fmt + 100 * alg)
    tp = T.Params(T.Algorithm(alg), tuple(weights), awa)
    got = T.compress_blocks(fmt, blocks, masks, tp)
    want = O.compress_blocks(fmt, blocks, masks, O.make_params(alg, weights, awa))
    diff = np.nonzero((got != want).any(axis=1))[0]
    assert diff.size == 0, (diff.size, [(int(i), hex(int(masks[i])), bytes(blocks[i].reshape(-1)).hex(), bytes(got[i]).hex(), bytes(want[i]).hex()) for i in diff[:3]])


@pytest.mark.parametrize("kind,w,h", [("smooth", 260, 131), ("noise_alpha", 129, 67), ("smooth", 1024, 512), ("noise_opaque", 5, 3)])
@pytest.mark.parametrize("alg", [1, 2])
@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_lane_images_bit_exact(TL, fmt, alg, kind, w, h):
    T = TL
    from texpresso_b200 import synth
    img = synth.generate(kind, w, h, seed=11)
    tp = T.Params(T.Algorithm(alg), tuple(O.PERCEPTUAL), False)
    got = T.Format(fmt).compress(img, w, h, tp)
    want = O.compress(fmt, img, w, h, O.make_params(alg, O.PERCEPTUAL, False), threads=8)
    bs = 8 if fmt == 0 else 16
    nd = int((got.reshape(-1, bs) != want.reshape(-1, bs)).any(axis=1).sum())
    assert nd == 0, (fmt, kind, nd)


def test_lane_equals_warp_kernel_and_auto_threshold():
    """The kernel structures give the same bytes on a single device-resident launch of 262144 blocks (the automatic
    threshold), and automatic mode takes the lane kernels there (one more launch per call than the warp structure for BC1
    IterativeClusterFit: setup + compress3 + compress4)."""
    import ctypes
    import torch
    import texpresso_b200 as T
    from texpresso_b200 import synth, _lib
    L = _lib.load()
    w = h = 2048
    for kind in ("smooth", "noise_alpha"):
        img = synth.generate(kind, w, h, seed=21)
        d = torch.from_numpy(img.reshape(-1)).cuda()
        for fmt, alg in ((0, 1), (2, 1), (0, 2), (1, 2)):
            bs = 8 if fmt == 0 else 16
            cp = T.Params(T.Algorithm(alg))._c()
            outs, launches = {}, {}
            for name, v in (("auto", 0), ("fused", 1), ("warp", 2), ("lane", 3), ("hybrid", 4)):
                _lib.check(L.txp_debug_set(2, 100 if name == "hybrid" else 0))   # hybrid: every partly filled last round goes to the warp kernel
                _lib.check(L.txp_debug_set(0, v))
                out = torch.zeros((w // 4) * (h // 4) * bs, dtype=torch.uint8, device="cuda")
                n0 = T.kernel_launches()
                _lib.check(L.txp_compress_device(fmt, ctypes.c_void_p(d.data_ptr()), w, h, ctypes.byref(cp), ctypes.c_void_p(out.data_ptr()), out.numel(), None))
                torch.cuda.synchronize()
                launches[name] = T.kernel_launches() - n0
                outs[name] = out.cpu().numpy()
            _lib.check(L.txp_debug_set(0, 0))
            _lib.check(L.txp_debug_set(2, 0))
            for name in ("fused", "warp", "lane", "hybrid"):
                assert np.array_equal(outs["auto"], outs[name]), (kind, fmt, alg, name)
            assert launches["auto"] == (launches["lane"] if alg == 1 else launches["warp"]), launches   # iterative: lane from 786432 blocks
            assert launches["fused"] == 1 and launches["warp"] == 2, launches


@pytest.mark.parametrize("alg", [1, 2])
@pytest.mark.parametrize("fmt", [0, 2])
def test_lane_mipchain(TL, fmt, alg):
    T = TL
    from texpresso_b200 import synth
    w, h = 200, 120
    img = synth.generate("smooth", w, h, seed=31)
    tp = T.Params(T.Algorithm(alg), tuple(O.PERCEPTUAL), False)
    got = T.compress_mipchain(fmt, img, w, h, tp)
    op = O.make_params(alg, O.PERCEPTUAL, False)
    want = np.concatenate([O.compress(fmt, lv, lv.shape[1], lv.shape[0], op) for lv in T.generate_mips(img, w, h)])
    assert np.array_equal(got, want)


def test_lane_chunked_launch_equals_warp_kernel():
    """More than 4 Mi blocks in one device-resident call: the lane path runs setup + search per chunk of 4 194 304 blocks
    (scratch is 292 B per block); the second chunk starts at a non-zero block offset."""
    import ctypes
    import torch
    import texpresso_b200 as T
    from texpresso_b200 import synth, _lib
    L = _lib.load()
    w, h = 8192, 8200                                  # 2048 x 2050 = 4 198 400 blocks
    d = torch.from_numpy(synth.generate("smooth", w, h, seed=41).reshape(-1)).cuda()
    for fmt, alg in ((0, 1), (2, 2)):
        bs = 8 if fmt == 0 else 16
        cp = T.Params(T.Algorithm(alg))._c()
        outs = {}
        for name, v in (("warp", 2), ("auto", 0)):
            _lib.check(L.txp_debug_set(0, v))
            out = torch.zeros((w // 4) * (h // 4) * bs, dtype=torch.uint8, device="cuda")
            _lib.check(L.txp_compress_device(fmt, ctypes.c_void_p(d.data_ptr()), w, h, ctypes.byref(cp), ctypes.c_void_p(out.data_ptr()), out.numel(), None))
            torch.cuda.synchronize()
            outs[name] = out
        _lib.check(L.txp_debug_set(0, 0))
        assert torch.equal(outs["warp"], outs["auto"]), (fmt, alg)
