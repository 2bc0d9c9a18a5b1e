"""Randomised differential test: many random blocks (random RGBA, random masks, clustered palettes) through the CUDA
path and the oracle.  Bit-exact for every format and algorithm (the ClusterFit bar of 99.9 % is reported, and met at 100 %
on every corpus so far; a mismatch prints the block so it can become a regression case)."""
import numpy as np
import pytest

from tests import oracle_lib as O

pytestmark = pytest.mark.gpu
N = 12000


def _corpus(seed):
    rng = np.random.default_rng(seed)
    blocks = rng.integers(0, 256, size=(N, 16, 4), dtype=np.uint8)
    # a third: few-colour palettes with small perturbations (duplicates, ties, near-degenerate axes)
    k = N // 3
    pal = rng.integers(0, 256, size=(k, 4, 3), dtype=np.uint8)
    pick = rng.integers(0, 4, size=(k, 16))
    blocks[:k, :, :3] = np.take_along_axis(pal, pick[..., None].repeat(3, axis=2), axis=1)
    blocks[: k // 2, :, :3] &= 0xF8                         # quantised -> exact ties on the 5:6:5 grid
    # a sixth: greys and single-channel ramps
    g = rng.integers(0, 256, size=(N // 6, 16), dtype=np.uint8)
    blocks[k:k + N // 6, :, 0] = g; blocks[k:k + N // 6, :, 1] = g; blocks[k:k + N // 6, :, 2] = g
    # alpha: mostly opaque-ish / binary / random mix
    mode = rng.integers(0, 3, size=N)
    blocks[mode == 0, :, 3] = 255
    blocks[mode == 1, :, 3] = np.where(rng.integers(0, 2, size=(int((mode == 1).sum()), 16)) == 1, 255, 0)
    masks = np.full(N, 0xFFFF, np.uint32)
    part = rng.random(N) < 0.12
    masks[part] = rng.integers(0, 1 << 16, size=int(part.sum()))
    return blocks, masks


@pytest.mark.parametrize("fmt,alg,weights,awa", [
    (0, 0, O.PERCEPTUAL, False), (0, 1, O.PERCEPTUAL, False), (0, 2, O.PERCEPTUAL, False), (0, 1, O.UNIFORM, True),
    (1, 1, O.PERCEPTUAL, False), (1, 2, O.UNIFORM, False),
    (2, 0, O.UNIFORM, True), (2, 1, O.PERCEPTUAL, False), (2, 2, O.PERCEPTUAL, True),
    (3, 1, O.PERCEPTUAL, False), (4, 1, O.PERCEPTUAL, False),
])
def test_random_blocks_bit_exact(fmt, alg, weights, awa):
    import texpresso_b200 as T
    blocks, masks = _corpus(1000 + 17 * fmt + alg)
    tp = T.Params(T.Algorithm(alg), tuple(weights), awa)
    got = T.compress_blocks(fmt, blocks, masks, tp)
    want = O.compress_blocks(fmt, blocks, masks, O.make_params(alg, weights, awa))
    diff = np.nonzero((got != want).any(axis=1))[0]
    assert diff.size == 0, (f"{diff.size}/{N} blocks differ",
                            [(int(i), hex(int(masks[i])), bytes(blocks[i].reshape(-1)).hex(), bytes(got[i]).hex(), bytes(want[i]).hex()) for i in diff[:3]])
    dec = T.decompress_blocks(fmt, want)
    assert np.array_equal(dec, O.decompress_blocks(fmt, want))
