"""GPU parity of the BC4/BC5 lattice kernel (txp_alpha_lattice.cuh) against the oracle: the exhaustive per-range corpus
of test_alpha_lattice.py, mixed regular / irregular blocks (so the per-warp queue and its drain are exercised with
every fill level), partial masks, and whole images through the image-mode entry point."""
import numpy as np
import pytest

from tests import oracle_lib as O

pytestmark = pytest.mark.gpu


def _exhaustive_values():
    rng = np.random.default_rng(7)
    blocks = []
    for r in range(7, 254):
        for lo in {1, 254 - r, int(rng.integers(1, 255 - r))}:
            xs = np.arange(r + 1)
            for c in range(0, r + 1, 14):
                fill = np.resize(xs[c:c + 14], 14)
                blocks.append(rng.permutation(np.concatenate(([0, r], fill)) + lo))
    return np.array(blocks, dtype=np.uint8)


def _mixed_blocks(n, seed, irregular_fraction):
    """RGBA blocks whose R and G channels are independently regular or irregular"""
    rng = np.random.default_rng(seed)
    blocks = np.zeros((n, 16, 4), dtype=np.uint8)
    for ch in (0, 1):
        lo = rng.integers(1, 240, size=n)
        hi = np.minimum(254, lo + rng.integers(7, 254, size=n))
        v = lo[:, None] + (rng.random((n, 16)) * (hi - lo + 1)[:, None]).astype(np.int64)
        v = v.clip(1, 254)
        irr = rng.random(n) < irregular_fraction
        kind = rng.integers(0, 4, size=n)
        z = irr & (kind == 0); v[z, rng.integers(0, 16)] = 0                      # a zero present
        f = irr & (kind == 1); v[f, rng.integers(0, 16)] = 255                    # a 255 present
        nrw = irr & (kind == 2); v[nrw] = lo[nrw, None] + rng.integers(0, 6, size=(int(nrw.sum()), 16))   # range < 7
        both = irr & (kind == 3); v[both, 0] = 0; v[both, 5] = 255
        blocks[:, :, ch] = v.clip(0, 255)
    blocks[:, :, 2] = rng.integers(0, 256, size=(n, 16))
    blocks[:, :, 3] = 255
    return blocks


@pytest.mark.parametrize("fmt", [O.BC4, O.BC5])
def test_exhaustive_ranges(fmt):
    import texpresso_b200 as T
    values = _exhaustive_values()
    n = len(values)
    blocks = np.zeros((n, 16, 4), dtype=np.uint8)
    blocks[:, :, 0] = values
    blocks[:, :, 1] = values[::-1]
    masks = np.full(n, 0xFFFF, np.uint32)
    got = T.compress_blocks(fmt, blocks, masks, T.Params())
    want = O.compress_blocks(fmt, blocks, masks)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, (bad.size, blocks[bad[0], :, :2].tolist(), bytes(got[bad[0]]).hex(), bytes(want[bad[0]]).hex())


@pytest.mark.parametrize("fmt", [O.BC4, O.BC5])
@pytest.mark.parametrize("frac", [0.0, 0.02, 0.12, 0.5, 1.0])
def test_mixed_regular_irregular(fmt, frac):
    import texpresso_b200 as T
    n = 40000 + 13                                            # not a multiple of 32: partial last tile
    blocks = _mixed_blocks(n, 11 + int(frac * 100), frac)
    masks = np.full(n, 0xFFFF, np.uint32)
    rng = np.random.default_rng(3)
    part = rng.random(n) < 0.03
    masks[part] = rng.integers(0, 1 << 16, size=int(part.sum()))
    got = T.compress_blocks(fmt, blocks, masks, T.Params())
    want = O.compress_blocks(fmt, blocks, masks)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, (bad.size, hex(int(masks[bad[0]])), blocks[bad[0], :, :2].tolist(), bytes(got[bad[0]]).hex(), bytes(want[bad[0]]).hex())


@pytest.mark.parametrize("fmt", [O.BC4, O.BC5])
@pytest.mark.parametrize("size", [(1024, 512), (513, 258), (36, 4), (4, 4), (7, 9)])
@pytest.mark.parametrize("kind", ["r_rg", "smooth"])
def test_images(fmt, size, kind):
    import texpresso_b200 as T
    from texpresso_b200 import synth
    w, h = size
    img = synth.generate(kind, w, (h + 3) // 4 * 4, seed=77)[:h]
    img = np.ascontiguousarray(img)
    got = T.Format(fmt).compress(img, w, h, T.Params())
    want = O.compress(fmt, img, w, h)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("fmt", [O.BC4, O.BC5])
@pytest.mark.parametrize("size", [(128, 4), (128, 7), (256, 130), (1152, 515), (2048, 64), (4096, 1030)])
@pytest.mark.parametrize("kind", ["r_rg", "smooth"])
def test_images_tma_staged(fmt, size, kind):
    """widths that are a multiple of 128 pixels take alpha_lattice_tma_kernel (one cp.async.bulk.tensor per 32-block strip):
    ragged heights (partial bottom block row through the literal path), more tiles than resident warps, and an output that is
    longer than needed (the extra block rows are encoded fully masked)."""
    import texpresso_b200 as T
    from texpresso_b200 import synth
    w, h = size
    img = np.ascontiguousarray(synth.generate(kind, w, (h + 3) // 4 * 4, seed=78)[:h])
    got = T.Format(fmt).compress(img, w, h, T.Params())
    assert np.array_equal(got, O.compress(fmt, img, w, h, threads=8))
    bs = 8 if fmt == O.BC4 else 16
    n = got.size + (w // 4) * bs + 5 * bs                      # one extra block row + a partial row
    out = np.full(n, 0xEE, np.uint8)
    T.Format(fmt).compress(img, w, h, T.Params(), output=out)
    assert np.array_equal(out, O.compress(fmt, img, w, h, out_len=n))


@pytest.mark.parametrize("fmt", [O.BC4, O.BC5])
def test_narrow_and_flat_blocks(fmt):
    """closed-form path for flat / narrow-range blocks (alpha_fit_narrow): every lo, every range 0..6, plus flat 0 / 255,
    ranges clamped at 255 (lo > 248) and narrow ranges that touch 0 or 255 (those stay on the literal path)"""
    import texpresso_b200 as T
    rng = np.random.default_rng(5)
    vals = []
    for lo in range(0, 256):
        for r in range(0, 7):
            if lo + r > 255:
                continue
            for rep in range(3):
                v = lo + rng.integers(0, r + 1, size=16)
                v[rng.integers(0, 16)] = lo; v[rng.integers(0, 16)] = lo + r
                if rep == 0 and r == 6:
                    v = np.where(v == lo + 5, lo + 4, v)
                vals.append(v)
    values = np.array(vals, dtype=np.uint8)
    n = len(values)
    blocks = np.zeros((n, 16, 4), dtype=np.uint8)
    blocks[:, :, 0] = values
    blocks[:, :, 1] = values[rng.permutation(n)]
    masks = np.full(n, 0xFFFF, np.uint32)
    got = T.compress_blocks(fmt, blocks, masks, T.Params())
    want = O.compress_blocks(fmt, blocks, masks)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, (bad.size, blocks[bad[0], :, :2].tolist(), bytes(got[bad[0]]).hex(), bytes(want[bad[0]]).hex())


@pytest.mark.parametrize("fmt", [O.BC4, O.BC5])
def test_one_sided_blocks(fmt):
    """semi-lattice path (zeros or 255s next to ordinary values): the corpora of tests/test_alpha_lattice.py on the GPU,
    including blocks that fail its side conditions and fall through to the literal search"""
    import texpresso_b200 as T
    from tests.test_alpha_lattice import _one_sided_corpus, _swept_one_sided
    values = np.concatenate([_one_sided_corpus(21, 60000), _swept_one_sided()])
    n = len(values)
    rng = np.random.default_rng(9)
    blocks = np.zeros((n, 16, 4), dtype=np.uint8)
    blocks[:, :, 0] = values
    blocks[:, :, 1] = values[rng.permutation(n)]
    masks = np.full(n, 0xFFFF, np.uint32)
    got = T.compress_blocks(fmt, blocks, masks, T.Params())
    want = O.compress_blocks(fmt, blocks, masks)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, (bad.size, blocks[bad[0], :, :2].tolist(), bytes(got[bad[0]]).hex(), bytes(want[bad[0]]).hex())
