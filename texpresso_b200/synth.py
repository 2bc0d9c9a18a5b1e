"""Synthetic RGBA inputs of SURVEY.md section 8(d): counter-based (splitmix64), so any sub-rectangle can be
generated independently and host / device generators agree.  Used by tests and bench.py."""
import numpy as np

_M = (1 << 64) - 1


def splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _hash_rows(seed, w, y0, y1):
    with np.errstate(over="ignore"):
        y = np.arange(y0, y1, dtype=np.uint64)[:, None]
        x = np.arange(w, dtype=np.uint64)[None, :]
        ctr = np.uint64((int(seed) << 32) & _M) + y * np.uint64(w) + x
        return splitmix64(ctr)


def _tri(t, p):
    return (255 * np.abs((t % (2 * p)) - p)) // p


def generate(kind, width, height, seed, y0=0, y1=None, rows_per_chunk=512):
    """Returns a (y1-y0, width, 4) uint8 array.  kind: noise_alpha | noise_opaque | smooth | r_rg | r_rg_smooth.
    r_rg is uniform byte noise in R and G (every value 0..255 equally likely: 11.8 % of the 4x4 blocks contain a 0 or a 255,
    the values the 5-point alpha book treats specially, alpha.rs:194-212); r_rg_smooth is the "smooth variant" of
    SURVEY 8(d): two triangle-wave gradients plus 4-5 bits of noise, values 20..235 (height / normal map like)."""
    y1 = height if y1 is None else y1
    out = np.empty((y1 - y0, width, 4), dtype=np.uint8)
    for a in range(y0, y1, rows_per_chunk):
        b = min(a + rows_per_chunk, y1)
        h = _hash_rows(seed, width, a, b)
        o = out[a - y0:b - y0]
        if kind in ("noise_alpha", "noise_opaque"):
            for c in range(4):
                o[..., c] = ((h >> np.uint64(8 * c)) & np.uint64(255)).astype(np.uint8)
            if kind == "noise_opaque":
                o[..., 3] = 255
        elif kind == "r_rg":
            o[..., 0] = (h & np.uint64(255)).astype(np.uint8)
            o[..., 1] = ((h >> np.uint64(8)) & np.uint64(255)).astype(np.uint8)
            o[..., 2] = 0
            o[..., 3] = 255
        elif kind == "r_rg_smooth":
            y = np.arange(a, b, dtype=np.int64)[:, None]
            x = np.arange(width, dtype=np.int64)[None, :]
            o[..., 0] = (24 + (_tri(x + y // 2, 193) * 3) // 4 + ((h >> np.uint64(16)) & np.uint64(15)).astype(np.int64)).astype(np.uint8)
            o[..., 1] = (20 + (_tri(y + x // 3, 167) * 3) // 4 + ((h >> np.uint64(24)) & np.uint64(31)).astype(np.int64)).astype(np.uint8)
            o[..., 2] = 0
            o[..., 3] = 255
        elif kind == "smooth":
            y = np.arange(a, b, dtype=np.int64)[:, None]
            x = np.arange(width, dtype=np.int64)[None, :]
            for c in range(3):
                base = ((_tri(x + 51 * c, 97) + _tri(y + 29 * c, 61)) // 2) & ~7
                bump = (((h >> np.uint64(8 * c)) & np.uint64(3)) == 0).astype(np.int64) * 8
                o[..., c] = np.minimum(255, base + bump).astype(np.uint8)
            o[..., 3] = _tri(x + y, 128).astype(np.uint8)
            # every 8th block column is flat: copy the block's top-left pixel (needs block-aligned chunks)
            assert a % 4 == 0
            for bx in range(0, (width + 3) // 4, 8):
                xs = slice(4 * bx, min(4 * bx + 4, width))
                for by in range(0, b - a, 4):
                    o[by:by + 4, xs, :] = o[by, 4 * bx, :]
        else:
            raise ValueError(kind)
    return out
