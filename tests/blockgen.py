"""Stratified 4x4 block generators for the differential tests: every dispatch
class and edge case of the reference's block path, seeded."""
import numpy as np


def _palette_blocks(rng, n, ncol, alpha_mode):
    """n blocks drawn from `ncol` random colours each."""
    pal = rng.integers(0, 256, size=(n, ncol, 3), dtype=np.uint8)
    pick = rng.integers(0, ncol, size=(n, 16))
    # make sure every palette entry is used when possible, so the distinct count is exactly ncol
    for c in range(min(ncol, 16)):
        pick[:, c] = c
    rgb = np.take_along_axis(pal, pick[..., None].repeat(3, axis=2), axis=1)
    out = np.empty((n, 16, 4), np.uint8)
    out[..., :3] = rgb
    if alpha_mode == "opaque":
        out[..., 3] = 255
    elif alpha_mode == "random":
        out[..., 3] = rng.integers(0, 256, size=(n, 16), dtype=np.uint8)
    elif alpha_mode == "binary":
        out[..., 3] = rng.choice(np.array([0, 255], np.uint8), size=(n, 16))
    elif alpha_mode == "around128":
        out[..., 3] = rng.integers(120, 136, size=(n, 16), dtype=np.uint8)
    return out


def colour_cases(seed=1234, per_class=48):
    """(blocks (N,16,4) uint8, masks (N,) uint32, tags list)"""
    rng = np.random.default_rng(seed)
    blocks, masks, tags = [], [], []

    def add(b, m, tag):
        blocks.append(b); masks.append(np.broadcast_to(np.asarray(m, np.uint32), (len(b),)).copy()); tags.extend([tag] * len(b))

    for ncol in range(1, 17):
        for amode in ("opaque", "random"):
            add(_palette_blocks(rng, per_class, ncol, amode), 0xFFFF, f"pal{ncol}_{amode}")
    add(_palette_blocks(rng, per_class, 16, "binary"), 0xFFFF, "binary_alpha")
    add(_palette_blocks(rng, per_class, 8, "around128"), 0xFFFF, "alpha_around_128")
    # smooth gradients (many exact ties on the 5:6:5 grid axes)
    g = np.empty((per_class, 16, 4), np.uint8)
    for t in range(per_class):
        o = rng.integers(0, 128, 3); dx = rng.integers(-8, 9, 3); dy = rng.integers(-8, 9, 3)
        for i in range(16):
            g[t, i, :3] = np.clip(o + dx * (i & 3) + dy * (i >> 2), 0, 255)
        g[t, :, 3] = rng.integers(0, 256)
    add(g, 0xFFFF, "gradient")
    # axis-aligned sets and grey ramps (tie constructions)
    a = np.zeros((per_class, 16, 4), np.uint8); a[..., 3] = 255
    for t in range(per_class):
        ch = t % 3
        a[t, :, ch] = rng.integers(0, 256, 16)
    add(a, 0xFFFF, "axis_aligned")
    gr = np.zeros((per_class, 16, 4), np.uint8); gr[..., 3] = 255
    v = rng.integers(0, 256, size=(per_class, 16), dtype=np.uint8)
    gr[..., 0] = v; gr[..., 1] = v; gr[..., 2] = v
    add(gr, 0xFFFF, "grey")
    # NaN-axis constructions: colour differences orthogonal to (1,1,1) with zero covariance row sums (Q7)
    nn = np.zeros((per_class, 16, 4), np.uint8); nn[..., 3] = 255
    for t in range(per_class):
        c0 = np.array([255, 0, 0]) if t % 2 == 0 else np.array([200, 10, 0])
        c1 = np.array([0, 255, 0]) if t % 2 == 0 else np.array([10, 200, 0])
        c2 = np.array([0, 0, 255])
        for i in range(16):
            nn[t, i, :3] = (c0, c1, c2)[(i + t // 2) % (2 + (t % 3 == 0))]
    add(nn, 0xFFFF, "nan_axis")
    # two-colour and near-duplicate sets
    add(_palette_blocks(rng, per_class, 2, "opaque"), 0xFFFF, "two_colour")
    nd = _palette_blocks(rng, per_class, 3, "opaque").astype(np.int16)
    nd[..., :3] += rng.integers(-1, 2, size=(per_class, 16, 3))
    add(np.clip(nd, 0, 255).astype(np.uint8), 0xFFFF, "near_duplicate")
    # extremes
    ex = _palette_blocks(rng, per_class, 4, "opaque"); ex[..., :3] = np.where(ex[..., :3] > 127, 255, 0)
    add(ex, 0xFFFF, "black_white_corners")
    # masks: single bit, 2x2, 4x1, 1x4, random, empty
    base = _palette_blocks(rng, 16, 16, "random")
    add(base, np.array([1 << i for i in range(16)], np.uint32), "mask_single_bit")
    add(_palette_blocks(rng, per_class, 16, "random"), 0x0033, "mask_2x2")
    add(_palette_blocks(rng, per_class, 16, "random"), 0x000F, "mask_4x1")
    add(_palette_blocks(rng, per_class, 16, "random"), 0x1111, "mask_1x4")
    add(_palette_blocks(rng, per_class, 16, "opaque"), 0x0777, "mask_3x3")
    rb = _palette_blocks(rng, 4 * per_class, 12, "random")
    add(rb, rng.integers(0, 1 << 16, size=len(rb)).astype(np.uint32), "mask_random")
    add(_palette_blocks(rng, 8, 16, "random"), 0, "mask_empty")
    z = _palette_blocks(rng, 8, 16, "opaque"); z[..., 3] = 0
    add(z, 0xFFFF, "all_alpha_zero")
    return np.concatenate(blocks), np.concatenate(masks), tags


def alpha_cases(seed=99, per_class=64):
    """Blocks whose R, G and A channels exercise the 5-/7-point alpha fit (alpha.rs:187-256)."""
    rng = np.random.default_rng(seed)
    blocks, masks, tags = [], [], []

    def add(vals, m, tag):
        n = len(vals)
        b = np.zeros((n, 16, 4), np.uint8)
        b[..., 0] = vals
        b[..., 1] = vals[:, ::-1]                       # a different arrangement in G
        b[..., 2] = rng.integers(0, 256, size=(n, 16))
        b[..., 3] = np.roll(vals, 5, axis=1)
        blocks.append(b); masks.append(np.broadcast_to(np.asarray(m, np.uint32), (n,)).copy()); tags.extend([tag] * n)

    add(rng.integers(0, 256, size=(per_class * 4, 16), dtype=np.uint8), 0xFFFF, "uniform_random")
    add(rng.choice(np.array([0, 255], np.uint8), size=(per_class, 16)), 0xFFFF, "only_0_255")
    for lo, hi in ((0, 4), (0, 8), (250, 256), (247, 256), (100, 103), (100, 108), (1, 255), (0, 255), (3, 130)):
        add(rng.integers(lo, hi, size=(per_class, 16), dtype=np.uint8), 0xFFFF, f"range_{lo}_{hi}")
    # values in a mid range plus a few 0 / 255 outliers: min5/max5 differ from min7/max7 (quirk Q1)
    v = rng.integers(40, 200, size=(per_class * 2, 16), dtype=np.uint8)
    v[:, 0] = 0; v[:per_class, 1] = 255
    add(v, 0xFFFF, "outliers_0_255")
    # smooth ramps -> 7-point mode
    r = (np.arange(16)[None, :] * rng.integers(1, 9, size=(per_class, 1)) + rng.integers(0, 120, size=(per_class, 1))).astype(np.uint8)
    add(r, 0xFFFF, "ramps")
    add(np.repeat(rng.integers(0, 256, size=(per_class, 1), dtype=np.uint8), 16, axis=1), 0xFFFF, "constant")
    rb = rng.integers(0, 256, size=(per_class * 2, 16), dtype=np.uint8)
    add(rb, rng.integers(0, 1 << 16, size=len(rb)).astype(np.uint32), "mask_random")
    add(rng.integers(0, 256, size=(16, 16), dtype=np.uint8), np.array([1 << i for i in range(16)], np.uint32), "mask_single_bit")
    add(rng.integers(0, 256, size=(4, 16), dtype=np.uint8), 0, "mask_empty")
    return np.concatenate(blocks), np.concatenate(masks), tags
