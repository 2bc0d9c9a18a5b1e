"""Exact fp32 identities the CUDA kernels rely on, checked exhaustively in numpy float32
(numpy float32 arithmetic is IEEE single, no contraction)."""
import numpy as np

f = np.float32


def roundf(x):                      # libm roundf: half away from zero
    return np.sign(x) * np.floor(np.abs(x) + f(0.5))


def test_pack565_of_snapped_endpoint_is_the_grid_index():
    # colourblock.rs:28-34 applied to k * fl(1/grid)  (cluster.rs:209-210, range.rs:97-98)
    for grid in (31, 63):
        k = np.arange(grid + 1, dtype=np.float32)
        v = k * (f(1.0) / f(grid))
        assert np.array_equal(np.clip(roundf(f(grid) * v), 0, grid), k)


def test_pack565_of_single_colour_lut_endpoint_is_the_lut_value():
    # single.rs:92-101: f32::from(s) / 31.0 then pack_565
    for grid in (31, 63):
        s = np.arange(grid + 1, dtype=np.float32)
        assert np.array_equal(np.clip(roundf(f(grid) * (s / f(grid))), 0, grid), s)


def test_single_colour_byte_roundtrip():
    # single.rs:60-64 on colourset.rs:65-67: round((c/255)*255) == c
    c = np.arange(256, dtype=np.float32)
    assert np.array_equal(np.clip(roundf((c / f(255.0)) * f(255.0)), 0, 255), c)


def test_weight_sums_are_exact_in_any_order():
    # colourset.rs:70-97: sums of (alpha+1)/256 over up to 16 pixels are multiples of 2^-8 below 2^5 -> exact
    rng = np.random.default_rng(0)
    for _ in range(2000):
        a = rng.integers(0, 256, size=16)
        w = (a + 1).astype(np.float32) / f(256.0)
        seq = f(0.0)
        for x in w:
            seq = f(seq + x)
        assert seq == f((a + 1).sum()) / f(256.0)


def test_alpha_key_argmin_matches_first_min_of_squared_distance():
    # alpha.rs:101-111 vs key_j = v*(-16*c_j) + (8*c_j^2 + j)  (txp_alpha.cuh fast path)
    rng = np.random.default_rng(1)
    for _ in range(3000):
        codes = rng.integers(0, 256, size=8)
        v = rng.integers(0, 256)
        d2 = (v - codes) ** 2
        ref = int(np.argmin(d2))                         # numpy argmin returns the first minimum
        key = v * (-16 * codes) + 8 * codes * codes + np.arange(8)
        assert int(np.argmin(key)) == ref and int(key.min()) & 7 == ref
        assert (int(key.min()) - ref) // 8 + v * v == int(d2.min())


def test_swap7_field_trick():
    # write_alpha_block7 (alpha.rs:172-178): 0->1, 1->0, x->9-x  ==  (9 - f) & 7
    for x in range(8):
        want = 1 if x == 0 else 0 if x == 1 else 9 - x
        assert (9 - x) & 7 == want


def test_division_free_alpha_interpolants():
    # txp_alpha.cuh alpha_codebooks_fast: ((N-i)*lo + i*hi)/N == lo + floor(i*(hi-lo)/N) with
    # floor(x/5) == (x*205)>>10 for x <= 1020 and floor(x/7) == (x*9363)>>16 for x <= 1530
    assert all((x * 205) >> 10 == x // 5 for x in range(1021))
    assert all((x * 9363) >> 16 == x // 7 for x in range(1531))
    for lo in range(0, 256, 5):
        for hi in range(lo, 256, 3):
            for i in range(1, 5):
                assert ((5 - i) * lo + i * hi) // 5 == lo + ((i * 205 * (hi - lo)) >> 10)
            for i in range(1, 7):
                assert ((7 - i) * lo + i * hi) // 7 == lo + ((i * 9363 * (hi - lo)) >> 16)


def _byte_perm(a, b, s):
    """CUDA __byte_perm / PRMT (default mode)."""
    src = [(a >> (8 * i)) & 255 for i in range(4)] + [(b >> (8 * i)) & 255 for i in range(4)]
    out = 0
    for i in range(4):
        n = (s >> (4 * i)) & 0xF
        v = src[n & 7]
        if n & 8:
            v = 0xFF if v & 0x80 else 0
        out |= v << (8 * i)
    return out


def test_decoder_prmt_selectors():
    # txp_decode.cuh: index spreaders and the plane -> pixel transpose selectors
    def spread2to4(x):
        y = (x | (x << 4)) & 0x0F0F
        return (y | (y << 2)) & 0x3333

    def spread3to4(x):
        y = (x & 0x3F) | ((x & 0xFC0) << 2)
        return (y & 0x0707) | ((y & 0x3838) << 1)

    for x in range(256):
        assert [(spread2to4(x) >> (4 * i)) & 0xF for i in range(4)] == [(x >> (2 * i)) & 3 for i in range(4)]
    for x in range(4096):
        assert [(spread3to4(x) >> (4 * i)) & 0xF for i in range(4)] == [(x >> (3 * i)) & 7 for i in range(4)]
    rng = np.random.default_rng(3)
    for _ in range(500):
        planes = [int(v) for v in rng.integers(0, 1 << 32, 4)]
        x = int(rng.integers(0, 256))
        sel = spread2to4(x)
        r4, g4, b4, a4 = [_byte_perm(p, 0, sel) for p in planes]
        rg_lo, rg_hi = _byte_perm(r4, g4, 0x5140), _byte_perm(r4, g4, 0x7362)
        ba_lo, ba_hi = _byte_perm(b4, a4, 0x5140), _byte_perm(b4, a4, 0x7362)
        px = [_byte_perm(rg_lo, ba_lo, 0x5410), _byte_perm(rg_lo, ba_lo, 0x7632), _byte_perm(rg_hi, ba_hi, 0x5410), _byte_perm(rg_hi, ba_hi, 0x7632)]
        for k in range(4):
            idx = (x >> (2 * k)) & 3
            assert px[k] == sum(((planes[c] >> (8 * idx)) & 255) << (8 * c) for c in range(4))
