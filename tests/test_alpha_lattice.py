"""CPU check of the BC4/BC5 closed-form fast path (texpresso_b200/csrc/txp_alpha_lattice.cuh).

The kernel's arithmetic for "regular" blocks (no 0 / 255, max - min >= 7) is restated in numpy from the generated
table (texpresso_b200/csrc/alpha_lattice_data.h) and compared with the C oracle (alpha.rs:187-256) block by block:
exhaustively over every range r = 7..255 and every offset inside the range, plus random regular blocks.  The GPU
tests (test_gpu_parity / test_gpu_fuzz / test_gpu_alpha_lattice) then check the kernel itself."""
import pathlib, re, struct, subprocess, sys
import numpy as np
import pytest

from tests import oracle_lib as O

ROOT = pathlib.Path(__file__).resolve().parent.parent
HDR = ROOT / "texpresso_b200" / "csrc" / "alpha_lattice_data.h"


def load_table():
    rows = re.findall(r"\{((?:0x[0-9A-F]{8}u(?:, )?){8})\}", HDR.read_text())
    assert len(rows) == 256
    tab = np.array([[int(x[:-1], 16) for x in r.split(", ")] for r in rows], dtype=np.uint32)
    return tab


def f32(bits):
    return np.frombuffer(np.asarray(bits, dtype=np.uint32).tobytes(), dtype=np.float32).astype(np.float64)


def emulate(values, tab):
    """values: (n, 16) uint8 regular blocks -> (n, 8) uint8 BC4 blocks, step by step as the kernel does."""
    v = values.astype(np.int64)
    lo, hi = v.min(axis=1), v.max(axis=1)
    r = hi - lo
    assert (lo > 0).all() and (hi < 255).all() and (r >= 7).all()
    row = tab[r]
    a5, b5, a7, b7 = f32(row[:, 0]), f32(row[:, 1]), f32(row[:, 4]), f32(row[:, 5])
    offs5 = np.stack([(row[:, 2 + k // 4] >> (8 * (k % 4))) & 255 for k in range(8)], axis=1).astype(np.int64)
    offs7 = np.stack([(row[:, 6 + k // 4] >> (8 * (k % 4))) & 255 for k in range(8)], axis=1).astype(np.int64)
    # d = vm - origin (exact), slot = low mantissa bits of fma(d, a, 1.5 * 2^23) = rint(d * a)
    d5 = v - (lo[:, None] - b5[:, None])
    s5 = np.rint(d5 * a5[:, None]).astype(np.int64)
    d7 = v - (hi[:, None] + b7[:, None])
    s7 = np.rint(d7 * -a7[:, None]).astype(np.int64)
    assert s5.min() >= 0 and s5.max() <= 5 and s7.min() >= 0 and s7.max() <= 7
    c5 = lo[:, None] + np.take_along_axis(offs5, s5, axis=1)
    c7 = hi[:, None] - np.take_along_axis(offs7, s7, axis=1)
    e5 = ((c5 - v) ** 2).sum(axis=1)
    e7 = ((c7 - v) ** 2).sum(axis=1)
    five = e5 <= e7
    map5 = np.array([0, 2, 3, 4, 5, 1, 0, 0]); map7 = np.array([0, 2, 3, 4, 5, 6, 7, 1])
    idx = np.where(five[:, None], map5[s5], map7[s7])
    a0 = np.where(five, lo, hi); a1 = np.where(five, hi, lo)
    bits = np.zeros(len(v), dtype=np.uint64)
    for i in range(16):
        bits |= idx[:, i].astype(np.uint64) << np.uint64(3 * i)
    out = np.zeros((len(v), 8), dtype=np.uint8)
    out[:, 0] = a0; out[:, 1] = a1
    for k in range(6):
        out[:, 2 + k] = ((bits >> np.uint64(8 * k)) & np.uint64(255)).astype(np.uint8)
    return out


def oracle_bc4(values):
    n = len(values)
    blocks = np.zeros((n, 16, 4), dtype=np.uint8)
    blocks[:, :, 0] = values
    blocks[:, :, 3] = 255
    return O.compress_blocks(O.BC4, blocks, np.full(n, 0xFFFF, np.uint32))


def test_table_is_current(tmp_path):
    """the committed header is what tools/gen_alpha_lattice.py generates (its own exhaustive proof runs while generating)"""
    before = HDR.read_text()
    subprocess.run([sys.executable, str(ROOT / "tools" / "gen_alpha_lattice.py")], check=True, stdout=subprocess.DEVNULL)
    assert HDR.read_text() == before


def test_every_range_every_offset():
    tab = load_table()
    rng = np.random.default_rng(7)
    blocks = []
    for r in range(7, 254):
        for lo in {1, 254 - r, int(rng.integers(1, 255 - r))}:
            xs = np.arange(r + 1)
            # 14 free pixels per block next to the two that pin lo and hi; cover every offset, in shuffled positions
            for c in range(0, r + 1, 14):
                fill = np.resize(xs[c:c + 14], 14)
                vals = np.concatenate(([0, r], fill)) + lo
                blocks.append(rng.permutation(vals))
    values = np.array(blocks, dtype=np.uint8)
    got = emulate(values, tab)
    want = oracle_bc4(values)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, (bad.size, values[bad[0]].tolist(), bytes(got[bad[0]]).hex(), bytes(want[bad[0]]).hex())


@pytest.mark.parametrize("seed", [1, 2])
def test_random_regular_blocks(seed):
    tab = load_table()
    rng = np.random.default_rng(seed)
    n = 60000
    lo = rng.integers(1, 248, size=n)
    hi = np.minimum(254, lo + rng.integers(7, 254, size=n))
    values = (lo[:, None] + (rng.random((n, 16)) * (hi - lo + 1)[:, None]).astype(np.int64)).clip(1, 254)
    # a third: values clustered near the lattice points / mid points (ties)
    values[: n // 3] = (lo[: n // 3, None] + np.round(rng.integers(0, 15, size=(n // 3, 16)) * (hi - lo)[: n // 3, None] / 14.0)).clip(1, 254)
    values = values.astype(np.uint8)
    keep = (values.max(axis=1).astype(int) - values.min(axis=1)) >= 7
    values = values[keep]
    got = emulate(values, tab)
    want = oracle_bc4(values)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, (bad.size, values[bad[0]].tolist(), bytes(got[bad[0]]).hex(), bytes(want[bad[0]]).hex())


# ---- narrow-range blocks (max - min <= 6, no 0 / 255, min <= 248): closed form of txp_alpha_lattice.cuh -----------------
def emulate_narrow(values):
    """fix_range (alpha.rs:70-77) widens both ranges: 5-point book codes lo..lo+5 (or lo+{0,1,2,3,4,6} when r == 6), 7-point
    book every integer lo..lo+7, so err7 == 0 always and err5 == 0 unless r == 6 and a pixel sits at lo+5."""
    v = values.astype(np.int64)
    lo, hi = v.min(axis=1), v.max(axis=1)
    r = hi - lo
    assert (lo >= 1).all() and (hi <= 254).all() and (r <= 6).all() and (lo <= 248).all()
    x = v - lo[:, None]
    seven = (r == 6) & (x == 5).any(axis=1)                       # the only case with err5 > err7 (alpha.rs:251)
    map5 = np.array([0, 2, 3, 4, 5, 1, 1, 0])                     # x -> index, block written (lo, hi5)
    map7 = np.array([1, 7, 6, 5, 4, 3, 0, 0])                     # x -> index, block written swapped (lo+7, lo)
    idx = np.where(seven[:, None], map7[x], map5[x])
    a0 = np.where(seven, lo + 7, lo)
    a1 = np.where(seven, lo, np.where(r < 5, lo + 5, hi))
    bits = np.zeros(len(v), dtype=np.uint64)
    for i in range(16):
        bits |= idx[:, i].astype(np.uint64) << np.uint64(3 * i)
    out = np.zeros((len(v), 8), dtype=np.uint8)
    out[:, 0] = a0; out[:, 1] = a1
    for k in range(6):
        out[:, 2 + k] = ((bits >> np.uint64(8 * k)) & np.uint64(255)).astype(np.uint8)
    return out


def test_narrow_blocks():
    rng = np.random.default_rng(5)
    blocks = []
    for lo in range(1, 249):
        for r in range(0, 7):
            if lo + r > 254:
                continue
            for rep in range(6):
                vals = lo + rng.integers(0, r + 1, size=16)
                vals[rng.integers(0, 16)] = lo; vals[rng.integers(0, 16)] = lo + r
                if (vals.max() - vals.min()) != r:
                    vals[0] = lo; vals[1] = lo + r
                if rep == 0 and r == 6:
                    vals[2:] = np.where(vals[2:] == lo + 5, lo + 4, vals[2:])      # r == 6 without a pixel at lo+5
                blocks.append(vals)
    values = np.array(blocks, dtype=np.uint8)
    got = emulate_narrow(values)
    want = oracle_bc4(values)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, (bad.size, values[bad[0]].tolist(), bytes(got[bad[0]]).hex(), bytes(want[bad[0]]).hex())


def test_flat_zero_and_255_blocks():
    """all-0 and all-255 blocks have constant encodings (used as literals by the kernel)"""
    z = oracle_bc4(np.zeros((1, 16), np.uint8))[0]
    f = oracle_bc4(np.full((1, 16), 255, np.uint8))[0]
    assert bytes(z).hex() == "0005000000000000"
    assert bytes(f).hex() == "0005ffffffffffff"


# ---- "one-sided" blocks: zeros (or 255s) next to ordinary values -- semi-lattice path of txp_alpha_lattice.cuh -----------
def semi_class(values):
    """0 = not eligible, 1 = zero side (zeros present, no 255), 2 = 255 side.  Mirrors alpha_semi_class()."""
    v = values.astype(np.int64)
    z = (v == 0).any(axis=1); f = (v == 255).any(axis=1)
    inner = (v != 0) & (v != 255)
    has = inner.any(axis=1)
    m = np.where(inner, v, 999).min(axis=1); M = np.where(inner, v, -1).max(axis=1)
    ok = has & (z ^ f) & (M - m >= 7)
    c1 = M // 7                                              # first interpolant of the 7-point lattice (0, M)
    c6 = m + (6 * (255 - m)) // 7                            # last interpolant of the lattice (m, 255)
    okz = ok & z & (m <= c1)
    okf = ok & f & (M >= c6)
    return np.where(okz, 1, np.where(okf, 2, 0)), m, M


def emulate_semi(values, tab, side):
    v = values.astype(np.int64)
    inner = (v != 0) & (v != 255)
    m = np.where(inner, v, 999).min(axis=1); M = np.where(inner, v, -1).max(axis=1)
    r5 = M - m
    lo7 = np.zeros_like(m) if side == 1 else m
    hi7 = M if side == 1 else np.full_like(M, 255)
    r7 = hi7 - lo7
    row5, row7 = tab[r5], tab[r7]
    a5, b5 = f32(row5[:, 0]), f32(row5[:, 1])
    a7, b7 = f32(row7[:, 4]), f32(row7[:, 5])
    offs5 = np.stack([(row5[:, 2 + k // 4] >> (8 * (k % 4))) & 255 for k in range(8)], axis=1).astype(np.int64)
    offs7 = np.stack([(row7[:, 6 + k // 4] >> (8 * (k % 4))) & 255 for k in range(8)], axis=1).astype(np.int64)
    # book 5: lattice (m, r5); the special pixels are sent to the slot-6 / slot-7 centres
    x6 = (m - b5) + 6.0 / a5; x7 = (m - b5) + 7.0 / a5
    v5 = np.where(v == 0, x6[:, None], np.where(v == 255, x7[:, None], v.astype(np.float64)))
    s5 = np.rint((v5 - (m - b5)[:, None]) * a5[:, None]).astype(np.int64)
    lut5 = m[:, None] + offs5
    lut5[:, 6] = 0; lut5[:, 7] = 255
    c5 = np.take_along_axis(lut5, s5, axis=1)
    e5 = ((c5 - v) ** 2).sum(axis=1)
    # book 7: lattice (lo7, r7) from the hi end; E0 = m / E1 = M replace the end codes, pixels they win are moved to the ends
    lut7 = hi7[:, None] - offs7
    f = lo7[:, None] + (np.arange(8)[None, :] * r7[:, None]) // 7            # interpolants c_i = lo7 + floor(i r7 / 7)
    if side == 1:
        c1, c2 = f[:, 1], f[:, 2]
        cup = np.where(c1 > m, c1, c2)
        B0 = (m + cup) // 2
        v7 = np.where(v <= B0[:, None], 0, v)
        lut7[:, 7] = m
    else:
        c6, c5i = f[:, 6], f[:, 5]
        cdn = np.where(c6 < M, c6, c5i)
        A1 = -((-(M + cdn)) // 2)
        v7 = np.where(v >= A1[:, None], 255, v)
        lut7[:, 0] = M
    s7 = np.rint((v7 - (hi7 + b7)[:, None]) * -a7[:, None]).astype(np.int64)
    assert s5.min() >= 0 and s5.max() <= 7 and s7.min() >= 0 and s7.max() <= 7
    c7 = np.take_along_axis(lut7, s7, axis=1)
    e7 = ((c7 - v) ** 2).sum(axis=1)
    five = e5 <= e7
    map5 = np.array([0, 2, 3, 4, 5, 1, 6, 7]); map7 = np.array([0, 2, 3, 4, 5, 6, 7, 1])
    idx = np.where(five[:, None], map5[s5], map7[s7])
    a0 = np.where(five, m, hi7); a1 = np.where(five, M, lo7)
    bits = np.zeros(len(v), dtype=np.uint64)
    for i in range(16):
        bits |= idx[:, i].astype(np.uint64) << np.uint64(3 * i)
    out = np.zeros((len(v), 8), dtype=np.uint8)
    out[:, 0] = a0; out[:, 1] = a1
    for k in range(6):
        out[:, 2 + k] = ((bits >> np.uint64(8 * k)) & np.uint64(255)).astype(np.uint8)
    return out


def _one_sided_corpus(seed, n):
    rng = np.random.default_rng(seed)
    lo = rng.integers(1, 248, size=n)
    hi = np.minimum(254, lo + rng.integers(7, 254, size=n))
    vals = (lo[:, None] + (rng.random((n, 16)) * (hi - lo + 1)[:, None]).astype(np.int64)).clip(1, 254)
    third = n // 3
    vals[:third] = rng.integers(1, 255, size=(third, 16))                      # plain noise
    vals[third:2 * third] = (lo[third:2 * third, None] + np.round(rng.integers(0, 15, size=(third, 16)) * (hi - lo)[third:2 * third, None] / 14.0)).clip(1, 254)
    special = np.where(rng.random(n) < 0.5, 0, 255)
    k = rng.integers(1, 5, size=n)
    for j in range(4):
        pos = rng.integers(0, 16, size=n)
        sel = k > j
        vals[np.nonzero(sel)[0], pos[sel]] = special[sel]
    return vals.astype(np.uint8)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_one_sided_blocks(seed):
    tab = load_table()
    values = _one_sided_corpus(seed, 90000)
    cls, m, M = semi_class(values)
    assert (cls == 1).sum() > 5000 and (cls == 2).sum() > 5000
    for side in (1, 2):
        sub = values[cls == side]
        got = emulate_semi(sub, tab, side)
        want = oracle_bc4(sub)
        bad = np.nonzero((got != want).any(axis=1))[0]
        assert bad.size == 0, (side, bad.size, len(sub), sub[bad[0]].tolist(), bytes(got[bad[0]]).hex(), bytes(want[bad[0]]).hex())


def _swept_one_sided():
    rng = np.random.default_rng(3)
    blocks = []
    for M in range(8, 255, 3):
        for m in range(1, M // 7 + 1):
            if M - m < 7:
                continue
            xs = np.arange(m, M + 1)
            for c in range(0, len(xs), 13 * 4):                 # a quarter of the chunks: keeps the CPU suite short
                fill = np.resize(xs[c:c + 13], 13)
                blocks.append(rng.permutation(np.concatenate(([0, m, M], fill))))
    for m in range(1, 247, 2):
        c6 = m + (6 * (255 - m)) // 7
        for M in range(max(c6, m + 7), 255):
            xs = np.arange(m, M + 1)
            c = int(rng.integers(0, max(1, len(xs) - 13)))
            fill = np.resize(xs[c:c + 13], 13)
            blocks.append(rng.permutation(np.concatenate(([255, m, M], fill))))
    return np.array(blocks, dtype=np.uint8)


def test_one_sided_blocks_swept():
    """every eligible (m, M) pair of the zero side (stride 3 in M) and of the 255 side, the 13 free pixels sweeping [m, M]"""
    tab = load_table()
    values = _swept_one_sided()
    cls, _, _ = semi_class(values)
    assert (cls > 0).all()
    for side in (1, 2):
        sub = values[cls == side]
        got = emulate_semi(sub, tab, side)
        want = oracle_bc4(sub)
        bad = np.nonzero((got != want).any(axis=1))[0]
        assert bad.size == 0, (side, bad.size, len(sub), sub[bad[0]].tolist(), bytes(got[bad[0]]).hex(), bytes(want[bad[0]]).hex())
