"""Second, independent restatement of the reference's block path in numpy -- TEST INFRASTRUCTURE.

Written from the Rust text (/root/reference/lib/src), NOT from oracle/txp_oracle.c, and with a different structure: the C oracle
walks one block at a time with scalar floats; this module walks the reference's loop nests once for ALL blocks of a batch, every
variable an array with one fp32 entry per block and every `if` a mask.  numpy's float32 +, -, *, /, sqrt are single IEEE
roundings with no contraction, which is the reference's arithmetic (scalar Rust f32, libm sqrtf / truncf / roundf).

Purpose ("parity unpinned"): the reference's own tests pin UNIFORM weights, <= 3 colours per block and 5-point alpha
only.  Two restatements that were written separately, in different languages and shapes, and agree bit for bit on 10^4 blocks per
configuration -- PERCEPTUAL weights, 16-colour blocks, 7-point alpha with its min5/max5 quirk, sort ties, NaN axes, iterations > 1,
punch-through, masks -- is the strongest pin available without a Rust toolchain.

Each function cites the reference lines it restates."""
import pathlib, re
import numpy as np

F = np.float32
BC1, BC2, BC3, BC4, BC5 = range(5)
RANGE_FIT, CLUSTER_FIT, ITERATIVE_CLUSTER_FIT = range(3)
F32_MAX = np.finfo(np.float32).max
F32_EPS = np.finfo(np.float32).eps


# ---- libm pieces ------------------------------------------------------------------------------------------------------------
def roundf(a):
    """libm::roundf: half away from zero (math.rs:101).  Exact: |a| + 0.5 is formed in float64."""
    a64 = a.astype(np.float64)
    return (np.sign(a64) * np.floor(np.abs(a64) + 0.5)).astype(F)


def f32_to_i32_clamped(a, limit):
    """math.rs:100-102: roundf(a).max(0).min(limit) as i32 (f32::max / min return the non-NaN operand; NaN as i32 = 0)"""
    r = roundf(a)
    r = np.fmin(np.fmax(r, F(0)), F(limit))
    return np.where(np.isnan(r), 0, r).astype(np.int64)


def rmax(a, b):  # f32::max
    return np.fmax(a, b)


def rmin(a, b):  # f32::min
    return np.fmin(a, b)


# ---- single colour tables (colourfit/single_lut.rs:52-1086, as re-laid-out data in oracle/single_lut_data.h) ------------------------------
_LUT = None


def single_lut():
    """-> uint8 array [table 0..3 = 5_3, 6_3, 5_4, 6_4][value 0..255][index 0..1][start, end, error]"""
    global _LUT
    if _LUT is None:
        text = (pathlib.Path(__file__).resolve().parent.parent / "oracle" / "single_lut_data.h").read_text()
        body = text[text.index("TXP_SINGLE_LUT_INIT {") + len("TXP_SINGLE_LUT_INIT {"):]
        vals = [int(x) for x in re.findall(r"\b\d+\b", body)]
        assert len(vals) >= 6144
        _LUT = np.array(vals[:6144], np.uint8).reshape(4, 256, 2, 3)
    return _LUT


# ---- alpha.rs ----------------------------------------------------------------------------------------------------------------
def compress_bc2(rgba, mask):
    """alpha.rs:27-51"""
    a = rgba[:, :, 3].astype(F) * F(15.0 / 255.0)
    q = f32_to_i32_clamped(a, 15)
    valid = ((mask[:, None] >> np.arange(16)[None, :]) & 1).astype(bool)
    q = np.where(valid, q, 0)
    return (q[:, 0::2] | (q[:, 1::2] << 4)).astype(np.uint8)


def _fix_range(mn, mx, steps):
    """alpha.rs:70-77"""
    mx = np.where(mx - mn < steps, np.minimum(mn + steps, 255), mx)
    mn = np.where(mx - mn < steps, np.maximum(mx - steps, 0), mn)
    return mn, mx


def _fit_codes(vals, valid, codes):
    """alpha.rs:79-119: first minimum of the squared distance over the 8 codes; masked pixels -> index 0, no error"""
    d = vals[:, :, None].astype(np.int64) - codes[:, None, :].astype(np.int64)
    d = d * d
    idx = np.argmin(d, axis=2)                            # argmin returns the FIRST minimum
    least = np.take_along_axis(d, idx[..., None], axis=2)[..., 0]
    idx = np.where(valid, idx, 0)
    err = np.where(valid, least, 0).sum(axis=1)
    return idx, err


def _write_alpha_block(a0, a1, idx):
    """alpha.rs:121-144"""
    n = len(a0)
    out = np.zeros((n, 8), np.uint8)
    out[:, 0] = a0; out[:, 1] = a1
    for g in range(2):
        value = np.zeros(n, np.int64)
        for j in range(8):
            value |= idx[:, 8 * g + j].astype(np.int64) << (3 * j)
        for j in range(3):
            out[:, 2 + 3 * g + j] = (value >> (8 * j)) & 0xFF
    return out


def compress_bc3(rgba, channel, mask):
    """alpha.rs:187-256 (+ write_alpha_block5 :146-165, write_alpha_block7 :167-185)"""
    vals = rgba[:, :, channel].astype(np.int64)
    valid = ((mask[:, None] >> np.arange(16)[None, :]) & 1).astype(bool)
    min7 = np.where(valid, vals, 255).min(axis=1)
    max7 = np.where(valid, vals, 0).max(axis=1)
    min5 = np.where(valid & (vals != 0), vals, 255).min(axis=1)
    max5 = np.where(valid & (vals != 255), vals, 0).max(axis=1)
    min5 = np.where(min5 > max5, max5, min5)              # :215-220
    min7 = np.where(min7 > max7, max7, min7)
    min5, max5 = _fix_range(min5, max5, 5)
    min7, max7 = _fix_range(min7, max7, 7)
    n = len(vals)
    codes5 = np.zeros((n, 8), np.int64)
    codes5[:, 0] = min5; codes5[:, 1] = max5
    for i in range(1, 5):
        codes5[:, 1 + i] = ((5 - i) * min5 + i * max5) // 5
    codes5[:, 6] = 0; codes5[:, 7] = 255
    codes7 = np.zeros((n, 8), np.int64)
    codes7[:, 0] = min5; codes7[:, 1] = max5             # :238-239 -- min5 / max5, not min7 / max7
    for i in range(1, 7):
        codes7[:, 1 + i] = ((7 - i) * min7 + i * max7) // 7
    idx5, err5 = _fit_codes(vals, valid, codes5)
    idx7, err7 = _fit_codes(vals, valid, codes7)
    # write_alpha_block5: swap if alpha0 > alpha1
    sw5 = min5 > max5
    m5 = np.array([1, 0, 5, 4, 3, 2, 6, 7])
    i5 = np.where(sw5[:, None], m5[idx5], idx5)
    b5 = _write_alpha_block(np.where(sw5, max5, min5), np.where(sw5, min5, max5), i5)
    # write_alpha_block7: swap if alpha0 < alpha1
    sw7 = min7 < max7
    m7 = np.array([1, 0, 7, 6, 5, 4, 3, 2])
    i7 = np.where(sw7[:, None], m7[idx7], idx7)
    b7 = _write_alpha_block(np.where(sw7, max7, min7), np.where(sw7, min7, max7), i7)
    return np.where((err5 <= err7)[:, None], b5, b7)


# ---- colourset.rs:35-112 -----------------------------------------------------------------------------------------------------------
class ColourSet:
    pass


def colour_set(rgba, mask, fmt, alpha_weighted):
    n = len(rgba)
    cs = ColourSet()
    cs.count = np.zeros(n, np.int64)
    cs.points = np.zeros((n, 16, 3), F)
    cs.weights = np.zeros((n, 16), F)
    cs.remap = np.zeros((n, 16), np.int64)
    cs.transparent = np.zeros(n, bool)
    rows = np.arange(n)
    for i in range(16):
        valid = ((mask >> i) & 1).astype(bool)
        cs.remap[~valid, i] = -1                                              # :47-51
        punched = valid & (fmt == BC1) & (rgba[:, i, 3] < 128)              # :54-58
        cs.remap[punched, i] = -1
        cs.transparent |= punched
        pending = valid & ~punched
        w = (rgba[:, i, 3].astype(np.int64) + 1).astype(F) / F(256)           # :70 / :94
        w = w if alpha_weighted else np.ones(n, F)
        for j in range(i + 1):
            if j == i:                                                        # :63-81: no duplicate found, new point
                r = rows[pending]
                c = cs.count[r]
                cs.points[r, c, 0] = rgba[r, i, 0].astype(F) / F(255)
                cs.points[r, c, 1] = rgba[r, i, 1].astype(F) / F(255)
                cs.points[r, c, 2] = rgba[r, i, 2].astype(F) / F(255)
                cs.weights[r, c] = w[r]
                cs.remap[r, i] = c
                cs.count[r] += 1
                break
            oldvalid = ((mask >> j) & 1).astype(bool)                         # :84-88
            dup = pending & oldvalid & (rgba[:, i, :3] == rgba[:, j, :3]).all(axis=1)
            if fmt == BC1:
                dup &= rgba[:, j, 3] >= 128
            r = rows[dup]
            index = cs.remap[r, j]
            cs.weights[r, index] = cs.weights[r, index] + w[r]                # :97
            cs.remap[r, i] = index
            pending = pending & ~dup
    cs.weights = np.sqrt(cs.weights)                                          # :107-109 (all 16 entries)
    return cs


def remap_indices(cs, source):
    """colourset.rs:130-141: source (n,16) per-point -> (n,16) per-pixel, 3 where remap == -1"""
    j = np.where(cs.remap < 0, 0, cs.remap)
    t = np.take_along_axis(source, j, axis=1)
    return np.where(cs.remap < 0, 3, t)


# ---- math.rs:44-97 -------------------------------------------------------------------------------------------------------------------
def weighted_covariance(cs):
    n = len(cs.count)
    total = np.zeros(n, F)
    cen = np.zeros((n, 3), F)
    for p in range(16):
        act = p < cs.count
        w = cs.weights[:, p]
        total = np.where(act, total + w, total)                               # weights.iter().sum()
        cen = np.where(act[:, None], cen + cs.points[:, p] * w[:, None], cen) # map(p * w).sum(), folded from 0
    big = total > F32_EPS
    with np.errstate(all="ignore"):
        cen = np.where(big[:, None], cen / total[:, None], cen)               # Div<f32> for Vec3: true division per lane
    cov = np.zeros((n, 6), F)
    for p in range(16):
        act = (p < cs.count)[:, None]
        a = cs.points[:, p] - cen
        b = a * cs.weights[:, p][:, None]
        add = np.stack([a[:, 0] * b[:, 0], a[:, 0] * b[:, 1], a[:, 0] * b[:, 2], a[:, 1] * b[:, 1], a[:, 1] * b[:, 2], a[:, 2] * b[:, 2]], axis=1)
        cov = np.where(act, cov + add, cov)
    return cov


def principle_component(cov):
    n = len(cov)
    z = np.zeros(n, F)
    row0 = np.stack([cov[:, 0], cov[:, 1], cov[:, 2], z], axis=1)
    row1 = np.stack([cov[:, 1], cov[:, 3], cov[:, 4], z], axis=1)
    row2 = np.stack([cov[:, 2], cov[:, 4], cov[:, 5], z], axis=1)
    v = np.ones((n, 4), F)
    with np.errstate(all="ignore"):
        for _ in range(8):
            w = row0 * v[:, 0:1]
            w = row1 * v[:, 1:2] + w
            w = row2 * v[:, 2:3] + w
            a = rmax(w[:, 0], rmax(w[:, 1], w[:, 2]))
            v = w * (F(1) / a)[:, None]                                      # a.reciprocal(), then multiply
    return v[:, :3]


# ---- colourblock.rs:28-94 ---------------------------------------------------------------------------------------------------------------
def pack_565(c):
    with np.errstate(all="ignore"):
        r = f32_to_i32_clamped(F(31) * c[:, 0], 31)
        g = f32_to_i32_clamped(F(63) * c[:, 1], 63)
        b = f32_to_i32_clamped(F(31) * c[:, 2], 31)
    return (r << 11) | (g << 5) | b


def write_block(a, b, idx):
    n = len(a)
    out = np.zeros((n, 8), np.uint8)
    out[:, 0] = a & 0xFF; out[:, 1] = a >> 8; out[:, 2] = b & 0xFF; out[:, 3] = b >> 8
    for i in range(4):
        out[:, 4 + i] = ((idx[:, 4 * i + 3] & 3) << 6) | ((idx[:, 4 * i + 2] & 3) << 4) | ((idx[:, 4 * i + 1] & 3) << 2) | (idx[:, 4 * i] & 3)
    return out


def write3(start, end, idx):
    a, b = pack_565(start), pack_565(end)
    sw = a > b
    m = np.array([1, 0, 2, 3])
    idx = np.where(sw[:, None], m[idx], idx)
    return write_block(np.where(sw, b, a), np.where(sw, a, b), idx)


def write4(start, end, idx):
    a, b = pack_565(start), pack_565(end)
    lt, gt = a < b, a > b
    out_idx = np.zeros_like(idx)                                              # a == b: index 0 everywhere
    out_idx = np.where(lt[:, None], (idx ^ 1) & 3, out_idx)
    out_idx = np.where(gt[:, None], idx, out_idx)
    return write_block(np.where(lt, b, a), np.where(lt, a, b), out_idx)


# ---- colourfit/single.rs:58-164 ----------------------------------------------------------------------------------------------------------
def single_colour_fit(cs, fmt):
    n = len(cs.count)
    lut = single_lut()
    p0 = cs.points[:, 0]
    colour = np.stack([f32_to_i32_clamped(p0[:, c] * F(255), 255) for c in range(3)], axis=1)
    best_error = np.full(n, 2**32 - 1, np.int64)
    best = np.zeros((n, 8), np.uint8)

    def endpoints(tabs):
        err_best = np.full(n, 2**32 - 1, np.int64)
        start = np.zeros((n, 3), F); end = np.zeros((n, 3), F); index = np.zeros(n, np.int64)
        for ix in range(2):
            src = [lut[tabs[c]][colour[:, c], ix] for c in range(3)]           # (n, 3): start, end, error
            e = sum(s[:, 2].astype(np.int64) ** 2 for s in src)
            better = e < err_best
            den = (F(31), F(63), F(31))
            s_ = np.stack([src[c][:, 0].astype(F) / den[c] for c in range(3)], axis=1)
            e_ = np.stack([src[c][:, 1].astype(F) / den[c] for c in range(3)], axis=1)
            start = np.where(better[:, None], s_, start); end = np.where(better[:, None], e_, end)
            index = np.where(better, 2 * ix, index)
            err_best = np.where(better, e, err_best)
        return start, end, index, err_best

    def run(tabs, writer, enabled):
        nonlocal best, best_error
        start, end, index, err = endpoints(tabs)
        win = enabled & (err < best_error)
        idx = remap_indices(cs, np.repeat(index[:, None], 16, axis=1))
        blk = writer(start, end, idx)
        best = np.where(win[:, None], blk, best)
        best_error = np.where(win, err, best_error)

    is_bc1 = fmt == BC1
    all_ = np.ones(n, bool)
    if is_bc1:                                                                # colourfit.rs:48-59
        run((0, 1, 0), write3, all_)
        run((2, 3, 2), write4, ~cs.transparent)
    else:
        run((2, 3, 2), write4, all_)
    return best


# ---- colourfit/range.rs:44-192 -------------------------------------------------------------------------------------------------------------
def _dot3(a, b):
    return a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1] + a[:, 2] * b[:, 2]          # vec3.rs:51-53


def _snap(v, grid, gridrcp):
    half = F(0.5)
    return np.trunc(grid * v + half) * gridrcp


def range_fit(cs, fmt, weights):
    n = len(cs.count)
    wv = np.array(weights, F)[None, :]
    with np.errstate(all="ignore"):
        principle = principle_component(weighted_covariance(cs))
        start = np.zeros((n, 3), F); end = np.zeros((n, 3), F)
        has = cs.count > 0
        start = np.where(has[:, None], cs.points[:, 0], start); end = start.copy()
        mn = _dot3(start, principle); mx = mn.copy()
        for p in range(1, 16):
            act = p < cs.count
            val = cs.points[:, p]
            d = _dot3(val, principle)
            lower = act & (d < mn)
            upper = act & ~(d < mn) & (d > mx)                               # else-if (range.rs:80)
            start = np.where(lower[:, None], val, start); mn = np.where(lower, d, mn)
            end = np.where(upper[:, None], val, end); mx = np.where(upper, d, mx)
        start = rmin(F(1), rmax(F(0), start)); end = rmin(F(1), rmax(F(0), end))   # one.min(zero.max(x))
        grid = np.array([31, 63, 31], F)[None, :]
        gridrcp = np.array([F(1.0 / 31.0), F(1.0 / 63.0), F(1.0 / 31.0)], F)[None, :]
        start = _snap(start, grid, gridrcp); end = _snap(end, grid, gridrcp)

        best_error = np.full(n, F32_MAX, F)
        best = np.zeros((n, 8), np.uint8)

        def helper(codes, writer, enabled):
            nonlocal best, best_error
            closest = np.zeros((n, 16), np.int64)
            error = np.zeros(n, F)
            for p in range(16):
                act = p < cs.count
                dist = np.full(n, F32_MAX, F); idx = np.zeros(n, np.int64)
                for j, code in enumerate(codes):
                    t = wv * (cs.points[:, p] - code)
                    d = t[:, 0] * t[:, 0] + t[:, 1] * t[:, 1] + t[:, 2] * t[:, 2]     # length2
                    less = d < dist
                    dist = np.where(less, d, dist); idx = np.where(less, j, idx)
                closest[:, p] = np.where(act, idx, 0)
                error = np.where(act, error + dist, error)
            win = enabled & (error < best_error)
            blk = writer(start, end, remap_indices(cs, closest))
            best = np.where(win[:, None], blk, best)
            best_error = np.where(win, error, best_error)

        all_ = np.ones(n, bool)
        c3 = [start, end, start * F(0.5) + end * F(0.5)]
        c4 = [start, end, start * F(2.0 / 3.0) + end * F(1.0 / 3.0), start * F(1.0 / 3.0) + end * F(2.0 / 3.0)]
        if fmt == BC1:
            helper(c3, write3, all_)
            helper(c4, write4, ~cs.transparent)
        else:
            helper(c4, write4, all_)
    return best


# ---- colourfit/cluster.rs:33-418 -------------------------------------------------------------------------------------------------------------
def _insertion_sort_order(count, dots):
    """cluster.rs:82-105: 16 (index, dot) pairs -- (0, f32::MAX) padding beyond count -- sorted with fcmp by the insertion sort
    core::slice::sort_unstable_by uses for short slices; returns the sorted indices (n, 16)."""
    n = len(count)
    idx = np.tile(np.arange(16), (n, 1))
    pad = np.arange(16)[None, :] >= count[:, None]
    idx = np.where(pad, 0, idx)
    val = np.where(pad, F32_MAX, dots)
    fin = np.isfinite(val)

    def less(av, af, bv, bf):       # fcmp(a, b) == Less
        with np.errstate(invalid="ignore"):
            return np.where(~af & ~bf, False, np.where(~af, False, np.where(~bf, True, av < bv)))

    for i in range(1, 16):
        tv, ti, tf = val[:, i].copy(), idx[:, i].copy(), fin[:, i].copy()
        moving = np.ones(n, bool)
        for j in range(i, 0, -1):
            sh = moving & less(tv, tf, val[:, j - 1], fin[:, j - 1])
            # where the predecessor is greater shift it right, elsewhere the element has found its place
            place = moving & ~sh
            val[:, j] = np.where(sh, val[:, j - 1], np.where(place, tv, val[:, j]))
            idx[:, j] = np.where(sh, idx[:, j - 1], np.where(place, ti, idx[:, j]))
            fin[:, j] = np.where(sh, fin[:, j - 1], np.where(place, tf, fin[:, j]))
            moving = sh
        val[:, 0] = np.where(moving, tv, val[:, 0]); idx[:, 0] = np.where(moving, ti, idx[:, 0]); fin[:, 0] = np.where(moving, tf, fin[:, 0])
    return idx


def cluster_fit(cs, fmt, weights, iterate):
    n = len(cs.count)
    count = cs.count
    wv = np.array(list(weights) + [1.0], F)[None, :]
    num_iterations = 8 if iterate else 1
    with np.errstate(all="ignore"):
        principle = principle_component(weighted_covariance(cs))
    pts4 = np.concatenate([cs.points, np.ones((n, 16, 1), F)], axis=2)
    fit_best_error = np.full(n, F32_MAX, F)
    fit_best = np.zeros((n, 8), np.uint8)
    rows = np.arange(n)
    two, one, zero, half = F(2), F(1), F(0), F(0.5)
    grid = np.array([31, 63, 31, 0], F)[None, :]
    gridrcp = np.array([F(1.0 / 31.0), F(1.0 / 63.0), F(1.0 / 31.0), F(0)], F)[None, :]

    def solve(alphax, betax, alphabeta):
        alpha2 = alphax[:, 3:4]; beta2 = betax[:, 3:4]; ab = alphabeta[:, None]
        factor = one / ((alpha2 * beta2) - ab * ab)
        a = ((alphax * beta2) - betax * ab) * factor
        b = ((betax * alpha2) - alphax * ab) * factor
        a = rmin(one, rmax(zero, a)); b = rmin(one, rmax(zero, b))
        a = np.trunc(grid * a + half) * gridrcp
        b = np.trunc(grid * b + half) * gridrcp
        e1 = (a * a) * alpha2 + (b * b * beta2)
        e2 = (a * b * ab) - a * alphax
        e3 = e2 - b * betax
        e4 = two * e3 + e1
        e5 = e4 * wv
        return a, b, e5[:, 0] + e5[:, 1] + e5[:, 2]

    def run_pass(three, enabled):
        nonlocal fit_best, fit_best_error
        order = np.zeros((n, 8, 16), np.int64)
        best_start = np.zeros((n, 4), F); best_end = np.zeros((n, 4), F)
        best_error = fit_best_error.copy()
        best_iteration = np.zeros(n, np.int64)
        best_i = np.zeros(n, np.int64); best_j = np.zeros(n, np.int64); best_k = np.zeros(n, np.int64)
        axis = principle.copy()
        alive = enabled.copy()                                                # blocks still inside the iteration loop
        h3 = np.array([0.5, 0.5, 0.5, 0.25], F)[None, :]
        t13 = np.array([F(1.0 / 3.0), F(1.0 / 3.0), F(1.0 / 3.0), F(1.0 / 9.0)], F)[None, :]
        t23 = np.array([F(2.0 / 3.0), F(2.0 / 3.0), F(2.0 / 3.0), F(4.0 / 9.0)], F)[None, :]
        t29 = F(2.0 / 9.0)
        for it in range(num_iterations):
            if not alive.any():
                break
            # ---- construct_ordering (cluster.rs:78-136)
            with np.errstate(all="ignore"):
                dots = np.stack([_dot3(cs.points[:, p], axis) for p in range(16)], axis=1)
            o = _insertion_sort_order(count, dots)
            order[:, it] = np.where(alive[:, None], o, order[:, it])
            same = np.zeros(n, bool)
            for prev in range(it):
                same |= (order[:, it] == order[:, prev]).all(axis=1)
            alive &= ~same                                                    # return false -> break
            pw = np.zeros((n, 16, 4), F)
            xsum = np.zeros((n, 4), F)
            for p in range(16):
                act = (p < count)[:, None]
                j = order[:, it, p]
                x = pts4[rows, j] * cs.weights[rows, j][:, None]
                pw[:, p] = np.where(act, x, pw[:, p])
                xsum = np.where(act, xsum + x, xsum)
            # ---- the loop nest, on the blocks that are alive (compressed views keep the cost proportional to the work)
            sel = rows[alive]
            if sel.size:
                cnt = count[sel]; spw = pw[sel]; sx = xsum[sel]
                be = best_error[sel].copy(); bs = best_start[sel].copy(); bn = best_end[sel].copy()
                bi = best_i[sel].copy(); bj = best_j[sel].copy(); bk = best_k[sel].copy(); bit = best_iteration[sel].copy()
                m = len(sel)
                part0 = np.zeros((m, 4), F)
                with np.errstate(all="ignore"):
                    for i in range(16):
                        ai = i < cnt
                        if not ai.any():
                            break
                        if three:
                            part1 = spw[:, 0].copy() if i == 0 else np.zeros((m, 4), F)
                            for j in range(1 if i == 0 else i, 17):
                                aj = ai & (j <= cnt)
                                if not aj.any():
                                    break
                                part2 = sx - part1 - part0
                                alphax = part1 * h3 + part0
                                betax = part1 * h3 + part2
                                alphabeta = (part1 * h3)[:, 3]
                                a, b, err = solve(alphax, betax, alphabeta)
                                win = aj & (err < be)
                                bs = np.where(win[:, None], a, bs); bn = np.where(win[:, None], b, bn)
                                bi = np.where(win, i, bi); bj = np.where(win, j, bj)
                                be = np.where(win, err, be); bit = np.where(win, it, bit)
                                if j < 16:
                                    adv = (aj & (j < cnt))[:, None]
                                    part1 = np.where(adv, part1 + spw[:, j], part1)
                        else:
                            part1 = np.zeros((m, 4), F)
                            for j in range(i, 17):
                                aj = ai & (j <= cnt)
                                if not aj.any():
                                    break
                                part2 = spw[:, 0].copy() if j == 0 else np.zeros((m, 4), F)
                                for k in range(1 if j == 0 else j, 17):
                                    ak = aj & (k <= cnt)
                                    if not ak.any():
                                        break
                                    part3 = sx - part2 - part1 - part0
                                    alphax = part2 * t13 + (part1 * t23 + part0)
                                    betax = part1 * t13 + (part2 * t23 + part3)
                                    alphabeta = t29 * (part1 + part2)[:, 3]
                                    a, b, err = solve(alphax, betax, alphabeta)
                                    win = ak & (err < be)
                                    bs = np.where(win[:, None], a, bs); bn = np.where(win[:, None], b, bn)
                                    bi = np.where(win, i, bi); bj = np.where(win, j, bj); bk = np.where(win, k, bk)
                                    be = np.where(win, err, be); bit = np.where(win, it, bit)
                                    if k < 16:
                                        adv = (ak & (k < cnt))[:, None]
                                        part2 = np.where(adv, part2 + spw[:, k], part2)
                                if j < 16:
                                    adv = (aj & (j < cnt))[:, None]
                                    part1 = np.where(adv, part1 + spw[:, j], part1)
                        part0 = np.where(ai[:, None], part0 + spw[:, i], part0)
                best_error[sel] = be; best_start[sel] = bs; best_end[sel] = bn
                best_i[sel] = bi; best_j[sel] = bj; best_k[sel] = bk; best_iteration[sel] = bit
            alive &= best_iteration == it                                     # :243 / :383
            axis = np.where(alive[:, None], (best_end - best_start)[:, :3], axis)   # :248 / :388
        # ---- save the block if necessary (:252-273 / :392-416)
        improved = enabled & (best_error < fit_best_error)
        ordb = order[rows, best_iteration]                                     # (n, 16)
        # the reference runs its `for m in a..b` loops one after the other (two for compress3, three for compress4): a point written
        # by an earlier loop is overwritten by a later one, also through a different m that names the same point (orderings with
        # repeated entries, NaN axis)
        unordered = np.zeros((n, 16), np.int64)
        loops = ((best_i, best_j, 2), (best_j, count, 1)) if three else ((best_i, best_j, 2), (best_j, count, 3), (best_k, count, 1))
        for lo, hi, code in loops:
            for mpos in range(16):
                hit = (mpos >= lo) & (mpos < hi)
                unordered[rows[hit], ordb[hit, mpos]] = code
        idx = remap_indices(cs, unordered)
        blk = (write3 if three else write4)(best_start[:, :3], best_end[:, :3], idx)
        fit_best = np.where(improved[:, None], blk, fit_best)
        fit_best_error = np.where(improved, best_error, fit_best_error)

    all_ = np.ones(n, bool)
    if fmt == BC1:                                                            # colourfit.rs:48-59
        run_pass(True, all_)
        run_pass(False, ~cs.transparent)
    else:
        run_pass(False, all_)
    return fit_best


# ---- lib.rs:188-234 -------------------------------------------------------------------------------------------------------------------------
def compress_blocks(fmt, rgba, mask, algorithm, weights, alpha_weighted=False):
    """n x compress_block_masked: rgba (n,16,4) uint8, mask (n,) -> (n, 8 | 16) uint8"""
    rgba = np.asarray(rgba, np.uint8).reshape(-1, 16, 4)
    mask = np.asarray(mask, np.int64).reshape(-1)
    n = len(rgba)
    if fmt == BC4:
        return compress_bc3(rgba, 0, mask)
    if fmt == BC5:
        return np.concatenate([compress_bc3(rgba, 0, mask), compress_bc3(rgba, 1, mask)], axis=1)
    out = np.zeros((n, 8 if fmt == BC1 else 16), np.uint8)
    if fmt == BC2:
        out[:, :8] = compress_bc2(rgba, mask)
    if fmt == BC3:
        out[:, :8] = compress_bc3(rgba, 3, mask)
    cs = colour_set(rgba, mask, fmt, alpha_weighted)
    colour = np.zeros((n, 8), np.uint8)
    single = cs.count == 1
    rng = ~single & ((algorithm == RANGE_FIT) | (cs.count == 0))
    clu = ~single & ~rng

    def subset(sel):
        s = ColourSet()
        s.count = cs.count[sel]; s.points = cs.points[sel]; s.weights = cs.weights[sel]; s.remap = cs.remap[sel]; s.transparent = cs.transparent[sel]
        return s

    if single.any():
        colour[single] = single_colour_fit(subset(single), fmt)
    if rng.any():
        colour[rng] = range_fit(subset(rng), fmt, weights)
    if clu.any():
        colour[clu] = cluster_fit(subset(clu), fmt, weights, algorithm == ITERATIVE_CLUSTER_FIT)
    out[:, (0 if fmt == BC1 else 8):] = colour
    return out
