"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol, size
arithmetic, sharding, argument validation (no compute calls without a GPU)."""
import ctypes, json, pathlib, re
import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
KAT = json.loads((ROOT / "tests" / "golden" / "kat.json").read_text())


def test_library_exports_every_declared_symbol():
    import texpresso_b200._lib as L
    lib = L.load()
    header = (ROOT / "include" / "texpresso_b200.h").read_text()
    declared = set(re.findall(r"TXP_API[^;(]*?\b(txp_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    for name in declared:
        assert hasattr(lib, name)


@pytest.mark.parametrize("fmt,w,h,size", KAT["sizes"])
def test_storage_requirements(fmt, w, h, size):       # reference lib.rs:350-361
    import texpresso_b200 as T
    assert T.Format(fmt).compressed_size(w, h) == size


def test_block_sizes_and_num_blocks():
    import texpresso_b200 as T
    assert [T.Format(f).block_size() for f in range(5)] == [8, 16, 16, 8, 16]
    assert [T.num_blocks(s) for s in (0, 1, 4, 5, 15, 16)] == [0, 1, 1, 2, 4, 4]
    assert T.Params().algorithm == T.Algorithm.ClusterFit and T.Params().weights == T.COLOUR_WEIGHTS_PERCEPTUAL


def test_shard_rows_partition():
    import texpresso_b200 as T
    for h in (1, 4, 6, 13, 1024, 8192, 8190):
        rows = (h + 3) // 4
        for world in (1, 2, 3, 4, 8):
            spans = [T.shard_rows(h, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == rows
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_argument_validation_without_gpu():
    import texpresso_b200 as T
    px = np.zeros(64, np.uint8)
    with pytest.raises(T.TexpressoError) as e:          # reference: assert!(output.len() >= compressed_size) lib.rs:295
        T.Format.Bc1.compress(px, 4, 4, output=np.zeros(4, np.uint8))
    assert e.value.code == -3
    with pytest.raises(T.TexpressoError) as e:          # rgba too short: slice panic lib.rs:324
        T.Format.Bc1.compress(px[:32], 4, 4)
    assert e.value.code == -3
    with pytest.raises(T.TexpressoError) as e:          # width 0: chunks_mut(0) panics
        T.Format.Bc1.compress(px, 0, 4, output=np.zeros(8, np.uint8))
    assert e.value.code == -2
    with pytest.raises(T.TexpressoError) as e:
        T.Format.Bc3.decompress(np.zeros(8, np.uint8), 4, 4)
    assert e.value.code == -3


def test_synth_is_counter_based():
    from texpresso_b200 import synth
    full = synth.generate("smooth", 64, 48, 7)
    part = synth.generate("smooth", 64, 48, 7, y0=16, y1=32)
    assert np.array_equal(full[16:32], part)
    n = synth.generate("noise_opaque", 32, 8, 1)
    assert (n[..., 3] == 255).all() and len(np.unique(n[..., 0])) > 50
    r = synth.generate("r_rg", 16, 4, 4)
    assert (r[..., 2] == 0).all() and (r[..., 3] == 255).all()


def test_mip_levels_and_sizes():
    import texpresso_b200 as T
    from texpresso_b200 import _lib
    L = _lib.load()
    assert T.mip_levels(1024, 1024)[-1] == (1, 1) and len(T.mip_levels(1024, 1024)) == 11
    assert T.mip_levels(100, 36) == [(100, 36), (50, 18), (25, 9), (12, 4), (6, 2), (3, 1), (1, 1)]
    for (w, h) in ((1024, 1024), (100, 36), (1, 1), (5, 3)):
        lv = T.mip_levels(w, h)
        assert L.txp_mip_levels(w, h) == len(lv)
        for fmt in range(5):
            assert L.txp_mipchain_compressed_size(fmt, w, h) == sum(T.Format(fmt).compressed_size(a, b) for a, b in lv)
    # 87 381 blocks per 1024^2 texture with its full chain (65536+16384+...+1+1+1)
    assert L.txp_mipchain_compressed_size(2, 1024, 1024) // 16 == 87383


def test_generate_mips_box_filter():
    import texpresso_b200 as T
    img = np.arange(4 * 6 * 4, dtype=np.uint8).reshape(4, 6, 4)
    lv = T.generate_mips(img, 6, 4)
    assert [l.shape[:2] for l in lv] == [(4, 6), (2, 3), (1, 1)]
    a = img.astype(np.int32)
    assert np.array_equal(lv[1][0, 0], (a[0, 0] + a[0, 1] + a[1, 0] + a[1, 1] + 2) >> 2)
    b = lv[1].astype(np.int32)                       # 3 wide -> 1 wide: columns 0,1 (clamped sampling never reaches col 2)
    assert np.array_equal(lv[2][0, 0], (b[0, 0] + b[0, 1] + b[1, 0] + b[1, 1] + 2) >> 2)


def test_batch_wrappers_validate_sizes_before_calling_c():
    """the batch entry points take no buffer lengths, so the Python mirror must reject short / strided buffers itself.
    These raise before any C call, i.e. they run without a GPU."""
    import numpy as np
    import pytest
    import texpresso_b200 as T
    good = np.zeros(16 * 16 * 4, np.uint8)
    with pytest.raises(ValueError):
        T.compress_batch(0, [(good[:100], 16, 16)])
    with pytest.raises(ValueError):
        T.compress_batch(0, [(good, 16, 16)], outputs=[np.zeros(8, np.uint8)])
    with pytest.raises(ValueError):
        T.compress_batch(0, [(good, 16, 16)], outputs=[np.zeros(256, np.uint8)[::2]])
    with pytest.raises(ValueError):
        T.compress_batch_mips(2, [(good[:100], 16, 16)])
    with pytest.raises(ValueError):
        T.compress_batch_mips(2, [(good, 16, 16)], outputs=[np.zeros(16, np.uint8)])
    with pytest.raises(ValueError):
        T.decompress_batch(0, [(np.zeros(8, np.uint8), 16, 16)])
    with pytest.raises(ValueError):
        T.compress_batch(0, [(good, 16, 16), (good, 16, 16)], outputs=[np.zeros(128, np.uint8)])
    for fn in (lambda o: T.compress_pixels(0, good, 16, 16, output=o), lambda o: T.compress_multi(0, good, 16, 16, output=o),
               lambda o: T.Format.Bc1.compress_block_masked(good[:64], 0xFFFF, output=o), lambda o: T.decompress_multi(0, np.zeros(128, np.uint8), 16, 16, output=o)):
        with pytest.raises(ValueError):
            fn(np.zeros(4096, np.uint8)[::2])               # non-contiguous output: would be written through a temporary copy


def test_debug_knobs_reject_bad_values():
    from texpresso_b200 import _lib
    L = _lib.load()
    assert L.txp_debug_set(1, -5) != 0
    assert L.txp_debug_set(1, 0) != 0
    assert L.txp_debug_set(0, 7) != 0


def test_copy_pool_matches_memcpy():
    """The multi-threaded staging copy used for pageable caller buffers (CopyPool): every size class (below the split threshold, odd sizes,
    unaligned ends, many parts) and several host threads copying at once give plain-memcpy results.  Host-only code: runs without a GPU."""
    import ctypes, threading
    from texpresso_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(5)
    for n in (0, 1, 4095, (512 << 10) - 1, 1 << 20, (1 << 20) + 1, (5 << 20) + 3, (24 << 20) + 4097):
        src = rng.integers(0, 256, n, dtype=np.uint8)
        dst = np.zeros(n + 64, np.uint8)
        assert L.txp_debug_host_copy(ctypes.c_void_p(dst.ctypes.data + 7), ctypes.c_void_p(src.ctypes.data), n) == 0
        assert np.array_equal(dst[7:7 + n], src) and not dst[:7].any() and not dst[7 + n:].any()
    srcs = [rng.integers(0, 256, (3 << 20) + 11 * i, dtype=np.uint8) for i in range(8)]
    dsts = [np.zeros(s.size, np.uint8) for s in srcs]
    def work(i):
        for _ in range(6):
            L.txp_debug_host_copy(ctypes.c_void_p(dsts[i].ctypes.data), ctypes.c_void_p(srcs[i].ctypes.data), srcs[i].size)
    ts = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert all(np.array_equal(a, b) for a, b in zip(srcs, dsts))


def _plan(fmt, alg, w, rows, sms=148):
    import texpresso_b200 as T
    from texpresso_b200 import _lib
    L = _lib.load()
    p = T.Params(T.Algorithm(alg))._c()
    out = (ctypes.c_size_t * 4096)()
    n = ctypes.c_size_t(0)
    lane = ctypes.c_uint64(0)
    _lib.check(L.txp_debug_plan(fmt, ctypes.byref(p), w, rows, sms, out, 4096, ctypes.byref(n), ctypes.byref(lane)))
    return list(out[:n.value]), lane.value


def test_pipeline_chunk_plans():
    """The host pipeline's chunk plan (pipeline_plan, host-only): the round-aligned plans of the metric texture and its shards on a 148-SM device,
    and the invariants every plan must keep for any width / height / algorithm (the chunks cover the rows exactly, none is empty, lane chunks of
    the round-aligned plan never exceed whole rounds of 148 x 6 x 128 lanes and come after the small warp-per-block head)."""
    wave = 148 * 6 * 128
    # one rank of eight: 256 block rows of the 8192-wide texture = 4.61 rounds -> 0.61 round first (8 + 27 rows), then 2 + 1 + 1 rounds
    assert _plan(2, 1, 8192, 256) == ([8, 27, 111, 55, 55], 0b11100)
    assert _plan(0, 1, 8192, 256) == ([8, 27, 111, 55, 55], 0b11100)
    rows, lane = _plan(2, 1, 8192, 512)                       # one rank of four: 9.2 rounds
    assert rows == [13, 111, 111, 111, 111, 55] and lane == 0b111110
    rows, lane = _plan(2, 1, 8192, 2048)                      # the whole texture: 36.9 rounds
    assert sum(rows) == 2048 and rows[2:-2] == [111] * 17 and rows[-2:] == [55, 55] and lane == ((1 << len(rows)) - 1) & ~0b11
    # below 4 rounds, above 40 rounds, RangeFit, IterativeClusterFit, BC4: the uniform plans, no forced lane chunks
    for fmt, alg, w, r in ((2, 1, 8192, 192), (2, 1, 16384, 4096), (2, 0, 8192, 2048), (0, 2, 8192, 2048), (3, 1, 16384, 4096)):
        rows, lane = _plan(fmt, alg, w, r)
        assert sum(rows) == r and lane == 0 and min(rows) >= 1, (fmt, alg, w, r, rows)
    assert _plan(0, 2, 8192, 2048)[0] == [64, 992, 992]       # IterativeClusterFit: half a lane launch, then two chunks of 124 MiB
    assert _plan(2, 1, 8192, 0) == ([], 0)
    rng = np.random.default_rng(11)
    for _ in range(3000):
        fmt, alg = int(rng.integers(0, 5)), int(rng.integers(0, 3))
        w = int(rng.choice([1, 3, 4, 64, 100, 1000, 2048, 4096, 6000, 8192, 16384, 30000, 500000]))
        r = int(rng.integers(0, 5000)) if w >= 1000 else int(rng.integers(0, 200000))
        sms = int(rng.choice([148, 132, 20, 1]))
        rows, lane = _plan(fmt, alg, w, r, sms)
        bw = (w + 3) // 4
        assert sum(rows) == r and all(x >= 1 for x in rows), (fmt, alg, w, r, sms, rows[:8])
        if lane:
            wv = sms * 6 * 128
            first_lane = min(i for i in range(len(rows)) if (lane >> i) & 1)
            assert fmt <= 2 and alg == 1 and 1 <= first_lane <= 2
            assert all((lane >> i) & 1 for i in range(first_lane, len(rows)))
            for i in range(first_lane, len(rows)):
                k = max(1, round(rows[i] * bw / wv))
                assert rows[i] * bw <= k * wv and rows[i] * bw >= 32768, (w, r, sms, rows, i)
            assert sum(rows[:first_lane]) * bw < 2 * wv + 16384 + 2 * bw
