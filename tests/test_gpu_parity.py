"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): bit-exact for RangeFit, SingleColourFit, BC2, the alpha/BC4/BC5 paths and
all decoders; ClusterFit / IterativeClusterFit >= 99.9 % identical blocks with every differing block no
worse in the reference's own weighted squared error."""
import json, pathlib
import numpy as np
import pytest

from tests import oracle_lib as O
from tests import blockgen

pytestmark = pytest.mark.gpu

KAT = json.loads((pathlib.Path(__file__).parent / "golden" / "kat.json").read_text())
SETS = sorted(KAT["sets"].items())
ALGS = [0, 1, 2]
WEIGHTS = {"uniform": O.UNIFORM, "perceptual": O.PERCEPTUAL, "odd": (0.3, 1.7, 0.05)}


@pytest.fixture(scope="module")
def T():
    import texpresso_b200 as T
    assert T.device_count() >= 1
    return T


def _params(T, alg, w, awa=False):
    return T.Params(T.Algorithm(alg), tuple(w), awa), O.make_params(alg, w, awa)


# ---- the reference's own known-answer tests, through the GPU path (lib.rs:363-504) -----------------------
@pytest.mark.parametrize("alg", ALGS)
@pytest.mark.parametrize("name,ds", SETS)
def test_kat_compression(T, name, ds, alg):
    tp, _ = _params(T, alg, O.UNIFORM)
    out = T.Format(ds["format"]).compress(np.array(ds["decoded"], np.uint8), 4, 4, tp)
    assert bytes(out).hex() == bytes(ds["encoded"]).hex()


@pytest.mark.parametrize("name,ds", SETS)
def test_kat_decompression(T, name, ds):
    out = T.Format(ds["format"]).decompress(np.array(ds["encoded"], np.uint8), 4, 4)
    assert out.tolist() == ds["decoded"]


def test_kat_decode_height_not_multiple_of_4(T):
    d = KAT["decode_4x6"]
    out = T.Format(d["format"]).decompress(np.array(d["encoded"], np.uint8), d["width"], d["height"])
    assert out.reshape(-1, 4).tolist() == [d["pixel"]] * 24


# ---- differential layer: stratified blocks through compress_block_masked semantics -----------------------
def _compare_blocks(T, fmt, blocks, masks, tags, alg, wname, awa):
    tp, op = _params(T, alg, WEIGHTS[wname], awa)
    got = T.compress_blocks(fmt, blocks, masks, tp)
    want = O.compress_blocks(fmt, blocks, masks, op)
    diff = np.nonzero((got != want).any(axis=1))[0]
    return got, want, diff, op


@pytest.mark.parametrize("awa", [False, True])
@pytest.mark.parametrize("wname", ["uniform", "perceptual", "odd"])
@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_rangefit_blocks_bit_exact(T, fmt, wname, awa):
    blocks, masks, tags = blockgen.colour_cases()
    got, want, diff, _ = _compare_blocks(T, fmt, blocks, masks, tags, 0, wname, awa)
    assert diff.size == 0, [(int(i), tags[i], bytes(got[i]).hex(), bytes(want[i]).hex()) for i in diff[:8]]


@pytest.mark.parametrize("awa", [False, True])
@pytest.mark.parametrize("wname", ["uniform", "perceptual", "odd"])
@pytest.mark.parametrize("alg", [1, 2])
@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_clusterfit_blocks(T, fmt, alg, wname, awa):
    """Default (warp-per-block) ClusterFit / IterativeClusterFit path on the stratified corpus: bit-exact, as DESIGN.md and
    INTEGRATION.md claim (the north star's >= 99.9 % tolerance is not used: there are no known divergent blocks)."""
    blocks, masks, tags = blockgen.colour_cases()
    got, want, diff, op = _compare_blocks(T, fmt, blocks, masks, tags, alg, wname, awa)
    assert diff.size == 0, [(int(i), tags[i], bytes(got[i]).hex(), bytes(want[i]).hex()) for i in diff[:8]]


@pytest.mark.parametrize("fmt", [2, 3, 4])
def test_alpha_paths_bit_exact(T, fmt):
    blocks, masks, tags = blockgen.alpha_cases()
    tp, op = _params(T, 0, O.UNIFORM)
    got = T.compress_blocks(fmt, blocks, masks, tp)
    want = O.compress_blocks(fmt, blocks, masks, op)
    n = 8 if fmt != 4 else 16
    diff = np.nonzero((got[:, :n] != want[:, :n]).any(axis=1))[0]
    assert diff.size == 0, [(int(i), tags[i], bytes(got[i]).hex(), bytes(want[i]).hex()) for i in diff[:8]]


def test_bc2_alpha_all_values(T):
    blocks = np.zeros((16, 16, 4), np.uint8)
    blocks[..., 3] = np.arange(256, dtype=np.uint8).reshape(16, 16)
    masks = np.full(16, 0xFFFF, np.uint32); masks[5] = 0x0F0F
    tp, op = _params(T, 0, O.UNIFORM)
    got = T.compress_blocks(1, blocks, masks, tp)
    want = O.compress_blocks(1, blocks, masks, op)
    assert np.array_equal(got, want)


def test_single_block_entry_points(T):
    blocks, masks, tags = blockgen.colour_cases(per_class=2)
    for i in range(0, len(blocks), 37):
        for fmt in range(5):
            tp, op = _params(T, 1, O.PERCEPTUAL)
            got = T.Format(fmt).compress_block_masked(blocks[i], int(masks[i]), tp)
            want = O.compress_block_masked(fmt, blocks[i], int(masks[i]), op)
            assert bytes(got) == bytes(want), (fmt, tags[i])
            assert np.array_equal(T.Format(fmt).decompress_block(want).reshape(-1), O.decompress_block(fmt, want))


# ---- decoders: bit exact on arbitrary (random) blocks -------------------------------------------------------
@pytest.mark.parametrize("fmt", range(5))
def test_decode_random_blocks(T, fmt):
    rng = np.random.default_rng(5 + fmt)
    bs = 8 if fmt in (0, 3) else 16
    n = 4096
    blocks = rng.integers(0, 256, size=(n, bs), dtype=np.uint8)
    blocks[: n // 4, 0:2] = blocks[: n // 4, 2:4]          # equal endpoints / a0 == a1 cases
    got = T.decompress_blocks(fmt, blocks)
    want = O.decompress_blocks(fmt, blocks)
    assert np.array_equal(got, want)


# ---- image layer -----------------------------------------------------------------------------------------------
SIZES = [(1, 1), (2, 2), (3, 5), (4, 6), (13, 7), (16, 4), (64, 64), (100, 36), (257, 63)]


@pytest.mark.parametrize("w,h", SIZES)
@pytest.mark.parametrize("fmt", range(5))
def test_image_encode_decode(T, fmt, w, h):
    from texpresso_b200 import synth
    img = synth.generate("smooth" if (w * h) % 2 else "noise_alpha", w, h, seed=w * 1000 + h)
    for alg in ((0, 1, 2) if fmt < 3 else (1,)):
        tp, op = _params(T, alg, O.PERCEPTUAL)
        got = T.Format(fmt).compress(img, w, h, tp)
        want = O.compress(fmt, img, w, h, op)
        assert np.array_equal(got, want), (fmt, alg, w, h)
    dec = T.Format(fmt).decompress(want, w, h)
    assert np.array_equal(dec, O.decompress(fmt, want, w, h))


@pytest.mark.parametrize("fmt", range(5))
def test_output_longer_than_needed_encodes_masked_rows(T, fmt):     # SURVEY Q13 (lib.rs:300-334)
    from texpresso_b200 import synth
    w, h = 20, 8
    img = synth.generate("noise_alpha", w, h, seed=3)
    bs = 8 if fmt in (0, 3) else 16
    n = T.Format(fmt).compressed_size(w, h) + 2 * 5 * bs + 3 * bs       # two extra rows + a partial row
    out = np.full(n, 0xEE, np.uint8)
    tp, op = _params(T, 1, O.PERCEPTUAL)
    T.Format(fmt).compress(img, w, h, tp, output=out)
    want = O.compress(fmt, img, w, h, op, out_len=n)
    assert np.array_equal(out, want)


def test_medium_image_all_formats(T):
    from texpresso_b200 import synth
    w = h = 256
    for kind in ("smooth", "noise_alpha", "noise_opaque"):
        img = synth.generate(kind, w, h, seed=11)
        for fmt in range(5):
            for alg in ((0, 1, 2) if fmt < 3 else (1,)):
                tp, op = _params(T, alg, O.PERCEPTUAL)
                got = T.Format(fmt).compress(img, w, h, tp)
                want = O.compress(fmt, img, w, h, op, threads=8)
                bs = 8 if fmt in (0, 3) else 16
                nd = int((got.reshape(-1, bs) != want.reshape(-1, bs)).any(axis=1).sum())
                assert nd == 0, (kind, fmt, alg, nd)


# ---- full-size properties (BASELINE.json sizes), oracle only on slices --------------------------------------
def test_fullsize_8192_bc1_bc3_clusterfit_properties(T):
    from texpresso_b200 import synth
    w = h = 8192
    img = synth.generate("noise_alpha", w, h, seed=3)
    tp, op = _params(T, 1, O.PERCEPTUAL)
    for fmt in (0, 2):
        a = T.Format(fmt).compress(img, w, h, tp)
        b = T.Format(fmt).compress(img, w, h, tp)
        assert np.array_equal(a, b), "run-to-run determinism"
        # block-row shards encode to the same bytes as the whole image (no cross-block state)
        bs = 8 if fmt == 0 else 16
        rowbytes = (w // 4) * bs
        r0, r1 = T.shard_rows(h, 3, 8)
        part = T.Format(fmt).compress(img[4 * r0:4 * r1], w, 4 * (r1 - r0), tp)
        assert np.array_equal(part, a[r0 * rowbytes:r1 * rowbytes])
        # oracle on 8 block rows taken from the middle
        y0 = 4096
        want = O.compress(fmt, img[y0:y0 + 32], w, 32, op, threads=8)
        got = a[(y0 // 4) * rowbytes:(y0 // 4 + 8) * rowbytes]
        nd = int((got.reshape(-1, bs) != want.reshape(-1, bs)).any(axis=1).sum())
        assert nd == 0, (fmt, nd)
        # decoder round trip equals the oracle's decode of the same blocks
        dec = T.Format(fmt).decompress(a[:rowbytes * 16], w, 64)
        assert np.array_equal(dec, O.decompress(fmt, a[:rowbytes * 16], w, 64))


def test_fullsize_16384_bc4_bc5(T):
    from texpresso_b200 import synth
    w = h = 16384
    img = synth.generate("r_rg", w, h, seed=4)
    tp, op = _params(T, 1, O.PERCEPTUAL)
    for fmt in (3, 4):
        bs = 8 if fmt == 3 else 16
        a = T.Format(fmt).compress(img, w, h, tp)
        rowbytes = (w // 4) * bs
        y0 = 8000
        want = O.compress(fmt, img[y0:y0 + 64], w, 64, op, threads=8)
        assert np.array_equal(a[(y0 // 4) * rowbytes:(y0 // 4 + 16) * rowbytes], want)
        # decoder on the same blocks == oracle decoder (a quality bound is not asserted: the reference's 7-point
        # codebook quirk Q1 legitimately produces large errors on blocks holding both 0/255 and mid values)
        dec = T.Format(fmt).decompress(a[:rowbytes * 64], w, 256)
        assert np.array_equal(dec, O.decompress(fmt, a[:rowbytes * 64], w, 256))
        del a


def test_multi_gpu_matches_single(T):
    if T.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from texpresso_b200 import synth
    w, h = 1024, 1000
    img = synth.generate("smooth", w, h, seed=8)
    tp, _ = _params(T, 1, O.PERCEPTUAL)
    for fmt in (0, 2, 4):
        one = T.Format(fmt).compress(img, w, h, tp)
        for n in (2, T.device_count()):
            assert np.array_equal(T.compress_multi(fmt, img, w, h, tp, n_gpus=n), one)


# ---- mip chains (extension, BASELINE config 5): device mip filter + one-launch multi-level encode -----------------
@pytest.mark.parametrize("w,h", [(64, 64), (100, 36), (16, 4), (1, 1), (5, 3), (256, 128)])
@pytest.mark.parametrize("fmt,alg", [(0, 0), (2, 1), (4, 1), (1, 2)])
def test_mipchain_matches_oracle_per_level(T, fmt, alg, w, h):
    from texpresso_b200 import synth
    img = synth.generate("smooth" if w >= 8 else "noise_alpha", w, h, seed=77)
    tp, op = _params(T, alg, O.PERCEPTUAL)
    got = T.compress_mipchain(fmt, img, w, h, tp)
    want = np.concatenate([O.compress(fmt, lv, lv.shape[1], lv.shape[0], op) for lv in T.generate_mips(img, w, h)])
    assert got.size == want.size
    assert np.array_equal(got, want)


def test_batch_with_mips_matches_single_calls(T):
    from texpresso_b200 import synth
    tp, _ = _params(T, 1, O.PERCEPTUAL)
    texs = [(synth.generate("smooth", 64 >> (t % 3), 32, seed=100 + t), 64 >> (t % 3), 32) for t in range(7)]
    outs = T.compress_batch_mips(2, texs, tp, n_gpus=1)
    for (img, w, h), o in zip(texs, outs):
        assert np.array_equal(o, T.compress_mipchain(2, img, w, h, tp))
    if T.device_count() >= 2:
        outs2 = T.compress_batch_mips(2, texs, tp, n_gpus=2)
        assert all(np.array_equal(a, b) for a, b in zip(outs, outs2))


@pytest.mark.parametrize("fmt,alg", [(0, 1), (2, 0), (3, 1), (1, 2)])
def test_batch_matches_single_calls(T, fmt, alg):
    """txp_compress_batch: textures pipelined through the slots (small ones) or the chunked path (> 32 MiB) give the bytes
    of lone Format.compress calls, ragged sizes included."""
    from texpresso_b200 import synth
    tp, _ = _params(T, alg, O.PERCEPTUAL)
    sizes = [(64, 32), (37, 23), (256, 256), (4, 4), (1, 7), (128, 20), (512, 64), (96, 96), (33, 3)]
    if alg == 0:
        sizes.append((4096, 2056))                           # > 32 MiB: chunked path in the middle of the batch
        sizes.append((40, 40))
    texs = [(synth.generate("smooth" if i % 2 else "noise_alpha", w, h, seed=300 + i), w, h) for i, (w, h) in enumerate(sizes)]
    outs = T.compress_batch(fmt, texs, tp, n_gpus=1)
    for (img, w, h), o in zip(texs, outs):
        assert np.array_equal(o, T.Format(fmt).compress(img, w, h, tp)), (w, h)
    if T.device_count() >= 2:
        outs2 = T.compress_batch(fmt, texs, tp, n_gpus=2)
        assert all(np.array_equal(a, b) for a, b in zip(outs, outs2))


def test_concurrent_host_calls_are_thread_safe(T):
    """The C ABI is documented as thread-safe: several host threads encoding different images at once (ctypes drops the
    GIL) must each get the bytes a lone call produces."""
    import threading
    from texpresso_b200 import synth
    jobs = [(fmt, alg, synth.generate("smooth" if k % 2 else "noise_alpha", 96 + 4 * k, 64 + 4 * k, seed=200 + k), 96 + 4 * k, 64 + 4 * k)
            for k, (fmt, alg) in enumerate([(0, 1), (2, 1), (3, 1), (1, 0), (0, 2), (4, 1), (2, 2), (0, 0)])]
    expect = [T.Format(f).compress(img, w, h, T.Params(T.Algorithm(a))) for f, a, img, w, h in jobs]
    got = [None] * len(jobs)
    errs = []

    def work(i):
        try:
            f, a, img, w, h = jobs[i]
            for _ in range(4):
                got[i] = T.Format(f).compress(img, w, h, T.Params(T.Algorithm(a)))
        except Exception as e:          # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errs, errs
    assert all(np.array_equal(g, e) for g, e in zip(got, expect))


# ---- round 2: host-API kernel selection, decode sharding / batching ------------------------------------------------
def _debug_get(key):
    import ctypes
    from texpresso_b200 import _lib
    v = ctypes.c_uint64(0)
    _lib.check(_lib.load().txp_debug_get(key, ctypes.byref(v)))
    return v.value


def test_fullsize_8192_bc1_iterative_auto_path_vs_oracle(T):
    """BASELINE config 3 through the public host API: the pipeline chunks must be large enough for the lane-per-block
    iterative kernels (>= 786 432 blocks per launch), and the bytes must equal the oracle's on slices of the full-size output --
    including blocks in the quarter-sized first chunk (warp-per-block kernel) and in later chunks (lane kernels, BC1 carry path)."""
    from texpresso_b200 import synth
    w = h = 8192
    img = synth.generate("noise_opaque", w, h, seed=3)
    tp, op = _params(T, 2, O.PERCEPTUAL)
    lane0, warp0 = _debug_get(3), _debug_get(4)
    a = T.Format.Bc1.compress(img, w, h, tp)
    lane, warp = _debug_get(3) - lane0, _debug_get(4) - warp0
    assert lane >= 2 and warp <= 1, (lane, warp)             # a 64-row first chunk (warp kernels), then two chunks of 124 MiB
    rowbytes = (w // 4) * 8
    for y0 in (0, 1024, 4096 + 512, h - 32):                  # first chunk, chunk interiors, last rows
        want = O.compress(0, img[y0:y0 + 32, :2048], 2048, 32, op, threads=8).reshape(8, -1)
        got = a[(y0 // 4) * rowbytes:(y0 // 4 + 8) * rowbytes].reshape(8, rowbytes)[:, :512 * 8]
        nd = int((got.reshape(-1, 8) != want.reshape(-1, 8)).any(axis=1).sum())
        assert nd == 0, (y0, nd)
    # the device-pointer entry point (one launch over 4 Mi blocks) produces the same bytes
    import torch, ctypes
    from texpresso_b200 import _lib
    d_in = torch.from_numpy(img.reshape(-1)).cuda()
    d_out = torch.empty(a.size, dtype=torch.uint8, device="cuda")
    cp = tp._c()
    _lib.check(_lib.load().txp_compress_device(0, ctypes.c_void_p(d_in.data_ptr()), w, h, ctypes.byref(cp), ctypes.c_void_p(d_out.data_ptr()),
                                               d_out.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), a)


def test_lane_chunk_rows_round_up(T):
    """Widths for which 16 MiB is not a whole number of block rows still reach the lane-per-block kernels (chunk rows are rounded up)."""
    from texpresso_b200 import synth
    w, h = 6000, 2400                                         # 1500 blocks per row: 174 rows would be 261 000 < 262 144
    img = synth.generate("noise_opaque", w, h, seed=12)
    tp, op = _params(T, 1, O.PERCEPTUAL)
    lane0 = _debug_get(2)
    a = T.Format.Bc1.compress(img, w, h, tp)
    assert _debug_get(2) - lane0 >= 3
    want = O.compress(0, img[800:832], w, 32, op, threads=8)
    assert np.array_equal(a[200 * 1500 * 8:208 * 1500 * 8], want)


@pytest.mark.parametrize("fmt", range(5))
def test_empty_and_one_pixel_images(T, fmt):
    """height 0: compressed_size is 0 and the reference's loops run zero times (lib.rs:300-305, :128-134); 1 x 1: one block with 15 masked
    pixels.  Every host entry point must agree with the oracle on both."""
    tp, op = _params(T, 1, O.PERCEPTUAL)
    F = T.Format(fmt)
    empty = np.zeros(0, np.uint8)
    assert F.compressed_size(16, 0) == 0
    assert F.compress(empty, 16, 0, tp).size == 0
    assert F.decompress(empty, 16, 0).size == 0
    assert T.compress_multi(fmt, empty, 16, 0, tp, n_gpus=1).size == 0
    px = np.array([[[200, 100, 50, 128]]], np.uint8)
    got = F.compress(px, 1, 1, tp)
    assert np.array_equal(got, O.compress(fmt, px, 1, 1, op))
    assert np.array_equal(F.decompress(got, 1, 1), O.decompress(fmt, got, 1, 1))
    outs = T.compress_batch(fmt, [(px, 1, 1), (empty, 8, 0), (px, 1, 1)], tp, n_gpus=1)
    assert np.array_equal(outs[0], got) and outs[1].size == 0 and np.array_equal(outs[2], got)


@pytest.mark.parametrize("fmt", [0, 2])
def test_round_aligned_plan_shard_vs_oracle(T, fmt):
    """One rank's shard of the metric texture at 8 ranks (8192 x 1024 = 4.6 rounds of the lane-per-block search) through the host API:
    the round-aligned chunk plan (two small warp-per-block chunks, then lane chunks of whole rounds) gives the bytes of the uniform plan
    and of the oracle on slices across every chunk boundary; a ragged height (partly masked last block row) takes the same path."""
    import ctypes
    from texpresso_b200 import synth, _lib
    L = _lib.load()
    w = 8192
    tp, op = _params(T, 1, O.PERCEPTUAL)
    bs = 8 if fmt == 0 else 16
    for h in (1024, 1022):
        img = synth.generate("noise_opaque" if fmt == 0 else "noise_alpha", w, h, seed=3)
        lane0, warp0 = _debug_get(2), _debug_get(4)
        a = T.Format(fmt).compress(img, w, h, tp)
        lane, warp = _debug_get(2) - lane0, _debug_get(4) - warp0
        assert lane == 3 and warp == 2, (lane, warp)         # chunks of 8, 27 | 111, 55, 55 block rows
        _lib.check(L.txp_debug_set(5, 0))
        try:
            b = T.Format(fmt).compress(img, w, h, tp)
        finally:
            _lib.check(L.txp_debug_set(5, 4))
        assert np.array_equal(a, b)
        rowbytes = (w // 4) * bs
        for r0 in (4, 31, 142, 197, 248):                     # block rows around the chunk boundaries 8, 35, 146, 201 and the end
            y0, y1 = 4 * r0, min(4 * r0 + 32, h)
            want = O.compress(fmt, img[y0:y1, 4096:6144], 2048, y1 - y0, op, threads=8).reshape(8, -1)
            got = a[r0 * rowbytes:(r0 + 8) * rowbytes].reshape(8, rowbytes)[:, 1024 * bs:1536 * bs]
            assert np.array_equal(got, want), (h, r0)


@pytest.mark.parametrize("fmt", range(5))
def test_decompress_multi_and_batch(T, fmt):
    from texpresso_b200 import synth
    tp, _ = _params(T, 0, O.PERCEPTUAL)
    sizes = [(1024, 1000), (37, 23), (256, 64), (4, 4), (1, 7), (2048, 1030)]
    texs = [(synth.generate("smooth" if i % 2 else "noise_alpha", w, h, seed=500 + i), w, h) for i, (w, h) in enumerate(sizes)]
    enc = [T.Format(fmt).compress(img, w, h, tp) for img, w, h in texs]
    want = [O.decompress(fmt, e, w, h) for e, (_, w, h) in zip(enc, texs)]
    for n in sorted({1, min(2, T.device_count()), T.device_count()}):
        for e, (_, w, h), wnt in zip(enc, texs, want):
            assert np.array_equal(T.decompress_multi(fmt, e, w, h, n_gpus=n), wnt), (w, h, n)
        got = T.decompress_batch(fmt, [(e, w, h) for e, (_, w, h) in zip(enc, texs)], n_gpus=n)
        assert all(np.array_equal(g, wnt) for g, wnt in zip(got, want)), n


def test_decompress_large_pipelined(T):
    """Format.decompress on an image of several pipeline chunks (ragged height) equals the oracle."""
    from texpresso_b200 import synth
    w, h = 4096, 4090
    img = synth.generate("noise_alpha", w, h, seed=21)
    enc = T.Format.Bc3.compress(img, w, h, T.Params(T.Algorithm.RangeFit))
    assert np.array_equal(T.Format.Bc3.decompress(enc, w, h), O.decompress(2, enc, w, h))


def test_multi_gpu_all_entry_points(T):
    """txp_compress_multi / txp_compress_batch{,_mips} / txp_decompress_{multi,batch} with n_gpus >= 2 (skipped on a 1-GPU box;
    bench.py runs the same comparison under torchrun: `multi_matches_single`)."""
    if T.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from texpresso_b200 import synth
    n = T.device_count()
    w, h = 4096, 4090
    img = synth.generate("noise_alpha", w, h, seed=31)
    for fmt, alg in ((0, 1), (2, 2), (4, 1)):
        tp, _ = _params(T, alg, O.PERCEPTUAL)
        one = T.Format(fmt).compress(img, w, h, tp)
        assert np.array_equal(T.compress_multi(fmt, img, w, h, tp, n_gpus=n), one)
        assert np.array_equal(T.decompress_multi(fmt, one, w, h, n_gpus=n), T.Format(fmt).decompress(one, w, h))


@pytest.mark.parametrize("fmt,alg", [(2, 1), (0, 2), (4, 1), (1, 0)])
def test_batch_groups_match_single_calls(T, fmt, alg):
    """txp_compress_batch{,_mips}: consecutive textures of one shape share ONE mip-chain launch and ONE encode launch
    (texture-group mode of BlockSource); runs of equal shapes, shape changes, odd sizes, pinned and pageable buffers mixed."""
    import torch
    from texpresso_b200 import synth
    tp, _ = _params(T, alg, O.PERCEPTUAL)
    shapes = [(128, 96)] * 19 + [(64, 64)] * 5 + [(100, 36)] * 3 + [(256, 256)] * 9 + [(1, 1)] * 2 + [(5, 3)] * 4 + [(192, 128)]
    texs = []
    keep = []
    for i, (w, h) in enumerate(shapes):
        img = synth.generate("smooth" if i % 3 else "noise_alpha", w, (h + 3) // 4 * 4, seed=700 + i)[:h]
        img = np.ascontiguousarray(img)
        if i % 4 == 1:                                          # every fourth texture in pinned memory (DMA'd directly)
            t = torch.from_numpy(img.reshape(-1).copy()).pin_memory(); keep.append(t); img = t.numpy()
        texs.append((img, w, h))
    outs = T.compress_batch_mips(fmt, texs, tp, n_gpus=1)
    for (img, w, h), o in zip(texs, outs):
        want = np.concatenate([O.compress(fmt, lv, lv.shape[1], lv.shape[0], O.make_params(alg, O.PERCEPTUAL, False)) for lv in T.generate_mips(img, w, h)])
        assert np.array_equal(o, want), (w, h)
    outs = T.compress_batch(fmt, texs, tp, n_gpus=1)
    for (img, w, h), o in zip(texs, outs):
        assert np.array_equal(o, O.compress(fmt, img, w, h, O.make_params(alg, O.PERCEPTUAL, False))), (w, h)


@pytest.mark.parametrize("w,h", [(1024, 1024), (2048, 512), (4096, 4096), (777, 333), (64, 64), (65, 129), (8192, 16)])
def test_mip_chain_kernel_levels(T, w, h):
    """the one-launch mip chain (64x64 tiles + last-CTA tail; per-level launches above 4096) equals the numpy statement of the
    filter on every level: checked through BC4, whose blocks decode the red channel exactly when a block is flat... so instead
    compare the encoded chain with the oracle's encoding of the numpy-generated levels"""
    from texpresso_b200 import synth
    img = synth.generate("noise_alpha", w, (h + 3) // 4 * 4, seed=91)[:h]
    img = np.ascontiguousarray(img)
    tp, op = _params(T, 0, O.PERCEPTUAL)
    got = T.compress_mipchain(0, img, w, h, tp)
    want = np.concatenate([O.compress(0, lv, lv.shape[1], lv.shape[0], op, threads=8) for lv in T.generate_mips(img, w, h)])
    assert np.array_equal(got, want)


@pytest.mark.parametrize("fmt,alg", [(0, 1), (2, 0), (4, 1), (1, 2)])
def test_device_tensor_entry_points(T, fmt, alg):
    """compress_device / decompress_device (torch CUDA tensors, asynchronous on torch's stream) give the bytes of the host calls; wrong
    tensors are rejected before anything is launched."""
    import torch
    from texpresso_b200 import synth
    w, h = 260, 134
    img = synth.generate("smooth", w, h, seed=5)
    tp, op = _params(T, alg, O.PERCEPTUAL)
    d_in = torch.from_numpy(img.reshape(-1)).cuda()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    d_out = T.compress_device(fmt, d_in, w, h, tp, stream=side)
    d_dec = T.decompress_device(fmt, d_out, w, h, stream=side)
    side.synchronize()
    want = O.compress(fmt, img, w, h, op)
    assert np.array_equal(d_out.cpu().numpy(), want)
    assert np.array_equal(d_dec.cpu().numpy(), O.decompress(fmt, want, w, h))
    assert np.array_equal(T.compress_device(fmt, d_in, w, h, tp).cpu().numpy(), want)          # torch's current stream
    with pytest.raises(TypeError):
        T.compress_device(fmt, torch.from_numpy(img.reshape(-1)), w, h, tp)                      # host tensor
    with pytest.raises(ValueError):
        T.compress_device(fmt, d_in[:100], w, h, tp)                                             # too short
    with pytest.raises(ValueError):
        T.compress_device(fmt, d_in, w, h, tp, output=torch.empty(8, dtype=torch.uint8, device="cuda"))
