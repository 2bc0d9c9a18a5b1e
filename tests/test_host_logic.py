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
