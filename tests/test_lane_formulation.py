"""CPU check of the FORMULATION behind the lane-per-block ClusterFit kernels (tests/restatements/lane_form.c): hoisted
(i, j) terms, unconditional running sums with a zero guard entry, (error, key) winner with recomputed endpoints, one-pass
index assignment -- restated in C inside the oracle's own ClusterFit state machine and compared with the oracle's literal
loop nest, byte for byte, on the stratified block corpus and on random blocks.  No GPU involved; the GPU kernels themselves
are compared with the oracle in tests/test_gpu_cluster_lane.py."""
import ctypes, pathlib, subprocess
import numpy as np
import pytest

from tests import oracle_lib as O
from tests import blockgen

HERE = pathlib.Path(__file__).resolve().parent
SRC = HERE / "restatements" / "lane_form.c"
OUT = HERE / "restatements" / "_build" / "liblane_form.so"


@pytest.fixture(scope="module")
def L():
    OUT.parent.mkdir(exist_ok=True)
    subprocess.run(["gcc", "-O2", "-std=c11", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-shared", "-o", str(OUT), str(SRC), "-lm", "-lpthread"], check=True)
    lib = ctypes.CDLL(str(OUT))
    lib.txl_compare.restype = ctypes.c_size_t
    lib.txl_compare.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(O.Params), ctypes.POINTER(ctypes.c_size_t)]
    return lib


def _compare(L, fmt, blocks, masks, alg, weights, awa):
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8)
    masks = np.ascontiguousarray(masks, dtype=np.uint32)
    p = O.make_params(alg, weights, awa)
    searched = ctypes.c_size_t(0)
    bad = L.txl_compare(fmt, blocks.ctypes.data, masks.ctypes.data, len(masks), ctypes.byref(p), ctypes.byref(searched))
    return bad, searched.value


@pytest.mark.parametrize("awa", [False, True])
@pytest.mark.parametrize("alg", [1, 2])
@pytest.mark.parametrize("fmt", [0, 2])
def test_lane_formulation_on_stratified_blocks(L, fmt, alg, awa):
    blocks, masks, _tags = blockgen.colour_cases()
    bad, searched = _compare(L, fmt, blocks, masks, alg, O.PERCEPTUAL, awa)
    assert searched > 1000 and bad == 0, (bad, searched)


@pytest.mark.parametrize("alg,weights", [(1, O.PERCEPTUAL), (1, (0.3, 1.7, 0.05)), (2, O.UNIFORM)])
def test_lane_formulation_on_random_blocks(L, alg, weights):
    rng = np.random.default_rng(99 + alg)
    n = 3000
    blocks = rng.integers(0, 256, size=(n, 16, 4), dtype=np.uint8)
    blocks[: n // 3, :, :3] &= 0xF0                         # few distinct colours, many exact ties
    blocks[n // 3: n // 2, :, 3] = 255
    masks = np.full(n, 0xFFFF, np.uint32)
    masks[::7] = rng.integers(1, 1 << 16, size=len(masks[::7]))
    for fmt in (0, 2):
        bad, searched = _compare(L, fmt, blocks, masks, alg, weights, False)
        assert searched > 2500 and bad == 0, (fmt, bad, searched)
