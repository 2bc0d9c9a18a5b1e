"""Randomised sweep of the public entry points against the oracle (tests/fuzz_api.py): random sizes (1 x 1 ... 384 x 131, and a few that
reach the lane-per-block kernels, texture groups, the TMA-staged kernel and several pipeline chunks), formats, algorithms, weights, alpha
weighting, over-long outputs, batches, mip chains, the compact-pixel entry point; encode and decode.  Bit-exact or it fails."""
import pathlib, sys
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent))


@pytest.mark.parametrize("seed", [101, 102])
def test_fuzz_small_images(seed):
    import fuzz_api
    bad = fuzz_api.run(seed, 300, large=False)
    assert not bad, bad[:10]


def test_fuzz_large_images():
    import fuzz_api
    bad = fuzz_api.run(103, 16, large=True)
    assert not bad, bad[:10]
