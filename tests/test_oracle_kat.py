"""Pins the CPU oracle to every known-answer vector the reference's own tests hold
(reference lib/src/lib.rs:350-504, lib/src/test_data.rs), via tests/golden/kat.json."""
import json, pathlib
import numpy as np
import pytest
from tests import oracle_lib as O

KAT = json.loads((pathlib.Path(__file__).parent / "golden" / "kat.json").read_text())
SETS = sorted(KAT["sets"].items())


@pytest.mark.parametrize("fmt,w,h,size", KAT["sizes"])
def test_storage_requirements(fmt, w, h, size):                 # lib.rs:350-361
    assert O.compressed_size(fmt, w, h) == size


@pytest.mark.parametrize("name,ds", SETS)
def test_decompression(name, ds):                               # lib.rs:363-367
    out = O.decompress(ds["format"], np.array(ds["encoded"], np.uint8), 4, 4)
    assert out.tolist() == ds["decoded"]


@pytest.mark.parametrize("alg", [O.CLUSTER_FIT, O.RANGE_FIT, O.ITERATIVE_CLUSTER_FIT])
@pytest.mark.parametrize("name,ds", SETS)
def test_compression(name, ds, alg):                            # lib.rs:369-393: all 3 algorithms, uniform weights
    p = O.make_params(alg, O.UNIFORM, False)
    out = O.compress(ds["format"], np.array(ds["decoded"], np.uint8), 4, 4, p)
    assert bytes(out).hex() == bytes(ds["encoded"]).hex()


def test_bc1_decompression_height_not_multiple_of_4():          # lib.rs:415-444
    d = KAT["decode_4x6"]
    out = O.decompress(d["format"], np.array(d["encoded"], np.uint8), d["width"], d["height"])
    assert out.reshape(-1, 4).tolist() == [d["pixel"]] * (d["width"] * d["height"])
