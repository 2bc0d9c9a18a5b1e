"""Builds tests/cpp/test_reference_units.cpp (the reference's unit tests, lib.rs:345-505, restated against the C++
mirror include/texpresso.hpp) with g++ and, on a GPU box, runs it against libtexpresso_b200.so."""
import pathlib, subprocess
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
BIN = ROOT / "tests" / "cpp" / "_build" / "test_reference_units"


def _build():
    from texpresso_b200 import build as B
    so = B.build()
    BIN.parent.mkdir(exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(ROOT / "tests/cpp/test_reference_units.cpp"),
                    "-o", str(BIN), f"-L{so.parent}", "-ltexpresso_b200", f"-Wl,-rpath,{so.parent}"], check=True)


def test_cpp_mirror_compiles_and_links():
    _build()
    assert BIN.exists()


@pytest.mark.gpu
def test_reference_unit_tests_through_cpp_mirror():
    _build()
    r = subprocess.run([str(BIN)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all reference unit tests passed" in r.stdout
