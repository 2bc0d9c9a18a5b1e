"""Builds texpresso_b200/libtexpresso_b200.so (CUDA kernels + C ABI) for sm_100a with nvcc, in-tree.

Flags that are part of the numeric contract: no FMA contraction, IEEE division
and square root, no flush-to-zero."""
import os, pathlib, shutil, subprocess, sys

HERE = pathlib.Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = pathlib.Path(os.environ.get("TXP_BUILD_OUT", HERE / "libtexpresso_b200.so"))
SOURCES = [CSRC / "txp_api.cu"]
HEADERS = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [HERE.parent / "include" / "texpresso_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2", "-cudart", "static",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and pathlib.Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not OUT.exists():
        return False
    t = OUT.stat().st_mtime
    return all(p.stat().st_mtime <= t for p in SOURCES + HEADERS + [pathlib.Path(__file__)])


def build(force=False, verbose=False, extra=()):
    if not force and up_to_date():
        return OUT
    extra = list(extra) + os.environ.get("TXP_BUILD_DEFS", "").split()
    cmd = [nvcc_path(), *NVCC_FLAGS, *extra, "-o", str(OUT), *map(str, SOURCES)]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, extra=(["-Xptxas", "-v"] if "-v" in sys.argv else []))
    print("built", OUT)
