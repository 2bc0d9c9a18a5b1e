"""Build-time contracts of the compiled library, checked without a GPU (cuobjdump on the in-tree .so):
every kernel family of DESIGN.md is present for every format it serves, the ClusterFit search kernels keep the register /
shared-memory budgets their occupancy is planned with (no spills in the search loops), and the numeric contract's compile
flags are the ones the build script passes."""
import re, shutil, subprocess
import pytest

from texpresso_b200 import _lib, build


def _resources():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([exe, "--dump-resource-usage", str(_lib.SO_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    res, name = {}, None
    for line in out.stdout.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and name:
            res[name] = tuple(int(x) for x in m.groups())
    return res


def test_kernel_families_present():
    names = "\n".join(_resources())
    for family, count in (("cluster_setup_kernel", 3), ("cluster_setup_sorted_kernel", 3), ("cluster_lane_kernel", 3),
                          ("cluster_lane_iter_kernel", 4), ("colour_search_kernel", 3), ("colour_encode_kernel", 3),
                          ("range_encode_kernel", 3), ("alpha_lattice_image_kernel", 2), ("alpha_lattice_tma_kernel", 2), ("alpha_lattice_kernel", 2),
                          ("decode_kernel", 5), ("mip_downsample_kernel", 1), ("expand_pixels_kernel", 4)):
        found = len(re.findall(r"\d+%s[IE]" % family, names))
        assert found >= count, (family, found, count)


def test_search_kernel_budgets():
    res = _resources()
    for name, (reg, stack, shared, local) in res.items():
        if "cluster_lane_kernel" in name:
            # 6 CTAs of 128 threads per SM: <= 85 registers, <= 37 KB of shared memory each; nothing on the stack
            assert reg <= 85 and shared <= 37 * 1024 and stack == 0 and local == 0, (name, reg, stack, shared, local)
        if "cluster_lane_iter_kernel" in name:
            # 5 CTAs per SM: <= 102 registers, <= 45 KB
            assert reg <= 102 and shared <= 45 * 1024 and stack == 0 and local == 0, (name, reg, stack, shared, local)
        if "range_encode_kernel" in name or "cluster_setup" in name:
            # rolled front end: 5-6 CTAs of 128 threads per SM (<= 102 registers, <= 38 KB), at most a few spilled words
            assert reg <= 102 and shared <= 38 * 1024 and stack <= 32, (name, reg, stack, shared, local)
        if "colour_search_kernel" in name:
            assert reg <= 64, (name, reg)                      # 8 CTAs x 4 warps per SM


def test_tma_staging_in_sass():
    """the BC4 / BC5 image kernels stage their strips with the tensor memory accelerator (UTMALDG = cp.async.bulk.tensor)"""
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([exe, "-sass", str(_lib.SO_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert out.stdout.count("UTMALDG") >= 2 and "SYNCS.PHASECHK" in out.stdout


def test_numeric_contract_flags():
    flags = " ".join(build.NVCC_FLAGS)
    for f in ("-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "arch=compute_100a,code=sm_100a", "-lineinfo"):
        assert f in flags, f
