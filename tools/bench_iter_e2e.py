#!/usr/bin/env python3
"""End-to-end (pinned host -> pinned host) timing of Format.compress for the ClusterFit family under different pipeline chunk plans
(txp_debug_set key 3 = chunk size override in MiB, key 4 = geometric growth of small shards), next to the device-resident time."""
import ctypes, json, sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import texpresso_b200 as T
from texpresso_b200 import synth, _lib
L = _lib.load(); T.set_device(0)
P = T.COLOUR_WEIGHTS_PERCEPTUAL
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(fmt, alg, rows, kind, chunk_mib=0, growth=3, reps=4):
    w = 8192; h = 4 * rows
    img = synth.generate(kind, w, 8192, 3, y0=0, y1=h)
    hin = torch.from_numpy(img.reshape(-1)).pin_memory()
    bs = 8 if fmt == 0 else 16
    hout = torch.empty(rows * 2048 * bs, dtype=torch.uint8).pin_memory()
    prm = T.Params(T.Algorithm(alg), P, False)
    L.txp_debug_set(3, chunk_mib); L.txp_debug_set(4, growth)
    ts = []
    for i in range(reps + 1):
        flush.fill_(i); torch.cuda.synchronize()
        t0 = time.perf_counter()
        T.Format(fmt).compress(hin.numpy(), w, h, prm, output=hout.numpy())
        ts.append(1e3 * (time.perf_counter() - t0))
    L.txp_debug_set(3, 0); L.txp_debug_set(4, 3)
    return min(ts[1:]), sum(ts[1:]) / reps


def dev(fmt, alg, rows, kind, reps=4):
    w = 8192; h = 4 * rows
    d = torch.from_numpy(synth.generate(kind, w, 8192, 3, y0=0, y1=h).reshape(-1)).cuda()
    bs = 8 if fmt == 0 else 16
    out = torch.empty(rows * 2048 * bs, dtype=torch.uint8, device="cuda")
    cp = T.Params(T.Algorithm(alg), P, False)._c()
    ts = []
    for i in range(reps + 1):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(L.txp_compress_device(fmt, ctypes.c_void_p(d.data_ptr()), w, h, ctypes.byref(cp), ctypes.c_void_p(out.data_ptr()), out.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts[1:])


QUICK = len(sys.argv) > 1
# BASELINE config 3 at N = 1: BC1 IterativeClusterFit, full 8192^2
d = dev(0, 2, 2048, "noise_opaque", reps=2)
print(json.dumps({"case": "bc1_iterative_8192_device", "ms": d}), flush=True)
for mib in ((0,) if QUICK else (0, 32, 64, 96, 128, 192)):
    best, mean = run(0, 2, 2048, "noise_opaque", chunk_mib=mib, reps=2)
    print(json.dumps({"case": "bc1_iterative_8192_e2e", "chunk_mib": mib or "auto", "ms_best": best, "ms_mean": mean, "of_device": d / best}), flush=True)
# one rank's shard at N = 8 (256 block rows), BC3 and BC1 ClusterFit
for fmt, kind in ((2, "noise_alpha"), (0, "noise_opaque")):
    d = dev(fmt, 1, 256, kind)
    print(json.dumps({"case": f"fmt{fmt}_cluster_256rows_device", "ms": d}), flush=True)
    for growth, mib in (((3, 0),) if QUICK else ((0, 0), (3, 0), (4, 0), (2, 0), (0, 8), (0, 4), (0, 32))):
        best, mean = run(fmt, 1, 256, kind, chunk_mib=mib, growth=growth, reps=6)
        print(json.dumps({"case": f"fmt{fmt}_cluster_256rows_e2e", "growth": growth, "chunk_mib": mib or "auto", "ms_best": best, "ms_mean": mean, "of_device": d / best}), flush=True)
# N = 1 metric workload, BC3 ClusterFit
d = dev(2, 1, 2048, "noise_alpha", reps=3)
print(json.dumps({"case": "bc3_cluster_8192_device", "ms": d}), flush=True)
for mib in ((0,) if QUICK else (0, 16, 32, 64)):
    best, mean = run(2, 1, 2048, "noise_alpha", chunk_mib=mib, reps=3)
    print(json.dumps({"case": "bc3_cluster_8192_e2e", "chunk_mib": mib or "auto", "ms_best": best, "ms_mean": mean, "of_device": d / best}), flush=True)
