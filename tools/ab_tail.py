#!/usr/bin/env python3
"""How much does a partly filled round of the lane-per-block search cost, and does more ILP per lane (k loop unrolled by 2, fewer CTAs
per SM) make it cheaper?  Forced lane-per-block launches (txp_debug_set(0, 3)) of 8192-wide strips, device-resident, interleaved.
usage: ab_tail.py name=lib.so ..."""
import ctypes, json, pathlib, statistics, sys
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import texpresso_b200 as T
from texpresso_b200 import synth, _lib

variants = []
for a in sys.argv[1:]:
    name, path = a.split("=")
    L = ctypes.CDLL(str(pathlib.Path(path).resolve()))
    L.txp_compress_device.restype = ctypes.c_int
    L.txp_compress_device.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.POINTER(_lib.CParams), ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    L.txp_debug_set.argtypes = [ctypes.c_int, ctypes.c_int]
    L.txp_debug_set(0, 3)
    variants.append((name, L))
torch.cuda.set_device(0)
w = 8192
img = synth.generate("noise_alpha", w, 1024, 3)
d3 = torch.from_numpy(img.reshape(-1)).cuda(); img[..., 3] = 255
d1 = torch.from_numpy(img.reshape(-1)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
cp = T.Params()._c()
for cname, fmt, d, bs in (("bc3", 2, d3, 16), ("bc1", 0, d1, 8)):
    for rows in (14, 20, 28, 34, 42, 55, 83, 111, 145, 256):          # 55.5 rows = one round of 113 664 lanes
        h = 4 * rows
        outs = {n: torch.zeros((w // 4) * rows * bs, dtype=torch.uint8, device="cuda") for n, _ in variants}
        times = {n: [] for n, _ in variants}
        for rep in range(8):
            for n, L in variants:
                flush.fill_(rep)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                rc = L.txp_compress_device(fmt, ctypes.c_void_p(d.data_ptr()), w, h, ctypes.byref(cp), ctypes.c_void_p(outs[n].data_ptr()), outs[n].numel(), stream)
                assert rc == 0
                b.record(); torch.cuda.synchronize()
                if rep >= 2: times[n].append(a.elapsed_time(b))
        first = variants[0][0]
        rec = {"fmt": cname, "rows": rows, "rounds": round(rows * 2048 / 113664, 2), "same": all(bool(torch.equal(outs[first], outs[n])) for n, _ in variants)}
        rec.update({n: round(statistics.median(v), 4) for n, v in times.items()})
        print(json.dumps(rec), flush=True)
