#!/usr/bin/env python3
"""Interleaved A/B timing of ClusterFit kernel variants in ONE process (cancels clock / thermal drift).
usage: ab_test.py name=lib.so[:auto|fused|warp|lane] ...   (each variant = a built library + a ClusterFit kernel structure)"""
import ctypes, json, pathlib, statistics, subprocess, sys, threading
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import texpresso_b200 as T
from texpresso_b200 import synth, _lib

variants = []
only = None
args = sys.argv[1:]
if args and args[0].startswith("--cases="):
    only = args.pop(0).split("=", 1)[1].split(",")
for a in args:
    name, spec = a.split("=")
    path, _, mode = spec.partition(":")
    L = ctypes.CDLL(str(pathlib.Path(path).resolve()))
    L.txp_compress_device.restype = ctypes.c_int
    L.txp_compress_device.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.POINTER(_lib.CParams), ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    L.txp_debug_set.argtypes = [ctypes.c_int, ctypes.c_int]
    variants.append((name, L, {"": 0, "auto": 0, "fused": 1, "warp": 2, "lane": 3}[mode]))

torch.cuda.set_device(0)
w = h = 8192
img = synth.generate("noise_alpha", w, h, 3)
d3 = torch.from_numpy(img.reshape(-1)).cuda(); img[..., 3] = 255
d1 = torch.from_numpy(img.reshape(-1)).cuda()
sm = torch.from_numpy(synth.generate("smooth", w, h, 5).reshape(-1)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
cases = [("bc3", 2, d3, 16, T.Params()), ("bc1", 0, d1, 8, T.Params()),
         ("bc1_iter", 0, d1, 8, T.Params(T.Algorithm.IterativeClusterFit)), ("bc3_smooth", 2, sm, 16, T.Params()),
         ("bc3_smooth_iter", 2, sm, 16, T.Params(T.Algorithm.IterativeClusterFit)), ("bc1_smooth", 0, sm, 8, T.Params())]
if only:
    cases = [c for c in cases if c[0] in only]
clk = []
p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [clk.append(l) for l in p.stdout], daemon=True).start()
res = {}
for cname, fmt, d, bs, prm in cases:
    out = torch.empty((w // 4) * (h // 4) * bs, dtype=torch.uint8, device="cuda")
    cp = prm._c()
    times = {n: [] for n, _, _ in variants}
    for rep in range(6):
        for n, L, mode in variants:
            L.txp_debug_set(0, mode)            # per call: two variants may share one library (same globals)
            flush.fill_(rep)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = L.txp_compress_device(fmt, ctypes.c_void_p(d.data_ptr()), w, h, ctypes.byref(cp), ctypes.c_void_p(out.data_ptr()), out.numel(), stream)
            assert rc == 0
            b.record(); torch.cuda.synchronize()
            if rep >= 1: times[n].append(a.elapsed_time(b))
    res[cname] = {n: round(statistics.median(v), 3) for n, v in times.items()}
p.terminate()
sm_clk = [float(l.split(",")[0]) for l in clk if "," in l]
print(json.dumps(res))
print("sm clock MHz: median %.0f min %.0f max %.0f (%d samples)" % (statistics.median(sm_clk), min(sm_clk), max(sm_clk), len(sm_clk)))
