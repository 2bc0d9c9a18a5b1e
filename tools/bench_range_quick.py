#!/usr/bin/env python3
"""RangeFit device-resident timing (BC1 / BC3, 8192^2 noise and smooth), CUDA events, L2 flushed; compares libraries given as name=path."""
import ctypes, json, pathlib, statistics, sys
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import texpresso_b200 as T
from texpresso_b200 import synth, _lib
libs = []
for a in sys.argv[1:] or ["default=" + str(ROOT / "texpresso_b200/libtexpresso_b200.so")]:
    name, path = a.split("=")
    L = ctypes.CDLL(str(pathlib.Path(path).resolve()))
    L.txp_compress_device.restype = ctypes.c_int
    L.txp_compress_device.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.POINTER(_lib.CParams), ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    libs.append((name, L))
torch.cuda.set_device(0)
w = h = 8192
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
cp = T.Params(T.Algorithm.RangeFit, T.COLOUR_WEIGHTS_PERCEPTUAL, False)._c()
for kind in ("noise_opaque", "noise_alpha", "smooth"):
    d = torch.from_numpy(synth.generate(kind, w, h, 3).reshape(-1)).cuda()
    for fmt, bs in ((0, 8), (2, 16)):
        outs = {n: torch.zeros((w // 4) * (h // 4) * bs, dtype=torch.uint8, device="cuda") for n, _ in libs}
        ts = {n: [] for n, _ in libs}
        for rep in range(9):
            for n, L in libs:
                flush.fill_(rep)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                assert L.txp_compress_device(fmt, ctypes.c_void_p(d.data_ptr()), w, h, ctypes.byref(cp), ctypes.c_void_p(outs[n].data_ptr()), outs[n].numel(), stream) == 0
                b.record(); torch.cuda.synchronize()
                if rep >= 2: ts[n].append(a.elapsed_time(b))
        first = libs[0][0]
        print(json.dumps({"input": kind, "fmt": "bc1" if fmt == 0 else "bc3", "same": all(bool(torch.equal(outs[first], outs[n])) for n, _ in libs),
                          **{n: round(statistics.median(v), 4) for n, v in ts.items()}}), flush=True)
