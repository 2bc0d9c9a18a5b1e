#!/usr/bin/env python3
"""Quick kernel-only timing of BC1/BC3 ClusterFit on 8192^2 noise (for A/B tests of kernel variants).
TEXPRESSO_B200_LIB selects the library."""
import ctypes, sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch, texpresso_b200 as T
from texpresso_b200 import synth, _lib
L = _lib.load(); T.set_device(0)
w = h = 8192
img = synth.generate("noise_alpha", w, h, 3)
d3 = torch.from_numpy(img.reshape(-1)).cuda()
img[..., 3] = 255
d1 = torch.from_numpy(img.reshape(-1)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {}
for name, fmt, bs, d in (("bc1", T.Format.Bc1, 8, d1), ("bc3", T.Format.Bc3, 16, d3)):
    out = torch.empty((w // 4) * (h // 4) * bs, dtype=torch.uint8, device="cuda")
    cp = T.Params()._c()
    ts = []
    for i in range(5):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(L.txp_compress_device(int(fmt), ctypes.c_void_p(d.data_ptr()), w, h, ctypes.byref(cp), ctypes.c_void_p(out.data_ptr()), out.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sum(ts[2:]) / 3
    res[name] = {"ms": round(ms, 3), "mpix_s": round(w * h / ms / 1e3, 1), "crc": int(out.to(torch.int64).sum().item())}
# iterative (BC1, noise) and smooth content (BC3 cluster)
def timeit(fmt, d, bs, params, reps=3):
    out = torch.empty((w // 4) * (h // 4) * bs, dtype=torch.uint8, device="cuda")
    cp = params._c(); ts = []
    for i in range(reps + 1):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(L.txp_compress_device(int(fmt), ctypes.c_void_p(d.data_ptr()), w, h, ctypes.byref(cp), ctypes.c_void_p(out.data_ptr()), out.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return round(sum(ts[1:]) / reps, 3)
res["bc1_iter_ms"] = timeit(T.Format.Bc1, d1, 8, T.Params(T.Algorithm.IterativeClusterFit))
sm = torch.from_numpy(synth.generate("smooth", w, h, 5).reshape(-1)).cuda()
res["bc3_smooth_ms"] = timeit(T.Format.Bc3, sm, 16, T.Params())
res["bc3_smooth_iter_ms"] = timeit(T.Format.Bc3, sm, 16, T.Params(T.Algorithm.IterativeClusterFit))
print(json.dumps(res))
