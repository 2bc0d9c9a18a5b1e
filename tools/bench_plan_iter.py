#!/usr/bin/env python3
"""Explicit pipeline chunk plans (TXP_PLAN, experiments only) for BC1 IterativeClusterFit over the whole 8192^2 texture (BASELINE config 3)
through Format.compress on pinned buffers, next to the device-resident launch."""
import json, os, sys, time, pathlib, subprocess, ctypes
ROOT = pathlib.Path(__file__).resolve().parent.parent
PLANS = {
    "auto": "",
    "64_992x2": "64,992x2;L=1",
    "128_1920": "128,1920;L=1",
    "192_1856": "192,1856;L=1",
    "256_1792": "256,1792;L=1",
    "256_1792_lane": "256,1792;L=0",
    "384_1664_lane": "384,1664;L=0",
    "32_224_1792": "32,224,1792;L=2",
}
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, str(ROOT))
    import numpy as np, torch
    import texpresso_b200 as T
    from texpresso_b200 import synth, _lib
    T.set_device(0)
    L = _lib.load()
    img = synth.generate("noise_opaque", 8192, 8192, 3)
    hin = torch.from_numpy(img.reshape(-1)).pin_memory()
    F = T.Format.Bc1
    out = torch.empty(F.compressed_size(8192, 8192), dtype=torch.uint8).pin_memory()
    prm = T.Params(T.Algorithm.IterativeClusterFit, T.COLOUR_WEIGHTS_PERCEPTUAL, False)
    ts = []
    for i in range(6):
        t0 = time.perf_counter(); F.compress(hin.numpy(), 8192, 8192, prm, output=out.numpy()); ts.append(1e3 * (time.perf_counter() - t0))
    res = {"plan": os.environ.get("TXP_PLAN", "auto"), "e2e_ms": [round(min(ts[1:]), 2), round(sorted(ts[1:])[2], 2)]}
    if not os.environ.get("TXP_PLAN"):
        d = hin.cuda(); o = torch.empty(out.numel(), dtype=torch.uint8, device="cuda"); cp = prm._c(); dt = []
        for i in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(L.txp_compress_device(0, ctypes.c_void_p(d.data_ptr()), 8192, 8192, ctypes.byref(cp), ctypes.c_void_p(o.data_ptr()), o.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
            b.record(); torch.cuda.synchronize(); dt.append(a.elapsed_time(b))
        res["device_ms"] = round(min(dt[1:]), 2)
        res["same"] = bool(torch.equal(o.cpu(), out))
    print(json.dumps(res), flush=True)
else:
    for name, plan in PLANS.items():
        env = dict(os.environ)
        if plan: env["TXP_PLAN"] = plan
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        lines = out.stdout.strip().splitlines()
        print(name, lines[-1] if lines else "no output: " + out.stderr[-300:], flush=True)
