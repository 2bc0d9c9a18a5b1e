#!/usr/bin/env python3
"""ClusterFit kernel structure vs launch size: warp-per-block search, lane-per-block search and the hybrid launch (full lane rounds
+ warp-per-block tail) on 4096-wide images whose block count is a chosen multiple of one lane round (113 664 blocks on 148 SMs).
Device-resident, CUDA events, L2 flushed, median of 7.  Decides TXP_LANE_MIN_BLOCKS and TXP_TAIL_FRAC."""
import ctypes, json, pathlib, statistics, sys
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import texpresso_b200 as T
from texpresso_b200 import synth, _lib
L = _lib.load(); T.set_device(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
cp = T.Params()._c()
W = 4096
WAVES = [float(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "0.3,0.6,0.9,1.0,1.15,1.3,1.6,1.9,2.0,2.3,2.6,3.0,3.3,4.6,9.2".split(","))]
L.txp_debug_set(2, 100)                                    # hybrid variant: every partial last round goes to the warp kernel
for kind in ("noise_alpha", "smooth"):
    for fmt, bs in ((2, 16), (0, 8)):
        for x in WAVES:
            rows = max(1, round(x * 111))
            h = 4 * rows
            img = synth.generate(kind, W, h, 9)
            if fmt == 0 and kind == "noise_alpha": img[..., 3] = 255
            d = torch.from_numpy(img.reshape(-1)).cuda()
            out = torch.empty(rows * (W // 4) * bs, dtype=torch.uint8, device="cuda")
            res, outs = {}, {}
            for name, v in (("warp", 2), ("lane", 3), ("hybrid", 4)):
                ts = []
                for rep in range(8):
                    L.txp_debug_set(0, v)
                    flush.fill_(rep)
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    rc = L.txp_compress_device(fmt, ctypes.c_void_p(d.data_ptr()), W, h, ctypes.byref(cp), ctypes.c_void_p(out.data_ptr()), out.numel(), stream)
                    assert rc == 0
                    b.record(); torch.cuda.synchronize()
                    if rep: ts.append(a.elapsed_time(b))
                res[name] = round(statistics.median(ts), 4)
                outs[name] = out.clone()
            L.txp_debug_set(0, 0)
            same = bool(torch.equal(outs["warp"], outs["lane"]) and torch.equal(outs["warp"], outs["hybrid"]))
            print(json.dumps({"input": kind, "fmt": "bc3" if fmt == 2 else "bc1", "rounds": x, "blocks": rows * (W // 4), **res, "identical": same}), flush=True)
L.txp_debug_set(2, 70)
