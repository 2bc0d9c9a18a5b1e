#!/usr/bin/env python3
"""ClusterFit kernel structure vs launch size: warp-per-block search against lane-per-block search on square images
from 128^2 to 4096^2 (device-resident, CUDA events, L2 flushed, median of 7).  Decides TXP_LANE_MIN_BLOCKS."""
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
for kind in ("noise_alpha", "smooth"):
    for fmt, bs in ((2, 16), (0, 8)):
        for side in (128, 256, 384, 512, 768, 1024, 1536, 2048, 4096):
            img = synth.generate(kind, side, side, 9)
            if fmt == 0 and kind == "noise_alpha": img[..., 3] = 255
            d = torch.from_numpy(img.reshape(-1)).cuda()
            out = torch.empty((side // 4) ** 2 * bs, dtype=torch.uint8, device="cuda")
            res = {}
            for name, v in (("warp", 2), ("lane", 3)):
                ts = []
                for rep in range(8):
                    L.txp_debug_set(0, v)
                    flush.fill_(rep)
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    rc = L.txp_compress_device(fmt, ctypes.c_void_p(d.data_ptr()), side, side, ctypes.byref(cp), ctypes.c_void_p(out.data_ptr()), out.numel(), stream)
                    assert rc == 0
                    b.record(); torch.cuda.synchronize()
                    if rep: ts.append(a.elapsed_time(b))
                res[name] = round(statistics.median(ts), 4)
            L.txp_debug_set(0, 0)
            print(json.dumps({"input": kind, "fmt": "bc3" if fmt == 2 else "bc1", "side": side, "blocks": (side // 4) ** 2, **res}), flush=True)
