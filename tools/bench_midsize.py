#!/usr/bin/env python3
"""End-to-end Format.compress (pinned host buffers) on medium images, ClusterFit: which kernel structure should the chunks take?"""
import json, os, pathlib, sys, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch, texpresso_b200 as T
from texpresso_b200 import synth
T.set_device(0)
for kind in ("noise_alpha", "smooth"):
    for side in (1024, 1536, 2048, 3072, 4096):
        img = torch.from_numpy(synth.generate(kind, side, side, 4).reshape(-1)).pin_memory()
        res = {}
        for name, fmt in (("bc3", T.Format.Bc3), ("bc1", T.Format.Bc1)):
            out = torch.empty(fmt.compressed_size(side, side), dtype=torch.uint8).pin_memory()
            for _ in range(3): fmt.compress(img.numpy(), side, side, T.Params(), output=out.numpy())
            ts = []
            for _ in range(9):
                t0 = time.perf_counter(); fmt.compress(img.numpy(), side, side, T.Params(), output=out.numpy()); ts.append(time.perf_counter() - t0)
            res[name + "_ms"] = round(sorted(ts)[len(ts) // 2] * 1e3, 3)
        print(json.dumps({"lib": os.environ.get("TEXPRESSO_B200_LIB", "default"), "input": kind, "side": side, **res}), flush=True)
