#!/usr/bin/env python3
"""Texture batches end to end from pinned host memory (txp_compress_batch, no mips): 256 textures 1024^2 `smooth`."""
import json, pathlib, sys, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch, texpresso_b200 as T
from texpresso_b200 import synth
T.set_device(0)
n = 256
pinned = [torch.from_numpy(synth.generate("smooth", 1024, 1024, 6_000_000 + t).reshape(-1)).pin_memory() for t in range(n)]
texs = [(p.numpy(), 1024, 1024) for p in pinned]
for fmt, prm, name in ((T.Format.Bc3, T.Params(), "bc3_clusterfit"), (T.Format.Bc1, T.Params(), "bc1_clusterfit"),
                       (T.Format.Bc1, T.Params(T.Algorithm.RangeFit), "bc1_rangefit"), (T.Format.Bc4, T.Params(), "bc4")):
    outs_t = [torch.empty(fmt.compressed_size(1024, 1024), dtype=torch.uint8).pin_memory() for _ in range(n)]
    outs = [o.numpy() for o in outs_t]
    T.compress_batch(fmt, texs, prm, n_gpus=1, outputs=outs)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); T.compress_batch(fmt, texs, prm, n_gpus=1, outputs=outs); ts.append(time.perf_counter() - t0)
    dt = min(ts)
    print(json.dumps({"case": name + "_batch_e2e", "textures": n, "size": "1024x1024", "ms": round(dt * 1e3, 2), "textures_per_s": round(n / dt),
                      "mpix_s": round(n * 1.048576 / dt), "h2d_gb_s": round(n * 4.194304e-3 / dt, 1)}), flush=True)
