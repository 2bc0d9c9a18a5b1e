#!/usr/bin/env python3
"""A/B of the round-aligned pipeline chunk plan (txp_debug_set key 5: smallest shard in lane rounds that takes it, 0 = off) for
Format.compress on pinned host buffers: block-row shards of the 8192-wide texture (256 rows = one rank of 8, 512 = one of 4) and
square medium images, noise and `smooth`, BC1 / BC3 ClusterFit.  Interleaved, median of REPS; outputs must agree byte for byte."""
import json, sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import texpresso_b200 as T
from texpresso_b200 import synth, _lib
L = _lib.load(); T.set_device(0)
P = T.COLOUR_WEIGHTS_PERCEPTUAL
REPS = 9
KNOBS = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 4, 2]

def case(name, w, h, kind, fmt):
    img = synth.generate(kind, w, h, 3)
    if fmt == 0 and kind.startswith("noise"):
        img = img.copy(); img[..., 3] = 255
    hin = torch.from_numpy(np.ascontiguousarray(img).reshape(-1)).pin_memory()
    F = T.Format(fmt)
    outs = {k: torch.empty(F.compressed_size(w, h), dtype=torch.uint8).pin_memory() for k in KNOBS}
    prm = T.Params(T.Algorithm.ClusterFit, P, False)
    ts = {k: [] for k in KNOBS}
    for i in range(REPS + 2):
        for k in KNOBS:
            L.txp_debug_set(5, k)
            t0 = time.perf_counter()
            F.compress(hin.numpy(), w, h, prm, output=outs[k].numpy())
            if i >= 2:
                ts[k].append(1e3 * (time.perf_counter() - t0))
    L.txp_debug_set(5, 4)
    same = all(torch.equal(outs[KNOBS[0]], outs[k]) for k in KNOBS)
    rec = {"case": name, "fmt": "bc1" if fmt == 0 else "bc3", "input": kind, "w": w, "h": h, "same": same}
    for k in KNOBS:
        rec[f"ms_wave{k}"] = round(sorted(ts[k])[len(ts[k]) // 2], 3)
    print(json.dumps(rec), flush=True)

for rows in (256, 512, 384, 192):
    for kind in ("noise_alpha", "smooth"):
        for fmt in (2, 0):
            case(f"shard_{rows}rows", 8192, 4 * rows, kind, fmt)
for side in (2048, 3072, 4096):
    for kind in ("noise_alpha", "smooth"):
        for fmt in (2, 0):
            case(f"square_{side}", side, side, kind, fmt)
