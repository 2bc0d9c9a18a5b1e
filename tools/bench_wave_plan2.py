#!/usr/bin/env python3
"""Round-aligned chunk plan for LARGE shards (txp_debug_set key 6 = largest shard in rounds that takes it): 512 / 1024 / 2048 block rows of
the 8192-wide texture (one rank of 4 / of 2 / the whole texture), BC1 / BC3 ClusterFit, pinned host buffers, interleaved, median."""
import json, sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import texpresso_b200 as T
from texpresso_b200 import synth, _lib
L = _lib.load(); T.set_device(0)
P = T.COLOUR_WEIGHTS_PERCEPTUAL
REPS = 7
KNOBS = [(8, 2), (64, 2), (64, 3), (64, 4)]
for rows in (512, 1024, 2048):
    for kind in ("noise_alpha", "smooth"):
        for fmt in (2, 0):
            w, h = 8192, 4 * rows
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
                    L.txp_debug_set(6, k[0]); L.txp_debug_set(7, k[1])
                    t0 = time.perf_counter()
                    F.compress(hin.numpy(), w, h, prm, output=outs[k].numpy())
                    if i >= 2:
                        ts[k].append(1e3 * (time.perf_counter() - t0))
            L.txp_debug_set(6, 40); L.txp_debug_set(7, 2)
            rec = {"rows": rows, "fmt": "bc1" if fmt == 0 else "bc3", "input": kind, "same": all(bool(torch.equal(outs[KNOBS[0]], outs[k])) for k in KNOBS)}
            for k in KNOBS:
                rec[f"ms_max{k[0]}_c{k[1]}"] = round(sorted(ts[k])[len(ts[k]) // 2], 3)
            print(json.dumps(rec), flush=True)
