#!/usr/bin/env python3
"""Explicit chunk plans (TXP_PLAN, experiments only) for one rank's shard of the metric texture at 8 ranks (8192 x 1024), BC3 + BC1 ClusterFit."""
import json, os, sys, time, pathlib, subprocess
ROOT = pathlib.Path(__file__).resolve().parent.parent
PLANS = {
    "auto(8,27,111,55,55)": "",
    "8,27,55,55,111": "8,27,55,55,111;L=2",
    "8,28,55x4": "8,28,55x4;L=2",
    "8,27,55,111,55": "8,27,55,111,55;L=2",
    "8,27,111,110": "8,27,111,110;L=2",
    "4,31,111,55,55": "4,31,111,55,55;L=2",
    "16,19,111,55,55": "16,19,111,55,55;L=2",
    "35,111,55,55": "35,111,55,55;L=1",
}
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, str(ROOT))
    import numpy as np, torch
    import texpresso_b200 as T
    from texpresso_b200 import synth
    T.set_device(0)
    res = {}
    for kind, fmt in (("noise_alpha", 2), ("noise_opaque", 0), ("smooth", 2)):
        img = synth.generate(kind, 8192, 1024, 3)
        hin = torch.from_numpy(img.reshape(-1)).pin_memory()
        F = T.Format(fmt)
        out = torch.empty(F.compressed_size(8192, 1024), dtype=torch.uint8).pin_memory()
        prm = T.Params(T.Algorithm.ClusterFit, T.COLOUR_WEIGHTS_PERCEPTUAL, False)
        ts = []
        for i in range(23):
            t0 = time.perf_counter(); F.compress(hin.numpy(), 8192, 1024, prm, output=out.numpy()); ts.append(1e3 * (time.perf_counter() - t0))
        ts = sorted(ts[3:])
        res[f"{kind}_bc{1 if fmt == 0 else 3}"] = [round(ts[0], 3), round(ts[len(ts) // 2], 3)]
    print(json.dumps(res), flush=True)
else:
    for rep in range(2):
        for name, plan in PLANS.items():
            env = dict(os.environ)
            if plan: env["TXP_PLAN"] = plan
            out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout.strip().splitlines()
            print(name, out[-1] if out else "no output", flush=True)
