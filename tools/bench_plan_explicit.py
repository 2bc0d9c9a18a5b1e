#!/usr/bin/env python3
"""Explicit pipeline chunk plans (TXP_PLAN, experiments only) for the whole 8192^2 texture through Format.compress on pinned buffers."""
import json, os, sys, time, pathlib, subprocess
ROOT = pathlib.Path(__file__).resolve().parent.parent
PLANS = {
    "old": "32,135x15;L=1",
    "new": "12,39,111x17,55,55;L=2",
    "a_oldstart_tailwarp": "32,111x18,18;L=1",
    "b_no_single_tail": "12,38,111x18;L=2",
    "c_single_head": "51,111x17,55,55;L=2",
    "d_4round": "12,39,222x8,111,55,55;L=2",
    "e_old_first_aligned": "50,111x18;L=1",
}
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, str(ROOT))
    import numpy as np, torch
    import texpresso_b200 as T
    from texpresso_b200 import synth
    T.set_device(0)
    res = {"plan": os.environ.get("TXP_PLAN", "auto")}
    for kind, fmt in (("noise_alpha", 2), ("smooth", 2), ("smooth", 0)):
        img = synth.generate(kind, 8192, 8192, 3)
        hin = torch.from_numpy(img.reshape(-1)).pin_memory()
        F = T.Format(fmt)
        out = torch.empty(F.compressed_size(8192, 8192), dtype=torch.uint8).pin_memory()
        prm = T.Params(T.Algorithm.ClusterFit, T.COLOUR_WEIGHTS_PERCEPTUAL, False)
        ts = []
        for i in range(13):
            t0 = time.perf_counter(); F.compress(hin.numpy(), 8192, 8192, prm, output=out.numpy()); ts.append(1e3 * (time.perf_counter() - t0))
        ts = sorted(ts[2:])
        res[f"{kind}_bc{1 if fmt == 0 else 3}"] = [round(ts[0], 3), round(ts[len(ts) // 2], 3)]
    print(json.dumps(res), flush=True)
else:
    for rep in range(2):
        for name, plan in PLANS.items():
            env = dict(os.environ, TXP_PLAN=plan)
            out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout.strip().splitlines()
            print(name, out[-1] if out else "no output", flush=True)
