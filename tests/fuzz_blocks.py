#!/usr/bin/env python3
"""Test infrastructure (calls the oracle): large one-off version of tests/test_gpu_fuzz.py -- N random blocks per configuration and seed through the CUDA
path (warp-per-block kernels; with TXP_COLOUR_VARIANT=lane the lane-per-block kernels) and the C oracle.  usage: fuzz_blocks.py <blocks> <seeds>"""
import sys, pathlib, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import texpresso_b200 as T
from tests import oracle_lib as O
from tests import test_gpu_fuzz as F
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
SEEDS = int(sys.argv[2]) if len(sys.argv) > 2 else 1
F.N = N
CONFIGS = [(0, 0, O.PERCEPTUAL, False), (0, 1, O.PERCEPTUAL, False), (0, 2, O.PERCEPTUAL, False), (0, 1, O.UNIFORM, True), (1, 1, O.PERCEPTUAL, False), (1, 2, O.UNIFORM, False),
           (2, 0, O.UNIFORM, True), (2, 1, O.PERCEPTUAL, False), (2, 2, O.PERCEPTUAL, True), (3, 1, O.PERCEPTUAL, False), (4, 1, O.PERCEPTUAL, False), (0, 2, (0.3, 1.7, 0.05), True)]
T.set_device(0)
tot = bad = 0
t0 = time.time()
for s in range(SEEDS):
    for fmt, alg, w, awa in CONFIGS:
        blocks, masks = F._corpus(50000 + 1000 * s + 17 * fmt + alg)
        got = T.compress_blocks(fmt, blocks, masks, T.Params(T.Algorithm(alg), tuple(w), awa))
        want = O.compress_blocks(fmt, blocks, masks, O.make_params(alg, w, awa))
        d = int((got != want).any(axis=1).sum())
        tot += N; bad += d
        if d:
            print("MISMATCH", s, fmt, alg, w, awa, d, flush=True)
    print(f"seed {s}: {tot} blocks, {bad} mismatches, {time.time() - t0:.0f} s", flush=True)
print(f"fuzz_blocks: {tot} blocks, {bad} mismatches")
