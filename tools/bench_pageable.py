#!/usr/bin/env python3
"""End-to-end Format.compress / decompress with PAGEABLE (plain numpy) buffers against pinned ones: what a caller who hands over ordinary slices
(the reference's &[u8] signature) gets.  Median of 7 calls."""
import json, sys, time, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import texpresso_b200 as T
from texpresso_b200 import synth
T.set_device(0)
def med(fn, n=7):
    ts = []
    for i in range(n + 2):
        t0 = time.perf_counter(); fn(); ts.append(1e3 * (time.perf_counter() - t0))
    return round(sorted(ts[2:])[n // 2], 3)
cases = [("bc3_cluster", 2, T.Params(), "noise_alpha", 8192), ("bc3_cluster", 2, T.Params(), "smooth", 8192), ("bc1_range", 0, T.Params(T.Algorithm.RangeFit), "smooth", 8192),
         ("bc4", 3, T.Params(), "r_rg", 8192), ("bc3_cluster", 2, T.Params(), "smooth", 2048)]
for name, fmt, prm, kind, side in cases:
    F = T.Format(fmt)
    img = synth.generate(kind, side, side, 3)
    pin_in = torch.from_numpy(img.reshape(-1)).pin_memory(); pin_out = torch.empty(F.compressed_size(side, side), dtype=torch.uint8).pin_memory()
    pag_in = img.reshape(-1).copy(); pag_out = np.empty(F.compressed_size(side, side), np.uint8)
    rec = {"case": name, "input": kind, "side": side,
           "pinned_ms": med(lambda: F.compress(pin_in.numpy(), side, side, prm, output=pin_out.numpy())),
           "pageable_ms": med(lambda: F.compress(pag_in, side, side, prm, output=pag_out))}
    assert np.array_equal(pag_out, pin_out.numpy())
    dec_pin = torch.empty(side * side * 4, dtype=torch.uint8).pin_memory(); dec_pag = np.empty(side * side * 4, np.uint8)
    rec["decode_pinned_ms"] = med(lambda: F.decompress(pin_out.numpy(), side, side, output=dec_pin.numpy()))
    rec["decode_pageable_ms"] = med(lambda: F.decompress(pag_out, side, side, output=dec_pag))
    print(json.dumps(rec), flush=True)
