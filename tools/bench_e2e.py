#!/usr/bin/env python3
"""End-to-end (pinned host -> H2D -> kernels -> D2H -> pinned host) timings through Format.compress / decompress
for the bandwidth-bound configurations; one JSON line per case."""
import json, pathlib, sys, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch, texpresso_b200 as T
from texpresso_b200 import synth
T.set_device(0)
P = T.COLOUR_WEIGHTS_PERCEPTUAL
cases = [("bc4", T.Format.Bc4, 16384, "r_rg", T.Params()), ("bc5", T.Format.Bc5, 16384, "r_rg", T.Params()),
         ("bc1_rangefit", T.Format.Bc1, 8192, "noise_opaque", T.Params(T.Algorithm.RangeFit, P, False)),
         ("bc3_rangefit", T.Format.Bc3, 8192, "noise_alpha", T.Params(T.Algorithm.RangeFit, P, False))]
for name, fmt, n, kind, prm in cases:
    img = torch.from_numpy(synth.generate(kind, n, n, 4).reshape(-1)).pin_memory()
    out = torch.empty(fmt.compressed_size(n, n), dtype=torch.uint8).pin_memory()
    dec = torch.empty(n * n * 4, dtype=torch.uint8).pin_memory()
    for _ in range(2): fmt.compress(img.numpy(), n, n, prm, output=out.numpy())
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); fmt.compress(img.numpy(), n, n, prm, output=out.numpy()); ts.append(time.perf_counter() - t0)
    enc = min(ts)
    for _ in range(2): fmt.decompress(out.numpy(), n, n, output=dec.numpy())
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); fmt.decompress(out.numpy(), n, n, output=dec.numpy()); ts.append(time.perf_counter() - t0)
    d = min(ts)
    print(json.dumps({"case": name, "size": f"{n}x{n}", "encode_e2e_ms": round(enc * 1e3, 2), "encode_mpix_s": round(n * n / enc / 1e6),
                      "encode_h2d_gb_s": round(n * n * 4 / enc / 1e9, 1), "decode_e2e_ms": round(d * 1e3, 2), "decode_d2h_gb_s": round(n * n * 4 / d / 1e9, 1)}), flush=True)
    del img, out, dec

# compact pixel layouts (txp_compress_pixels): the same BC4 / BC5 images as L8 / LA8-style 1- and 2-byte pixels
import numpy as np
for name, fmt, n, ch in (("bc4_from_L8", T.Format.Bc4, 16384, 1), ("bc5_from_RG8", T.Format.Bc5, 16384, 2), ("bc1_clusterfit_from_RGB8", T.Format.Bc1, 8192, 3)):
    src = synth.generate("r_rg" if ch < 3 else "noise_opaque", n, n, 4).reshape(n, n, 4)
    pix = np.ascontiguousarray(src[..., :ch])
    img = torch.from_numpy(pix.reshape(-1)).pin_memory()
    out = torch.empty(fmt.compressed_size(n, n), dtype=torch.uint8).pin_memory()
    prm = T.Params()
    for _ in range(2): T.compress_pixels(fmt, img.numpy(), n, n, prm, output=out.numpy(), layout=(T.PIXELS_RG8 if ch == 2 else ch))
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); T.compress_pixels(fmt, img.numpy(), n, n, prm, output=out.numpy(), layout=(T.PIXELS_RG8 if ch == 2 else ch)); ts.append(time.perf_counter() - t0)
    enc = min(ts)
    print(json.dumps({"case": name, "size": f"{n}x{n}", "bytes_per_pixel": ch, "encode_e2e_ms": round(enc * 1e3, 2), "encode_mpix_s": round(n * n / enc / 1e6),
                      "encode_h2d_gb_s": round(n * n * ch / enc / 1e9, 1)}), flush=True)
    del img, out
