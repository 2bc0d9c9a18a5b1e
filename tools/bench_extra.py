#!/usr/bin/env python3
"""Kernel-only timings of the other BASELINE.json configs (device-resident inputs, CUDA events, L2 flushed
between launches).  One JSON line per case.  Not the driver's bench (that is /bench.py)."""
import argparse, ctypes, json, pathlib, sys
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import texpresso_b200 as T
from texpresso_b200 import synth, _lib

L = _lib.load()
PEAKS = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
HBM = float(PEAKS.get("hbm_gbs", 6650.0))


def time_kernel(fn, reps, flush):
    ts = []
    for _ in range(reps + 2):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts = ts[2:]
    return sum(ts) / len(ts), min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="cfg2,bc4,bc5,range,iter,smooth,decode")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    torch.cuda.set_device(0); T.set_device(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    cases = args.cases.split(",")

    def enc(fmt, d_in, w, h, params, d_out):
        cp = params._c()
        _lib.check(L.txp_compress_device(int(fmt), ctypes.c_void_p(d_in.data_ptr()), w, h, ctypes.byref(cp),
                                         ctypes.c_void_p(d_out.data_ptr()), d_out.numel(), ctypes.c_void_p(stream)))

    def dec(fmt, d_in, w, h, d_out):
        _lib.check(L.txp_decompress_device(int(fmt), ctypes.c_void_p(d_in.data_ptr()), w, h,
                                           ctypes.c_void_p(d_out.data_ptr()), d_out.numel(), ctypes.c_void_p(stream)))

    def report(name, w, h, ms, best, bytes_per_block, extra=None):
        blocks = (w // 4) * (h // 4)
        gbs = bytes_per_block * blocks / (ms / 1e3) / 1e9
        line = {"case": name, "size": f"{w}x{h}", "ms": ms, "ms_best": best, "mpix_s": w * h / (ms / 1e3) / 1e6,
                "algorithmic_gb_s": gbs, "hbm_frac_of_measured": gbs / HBM}
        if extra:
            line.update(extra)
        print(json.dumps(line), flush=True)

    P = T.COLOUR_WEIGHTS_PERCEPTUAL
    if "bc4" in cases or "bc5" in cases or "decode" in cases:
        w = h = 16384
        for kind in ("r_rg", "r_rg_smooth", "smooth"):
            if kind != "r_rg" and not ("bc4" in cases or "bc5" in cases):
                continue
            img = torch.from_numpy(synth.generate(kind, w, h, 4).reshape(-1)).cuda()
            for name, fmt, bs in (("bc4", T.Format.Bc4, 8), ("bc5", T.Format.Bc5, 16)):
                if name not in cases and "decode" not in cases:
                    continue
                out = torch.empty((w // 4) * (h // 4) * bs, dtype=torch.uint8, device="cuda")
                ms, best = time_kernel(lambda: enc(fmt, img, w, h, T.Params(), out), args.reps, flush)
                if name in cases:
                    report(f"{name}_encode_{kind}", w, h, ms, best, 64 + bs)
                if "decode" in cases and kind == "r_rg":
                    dimg = torch.empty(w * h * 4, dtype=torch.uint8, device="cuda")
                    ms, best = time_kernel(lambda: dec(fmt, out, w, h, dimg), args.reps, flush)
                    report(f"{name}_decode", w, h, ms, best, 64 + bs)
                    del dimg
                del out
            del img
    if "decode" in cases:
        w = h = 8192
        img = torch.from_numpy(synth.generate("noise_alpha", w, h, 3).reshape(-1)).cuda()
        for name, fmt, bs in (("bc1", T.Format.Bc1, 8), ("bc2", T.Format.Bc2, 16), ("bc3", T.Format.Bc3, 16)):
            out = torch.empty((w // 4) * (h // 4) * bs, dtype=torch.uint8, device="cuda")
            enc(fmt, img, w, h, T.Params(T.Algorithm.RangeFit, P, False), out)
            dimg = torch.empty(w * h * 4, dtype=torch.uint8, device="cuda")
            ms, best = time_kernel(lambda: dec(fmt, out, w, h, dimg), args.reps, flush)
            report(f"{name}_decode", w, h, ms, best, 64 + bs)
            del out, dimg
        del img
    if "range" in cases:
        for (w, h) in ((1024, 1024), (8192, 8192)):
            img = torch.from_numpy(synth.generate("noise_opaque", w, h, 1).reshape(-1)).cuda()
            for name, fmt, bs in (("bc1", T.Format.Bc1, 8), ("bc3", T.Format.Bc3, 16)):
                out = torch.empty((w // 4) * (h // 4) * bs, dtype=torch.uint8, device="cuda")
                ms, best = time_kernel(lambda: enc(fmt, img, w, h, T.Params(T.Algorithm.RangeFit, P, False), out), args.reps, flush)
                report(f"{name}_rangefit_noise_opaque", w, h, ms, best, 64 + bs)
            del img
    if "cfg2" in cases:
        w = h = 4096                                      # BASELINE config 2: BC3 ClusterFit, 4096^2 noise_alpha, seed 2, default Params
        img = torch.from_numpy(synth.generate("noise_alpha", w, h, 2).reshape(-1)).cuda()
        out = torch.empty((w // 4) * (h // 4) * 16, dtype=torch.uint8, device="cuda")
        ms, best = time_kernel(lambda: enc(T.Format.Bc3, img, w, h, T.Params(), out), args.reps, flush)
        report("cfg2_bc3_cluster_noise_alpha", w, h, ms, best, 80, {"fp32_issue_frac": 967 * 159 * (w // 4) * (h // 4) / (ms / 1e3) / 37.22e12})
        del img, out
    if "iter" in cases:
        w = h = 8192
        img = torch.from_numpy(synth.generate("noise_opaque", w, h, 3).reshape(-1)).cuda()
        out = torch.empty((w // 4) * (h // 4) * 8, dtype=torch.uint8, device="cuda")
        ms, best = time_kernel(lambda: enc(T.Format.Bc1, img, w, h, T.Params(T.Algorithm.IterativeClusterFit, P, False), out), max(2, args.reps // 2), flush)
        report("bc1_iterative_noise_opaque", w, h, ms, best, 72)
        del img, out
    if "smooth" in cases:
        w = h = 8192
        img = torch.from_numpy(synth.generate("smooth", w, h, 5).reshape(-1)).cuda()
        for name, fmt, bs in (("bc1", T.Format.Bc1, 8), ("bc3", T.Format.Bc3, 16)):
            out = torch.empty((w // 4) * (h // 4) * bs, dtype=torch.uint8, device="cuda")
            for aname, alg in (("cluster", T.Algorithm.ClusterFit), ("iterative", T.Algorithm.IterativeClusterFit)):
                ms, best = time_kernel(lambda: enc(fmt, img, w, h, T.Params(alg, P, False), out), args.reps, flush)
                report(f"{name}_{aname}_smooth", w, h, ms, best, 64 + bs)
        del img


def mips_case(n_tex, n_gpus):
    """BASELINE config 5 (scaled down): textures 1024^2 `smooth` + full mip chains, BC3 ClusterFit, end to end."""
    import time
    texs = [(synth.generate("smooth", 1024, 1024, 5_000_000 + t), 1024, 1024) for t in range(n_tex)]
    pinned = [torch.from_numpy(t[0].reshape(-1)).pin_memory() for t in texs]
    texs = [(p.numpy(), 1024, 1024) for p in pinned]
    size = L.txp_mipchain_compressed_size(2, 1024, 1024)
    outs_t = [torch.empty(size, dtype=torch.uint8).pin_memory() for _ in range(n_tex)]
    outs = [o.numpy() for o in outs_t]
    T.compress_batch_mips(T.Format.Bc3, texs, T.Params(), n_gpus=n_gpus, outputs=outs)      # warm-up
    t0 = time.perf_counter()
    T.compress_batch_mips(T.Format.Bc3, texs, T.Params(), n_gpus=n_gpus, outputs=outs)
    dt = time.perf_counter() - t0
    pix = n_tex * sum(w * h for w, h in T.mip_levels(1024, 1024))
    print(json.dumps({"case": "bc3_cluster_batch_mips_e2e", "textures": n_tex, "n_gpus": n_gpus, "s": dt,
                      "mpix_s": pix / dt / 1e6, "textures_per_s": n_tex / dt}), flush=True)


if __name__ == "__main__":
    if "--mips" in sys.argv:
        i = sys.argv.index("--mips")
        torch.cuda.set_device(0); T.set_device(0)
        mips_case(int(sys.argv[i + 1]), int(sys.argv[i + 2]) if len(sys.argv) > i + 2 else 1)
        sys.exit(0)
    main()
