#!/usr/bin/env python3
"""Test infrastructure (it calls the oracle, so it lives under tests/): randomised sweep of the public entry points against the oracle (beyond the fixed cases): random sizes, formats,
algorithms, weights, alpha weighting, over-long outputs, batches, mip chains, compact pixel layouts, multi-call decode.  Prints a summary."""
import sys, pathlib, random
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import texpresso_b200 as T
from texpresso_b200 import synth
from tests import oracle_lib as O
W = {"u": (O.UNIFORM, T.COLOUR_WEIGHTS_UNIFORM), "p": (O.PERCEPTUAL, T.COLOUR_WEIGHTS_PERCEPTUAL)}


def run(seed=1, n=250, large=False):
    """returns the list of failing cases (empty = every output equals the oracle's)"""
    T.set_device(0)
    rng = random.Random(seed)
    N, LARGE = n, large
    bad = []
    for it in range(N):
        fmt = rng.randrange(5); alg = rng.randrange(3); wn = rng.choice("up"); awa = rng.random() < 0.3
        w = rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 31, 33, 64, 100, 127, 128, 129, 200, 256, 260, 384]); h = rng.choice([1, 2, 3, 4, 5, 6, 8, 11, 16, 30, 64, 65, 131])
        if LARGE:
            w = rng.choice([512, 640, 1024, 1028, 2048, 4096]); h = rng.choice([64, 130, 256, 512, 515, 700])
        kind = rng.choice(["smooth", "noise_alpha", "noise_opaque", "r_rg", "flat", "two", "alpha_edge"])
        if kind == "flat":
            img = np.full((h, w, 4), rng.randrange(256), np.uint8); img[..., 3] = rng.choice([0, 127, 128, 255])
        elif kind == "two":
            img = np.zeros((h, w, 4), np.uint8); img[::2] = (255, 0, 10, 255); img[1::2] = (0, 255, 10, rng.choice([0, 255]))
        elif kind == "alpha_edge":
            img = synth.generate("smooth", w, h, it); img[..., 3] = np.where(np.arange(w)[None, :] % 3 == 0, 127, 128).astype(np.uint8)
        else:
            img = synth.generate(kind, w, h, it)
        img = np.ascontiguousarray(img)
        tp = T.Params(T.Algorithm(alg), W[wn][1], awa); op = O.make_params(alg, W[wn][0], awa)
        tag = (it, fmt, alg, wn, awa, w, h, kind)
        try:
            want = O.compress(fmt, img, w, h, op, threads=8)
            mode = rng.randrange(6)
            if mode == 0:
                got = T.Format(fmt).compress(img, w, h, tp)
            elif mode == 1:                                      # over-long output: extra block rows are encoded fully masked
                extra = rng.randrange(1, 3) * ((w + 3) // 4) * T.Format(fmt).block_size()
                out = np.zeros(want.size + extra, np.uint8)
                T.Format(fmt).compress(img, w, h, tp, output=out)
                want = O.compress(fmt, img, w, h, op, out_len=out.size, threads=8); got = out
            elif mode == 2:
                texs = [(img, w, h)] * rng.randrange(1, 4) + [(np.ascontiguousarray(img[: max(1, h // 2)]), w, max(1, h // 2))]
                outs = T.compress_batch(fmt, texs, tp, n_gpus=1)
                got = outs[0]
                if not np.array_equal(outs[-1], O.compress(fmt, texs[-1][0], w, texs[-1][2], op)): bad.append(tag + ("batch tail",))
                if any(not np.array_equal(o, got) for o in outs[:-1]): bad.append(tag + ("batch dup",))
            elif mode == 3:
                got = T.compress_mipchain(fmt, img, w, h, tp)
                want = np.concatenate([O.compress(fmt, lv, lv.shape[1], lv.shape[0], op, threads=8) for lv in T.generate_mips(img, w, h)])
            elif mode == 4:
                got = T.compress_pixels(fmt, img, w, h, tp)      # RGBA8 layout
            else:
                texs = [(img, w, h)] * rng.randrange(2, 5)
                outs = T.compress_batch_mips(fmt, texs, tp, n_gpus=1)
                got = outs[-1]
                want = np.concatenate([O.compress(fmt, lv, lv.shape[1], lv.shape[0], op, threads=8) for lv in T.generate_mips(img, w, h)])
            if got.size != want.size or not np.array_equal(got, want):
                bad.append(tag + (f"mode {mode}: {int((got.reshape(-1) != want.reshape(-1)).sum()) if got.size == want.size else 'size'} bytes differ",))
            if mode == 0:
                dec = T.Format(fmt).decompress(got, w, h)
                if not np.array_equal(dec, O.decompress(fmt, got, w, h)): bad.append(tag + ("decode",))
        except Exception as e:                                   # noqa: BLE001
            bad.append(tag + (f"exception {type(e).__name__}: {e}",))
    return bad


if __name__ == "__main__":
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 250
    bad = run(int(sys.argv[1]) if len(sys.argv) > 1 else 1, N, len(sys.argv) > 3 and sys.argv[3] == "large")
    print(f"fuzz_api: {N} cases, {len(bad)} failures")
    for b in bad[:20]:
        print("  ", b)
