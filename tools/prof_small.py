"""Driver for ncu captures of the kernels that small launches take: warp-per-block search + its setup kernel (1024^2), the mip-chain kernel and
the texture-group encode (8 textures 1024^2 + mips)."""
import ctypes, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch, texpresso_b200 as T
from texpresso_b200 import synth, _lib
L = _lib.load(); T.set_device(0)
img = synth.generate("noise_alpha", 1024, 1024, 3)
for _ in range(2):
    T.Format.Bc3.compress(img, 1024, 1024, T.Params())
texs = [(synth.generate("smooth", 1024, 1024, 40 + i), 1024, 1024) for i in range(8)]
for _ in range(2):
    T.compress_batch_mips(T.Format.Bc3, texs, T.Params(), n_gpus=1)
