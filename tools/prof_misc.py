import ctypes, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch, texpresso_b200 as T
from texpresso_b200 import synth, _lib
L = _lib.load(); T.set_device(0)
w = h = 8192
img = torch.from_numpy(synth.generate("noise_alpha", w, h, 3).reshape(-1)).cuda()
P = T.COLOUR_WEIGHTS_PERCEPTUAL
for fmt, bs in ((T.Format.Bc1, 8), (T.Format.Bc3, 16)):
    out = torch.empty((w // 4) * (h // 4) * bs, dtype=torch.uint8, device="cuda")
    cp = T.Params(T.Algorithm.RangeFit, P, False)._c()
    for _ in range(2):
        _lib.check(L.txp_compress_device(int(fmt), ctypes.c_void_p(img.data_ptr()), w, h, ctypes.byref(cp), ctypes.c_void_p(out.data_ptr()), out.numel(), None))
    dimg = torch.empty(w * h * 4, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        _lib.check(L.txp_decompress_device(int(fmt), ctypes.c_void_p(out.data_ptr()), w, h, ctypes.c_void_p(dimg.data_ptr()), dimg.numel(), None))
    torch.cuda.synchronize()
