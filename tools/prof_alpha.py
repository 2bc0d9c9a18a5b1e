import ctypes, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch, texpresso_b200 as T
from texpresso_b200 import synth, _lib
L = _lib.load(); T.set_device(0)
w = h = 8192
img = torch.from_numpy(synth.generate("r_rg", w, h, 4).reshape(-1)).cuda()
for fmt, bs in ((T.Format.Bc4, 8), (T.Format.Bc5, 16)):
    out = torch.empty((w // 4) * (h // 4) * bs, dtype=torch.uint8, device="cuda")
    cp = T.Params()._c()
    for _ in range(3):
        _lib.check(L.txp_compress_device(int(fmt), ctypes.c_void_p(img.data_ptr()), w, h, ctypes.byref(cp), ctypes.c_void_p(out.data_ptr()), out.numel(), None))
    torch.cuda.synchronize()
    dimg = torch.empty(w * h * 4, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        _lib.check(L.txp_decompress_device(int(fmt), ctypes.c_void_p(out.data_ptr()), w, h, ctypes.c_void_p(dimg.data_ptr()), dimg.numel(), None))
    torch.cuda.synchronize()
