import ctypes, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch, texpresso_b200 as T
from texpresso_b200 import synth, _lib
L = _lib.load(); T.set_device(0)
w = h = 8192 if len(sys.argv) < 2 else int(sys.argv[1])
kind = "noise_alpha" if len(sys.argv) < 3 else sys.argv[2]
img = torch.from_numpy(synth.generate(kind, w, h, 3).reshape(-1)).cuda()
out = torch.empty((w // 4) * (h // 4) * 16, dtype=torch.uint8, device="cuda")
cp = T.Params()._c()
for _ in range(2):
    _lib.check(L.txp_compress_device(2, ctypes.c_void_p(img.data_ptr()), w, h, ctypes.byref(cp), ctypes.c_void_p(out.data_ptr()), out.numel(), None))
torch.cuda.synchronize()
