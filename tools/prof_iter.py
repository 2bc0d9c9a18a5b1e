import ctypes, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch, texpresso_b200 as T
from texpresso_b200 import synth, _lib
L = _lib.load(); T.set_device(0)
w = h = 4096
img = synth.generate("noise_opaque", w, h, 3)
d = torch.from_numpy(img.reshape(-1)).cuda()
out = torch.empty((w // 4) * (h // 4) * 8, dtype=torch.uint8, device="cuda")
cp = T.Params(T.Algorithm.IterativeClusterFit)._c()
for _ in range(2):
    _lib.check(L.txp_compress_device(0, ctypes.c_void_p(d.data_ptr()), w, h, ctypes.byref(cp), ctypes.c_void_p(out.data_ptr()), out.numel(), None))
torch.cuda.synchronize()
