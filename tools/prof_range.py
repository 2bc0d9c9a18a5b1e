"""ncu target: RangeFit BC1 / BC3 and the ClusterFit launch pair (BC3, noise + smooth) on 4096^2 device-resident images."""
import ctypes, sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch, texpresso_b200 as T
from texpresso_b200 import synth, _lib
L = _lib.load(); T.set_device(0)
w = h = 4096
P = T.COLOUR_WEIGHTS_PERCEPTUAL
for kind in ("noise_opaque", "smooth"):
    img = torch.from_numpy(synth.generate(kind, w, h, 3).reshape(-1)).cuda()
    for fmt, bs, alg in ((0, 8, 0), (2, 16, 0), (2, 16, 1)):
        out = torch.empty((w // 4) * (h // 4) * bs, dtype=torch.uint8, device="cuda")
        cp = T.Params(T.Algorithm(alg), P, False)._c()
        for _ in range(2):
            _lib.check(L.txp_compress_device(fmt, ctypes.c_void_p(img.data_ptr()), w, h, ctypes.byref(cp), ctypes.c_void_p(out.data_ptr()), out.numel(), None))
        torch.cuda.synchronize()
