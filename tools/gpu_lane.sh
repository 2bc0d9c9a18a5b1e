python -m pytest tests/test_gpu_parity.py -x -q -k "batch or thread or multi" 2>&1 | tail -3
python - <<'P'
import time, numpy as np, torch, texpresso_b200 as T
from texpresso_b200 import synth
T.set_device(0)
n = 256
pinned = [torch.from_numpy(synth.generate("smooth", 1024, 1024, 6_000_000 + t).reshape(-1)).pin_memory() for t in range(n)]
texs = [(p.numpy(), 1024, 1024) for p in pinned]
for fmt, prm, name in ((T.Format.Bc3, T.Params(), "bc3_cluster"), (T.Format.Bc1, T.Params(T.Algorithm.RangeFit), "bc1_range"), (T.Format.Bc4, T.Params(), "bc4")):
    T.compress_batch(fmt, texs, prm, n_gpus=1)
    t0 = time.perf_counter(); T.compress_batch(fmt, texs, prm, n_gpus=1); dt = time.perf_counter() - t0
    print(name, "batch 256 x 1024^2: %.1f ms, %.0f textures/s, %.0f Mpix/s" % (dt * 1e3, n / dt, n * 1.048576 / dt))
P
