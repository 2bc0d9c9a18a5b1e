"""world_size-2 gloo test (CPU) of the multi-rank path: block-row sharding as bench.py / txp_compress_multi do it.
Each rank encodes only its own block rows (with the CPU oracle standing in for the kernels -- there is no GPU
here), the slices are gathered, and the result must equal the whole-image encode byte for byte."""
import os, socket, sys, pathlib
import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, fmt, w, h, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import torch, torch.distributed as dist
    import texpresso_b200 as T
    from texpresso_b200 import synth
    from tests import oracle_lib as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bs = T.Format(fmt).block_size()
        bw = T.num_blocks(w)
        r0, r1 = T.shard_rows(h, rank, world)
        # the rank generates and encodes only its rows (counter-based generator), as bench.py does
        y0, y1 = 4 * r0, min(4 * r1, h)
        img = synth.generate("noise_alpha", w, h, seed=9, y0=y0, y1=y1)
        p = O.make_params(O.CLUSTER_FIT, O.PERCEPTUAL, False)
        part = O.compress(fmt, img, w, y1 - y0, p) if y1 > y0 else np.zeros(0, np.uint8)
        assert part.size == (r1 - r0) * bw * bs
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([part.size], dtype=torch.int64))
        mx = int(max(s.item() for s in sizes))
        buf = torch.zeros(mx, dtype=torch.uint8); buf[:part.size] = torch.from_numpy(part)
        outs = [torch.zeros(mx, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(outs, buf)
        # device-time style reduction used by bench.py: max over ranks
        t = torch.tensor([float(rank + 1)]); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
        if rank == 0:
            whole = np.concatenate([o.numpy()[: int(s.item())] for o, s in zip(outs, sizes)])
            full = synth.generate("noise_alpha", w, h, seed=9)
            want = O.compress(fmt, full, w, h, p)
            ret["ok"] = bool(np.array_equal(whole, want))
            ret["n"] = int(whole.size)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fmt,w,h", [(0, 64, 40), (2, 36, 30), (4, 20, 7)])
def test_two_rank_block_row_sharding(fmt, w, h):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, fmt, w, h, ret)) for r in range(2)]
    for p in procs: p.start()
    for p in procs: p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret.get("ok") is True
