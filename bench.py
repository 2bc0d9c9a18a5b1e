#!/usr/bin/env python3
"""bench.py -- BC1/BC3 ClusterFit Mpix/s on an 8192x8192 synthetic RGBA texture (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W             (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                      reference algorithm on the host CPU cores

A "step" = one BC1 ClusterFit pass + one BC3 ClusterFit pass over the texture (PERCEPTUAL weights, the
reference's default Params).  With N ranks the texture is sharded by block rows (reference grain:
lib.rs:300-305), every rank encodes its own slice, no collective on the data path -> strong scaling.

value  : whole-job Mpix/s, inputs resident in HBM, device time (CUDA events on the launching stream), max over ranks
e2e    : same metric through the public host API (Format.compress on pinned host buffers: H2D + kernels + D2H)
roofline : dominant kernels (cluster_setup_sorted_kernel<BC3> + cluster_lane_kernel<BC3>, one BC3 ClusterFit launch pair),
           algorithmic fp32 ops / event time vs the non-FMA FP32 issue peak (SMs x 128 lanes x clock).
           Not a tensor/HBM kernel: see DESIGN.md.
cpu_baseline : the CPU oracle (a C port of the reference algorithm; the Rust reference cannot be built
           in this image) on all host cores, on a bounded crop of the same workload.
"""
import argparse, json, os, statistics, subprocess, sys, threading, time, pathlib

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W = H = 8192
SEED = 3
FLOPS_BC3 = 967 * 159                 # fp32 ops per 16-colour block, SURVEY.md App. C
FLOPS_BC1 = 151 * 138 + 967 * 159
UNIT = "Mpix/s"
# from the ncu --set full capture of the BC3 launch pair (profiles/, per round): FMA-pipe utilisation of cluster_lane_kernel<BC3> and
# DRAM bytes (read + write) of setup + search for one 8192^2 launch
NCU_FMA_PIPE_FRAC = 0.807
NCU_FMA_PIPE_SOURCE = "ncu sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active of cluster_lane_kernel<BC3>, profiles/ncu_lane_r02_summary.txt"
NCU_TRAFFIC_BYTES_8192 = 1389.6e6
NCU_TRAFFIC_SOURCE = "profiles/ncu_setup_r02_summary.txt + ncu_lane_r02_summary.txt (dram read + write): setup 378.3 MB + 472.6 MB, search 476.1 MB + 62.6 MB (96-byte point records + 16-byte block records of scratch between the two kernels)"
METRIC = "BC1/BC3 ClusterFit Mpix/s (8192^2 synthetic RGBA)"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_time(crop, threads, repeats=1, keep=None):
    """Times the CPU oracle (port of the reference algorithm) on a crop x crop sample of the workload.
    Returns (Mpix/s over BC1+BC3, seconds).  keep (dict): receives the oracle's blocks and the crop's pixels for the parity check."""
    import numpy as np
    from texpresso_b200 import synth
    from tests import oracle_lib as O                     # allowed here: cpu_baseline / --impl reference leg
    img = synth.generate("noise_alpha", W, H, SEED, y0=0, y1=crop)[:, :crop].copy()
    opaque = img.copy(); opaque[..., 3] = 255
    p = O.make_params(O.CLUSTER_FIT, O.PERCEPTUAL, False)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        o1 = O.compress(O.BC1, opaque, crop, crop, p, threads=threads)
        o3 = O.compress(O.BC3, img, crop, crop, p, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    if keep is not None:
        keep.update(crop=crop, bc1=o1, bc3=o3, opaque=opaque, img=img)
    return 2 * crop * crop / best / 1e6, best


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    threads = host_threads()
    crop = 256
    mp, dt = cpu_reference_time(crop, threads)            # calibrate
    # size each step at roughly 2-3 s so warmup+steps stay within a few minutes
    while dt < 2.0 and crop < 2048:
        crop *= 2
        mp, dt = cpu_reference_time(crop, threads)
    times = []
    for i in range(args.warmup + args.steps):
        mp_i, dt_i = cpu_reference_time(crop, threads)
        if i >= args.warmup:
            times.append(dt_i)
    ms = 1e3 * sum(times) / len(times)
    val = 2 * crop * crop / (ms / 1e3) / 1e6
    sample = f"{crop}x{crop} crop (top-left) of the 8192^2 workload per step, BC1+BC3 ClusterFit"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BC1+BC3 ClusterFit, 8192x8192 synthetic RGBA (noise, seed 3), PERCEPTUAL weights; "
                                   "CPU arm timed on a bounded crop", "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "note": "C port of the reference algorithm (oracle/txp_oracle.c); the Rust reference cannot be built here (no cargo/rustc)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---- parity of GPU output against the CPU oracle (SURVEY 8(d) "Metric", BASELINE.md 4 "parity columns") ----------------
def block_rows_of(flat_u8, bs, bw_total, row_lo, row_hi, nbx):
    """blocks [row_lo,row_hi) x [0,nbx) of a block-row-major output with bw_total blocks per row -> (n, bs) array"""
    import numpy as np
    a = np.asarray(flat_u8).reshape(-1, bw_total, bs)
    return a[row_lo:row_hi, :nbx].reshape(-1, bs)


def parity_counts(fmt, got_blocks, want_blocks, pixels, op):
    """identical blocks, and for differing ones whether ours is worse in weighted squared error than the oracle's choice"""
    import numpy as np
    from tests import oracle_lib as O
    n = got_blocks.shape[0]
    diff = np.nonzero((got_blocks != want_blocks).any(axis=1))[0]
    worse = 0
    if diff.size:
        hh, ww = pixels.shape[0], pixels.shape[1]
        blk = pixels.reshape(hh // 4, 4, ww // 4, 4, 4).transpose(0, 2, 1, 3, 4).reshape(-1, 64)
        off = 0 if fmt == 0 else 8
        for i in diff[:4096]:
            if off and bytes(got_blocks[i][:off]) != bytes(want_blocks[i][:off]):
                worse += 1                                  # alpha halves must be bit exact
                continue
            eg = O.colour_block_error(fmt, blk[i], 0xFFFF, op, got_blocks[i][off:off + 8])
            ew = O.colour_block_error(fmt, blk[i], 0xFFFF, op, want_blocks[i][off:off + 8])
            if eg > ew * (1 + 1e-6) + 1e-12:
                worse += 1
    return n, n - int(diff.size), worse


def run_ours(args):
    import numpy as np
    import torch
    import texpresso_b200 as T
    from texpresso_b200 import synth, _lib
    import ctypes

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    T.set_device(local)
    dist = None
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        if dist is None:
            return list(vals)
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def sum_over_ranks(vals):
        if dist is None:
            return [int(v) for v in vals]
        t = torch.tensor([int(v) for v in vals], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        return [int(v) for v in t.tolist()]

    # ---- this rank's shard: block rows [r0, r1) -----------------------------------------------------------
    r0, r1 = T.shard_rows(H, rank, world)
    if args.shard_of > 1 and world == 1:                  # tuning aid: rank 0's shard of an N-rank run on one GPU (values are per shard)
        r0, r1 = T.shard_rows(H, 0, args.shard_of)
    rows = r1 - r0
    hs = 4 * rows
    img = synth.generate("noise_alpha", W, H, SEED, y0=4 * r0, y1=4 * r1)
    h_bc3 = torch.from_numpy(img.reshape(-1)).pin_memory()
    opq = img.copy(); opq[..., 3] = 255
    h_bc1 = torch.from_numpy(opq.reshape(-1)).pin_memory()
    del img, opq
    d_bc1, d_bc3 = h_bc1.cuda(), h_bc3.cuda()
    nblk = rows * (W // 4)
    d_out1 = torch.empty(nblk * 8, dtype=torch.uint8, device="cuda")
    d_out3 = torch.empty(nblk * 16, dtype=torch.uint8, device="cuda")
    h_out1 = torch.empty(nblk * 8, dtype=torch.uint8).pin_memory()
    h_out3 = torch.empty(nblk * 16, dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    L = _lib.load()
    iterative = args.workload == "iterative"
    PERC = T.COLOUR_WEIGHTS_PERCEPTUAL
    params = T.Params(T.Algorithm.IterativeClusterFit if iterative else T.Algorithm.ClusterFit, PERC, False)
    cp = params._c()
    stream = torch.cuda.current_stream().cuda_stream

    def dev_call(fmt, d_in, w, h, prm, d_out):
        c = prm._c()
        _lib.check(L.txp_compress_device(int(fmt), ctypes.c_void_p(d_in.data_ptr()), w, h, ctypes.byref(c),
                                         ctypes.c_void_p(d_out.data_ptr()), d_out.numel(), ctypes.c_void_p(stream)))

    def dev_encode(fmt, d_in, d_out):
        dev_call(fmt, d_in, W, hs, params, d_out)

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timed(fn, reps=3, warm=1):
        """mean CUDA-event ms of fn() over reps launches, L2 flushed before each (outside the events)"""
        ts = []
        for i in range(warm + reps):
            flush.fill_(i & 255)
            a, b = ev(), ev()
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            if i >= warm:
                ts.append(a.elapsed_time(b))
        return sum(ts) / len(ts)

    def device_step():
        """returns (ms_bc1, ms_bc3) for one step; L2 flushed before each kernel, outside the timed events"""
        flush.fill_(rank & 255)
        a, b = ev(), ev()
        a.record(); dev_encode(T.Format.Bc1, d_bc1, d_out1); b.record()
        if iterative:                                     # config 3 is BC1 only
            torch.cuda.synchronize()
            return a.elapsed_time(b), 0.0
        flush.fill_((rank + 1) & 255)
        c, d = ev(), ev()
        c.record(); dev_encode(T.Format.Bc3, d_bc3, d_out3); d.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b), c.elapsed_time(d)

    def host_step():
        T.Format.Bc1.compress(h_bc1.numpy(), W, hs, params, output=h_out1.numpy())
        if not iterative:
            T.Format.Bc3.compress(h_bc3.numpy(), W, hs, params, output=h_out3.numpy())

    # ---- kernel-only (device resident) ---------------------------------------------------------------------
    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = T.kernel_launches()
    t1 = t3 = 0.0
    for _ in range(args.steps):
        a, b = device_step()
        t1 += a; t3 += b
    barrier()
    launches = T.kernel_launches() - launches0
    dev_ms, t1, t3 = max_over_ranks([t1 + t3, t1, t3])
    launches = sum_over_ranks([launches])[0]

    # ---- end to end through the host API (pinned host in, pinned host out) --------------------------------------
    for _ in range(max(1, min(args.warmup, 2))):
        host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    barrier()
    e2e_ms = max_over_ranks([1e3 * (time.perf_counter() - t0)])[0]
    clocks = sampler.stop() if rank == 0 else None

    # correctness guard: device-resident and host-API paths must agree byte for byte
    same = bool(torch.equal(d_out1.cpu(), h_out1) and (iterative or torch.equal(d_out3.cpu(), h_out3)))
    same = sum_over_ranks([0 if same else 1])[0] == 0

    # ---- parity against the CPU oracle --------------------------------------------------------------------------------------
    # N = 1: the whole cpu_baseline crop (its oracle output is the checker's answer); N > 1: every rank checks a slice of its
    # own shard (top 64 pixel rows x 1024 pixels) and the counts are summed over ranks.
    from tests import oracle_lib as O                     # the checker; never on the timed path
    op = O.make_params(O.ITERATIVE_CLUSTER_FIT if iterative else O.CLUSTER_FIT, O.PERCEPTUAL, False)
    cpu_line, keep = None, {}
    do_cpu = world == 1 and not args.no_cpu_baseline and not iterative
    if do_cpu:
        threads = host_threads()
        crop = 256
        mp, dt = cpu_reference_time(crop, threads, keep=keep)
        while dt < 4.0 and crop < 4096:
            crop *= 2
            mp, dt = cpu_reference_time(crop, threads, keep=keep)
        cpu_line = {"value": mp, "unit": UNIT, "cores": threads, "kind": "port",
                    "sample": f"{crop}x{crop} crop of the same 8192^2 workload, BC1+BC3 ClusterFit, {dt:.1f} s",
                    "note": "C port of the reference algorithm (oracle/txp_oracle.c); the Rust reference cannot be built here"}
        pw, ph = keep["crop"], keep["crop"]
        want1, want3, pix1, pix3 = keep["bc1"], keep["bc3"], keep["opaque"], keep["img"]
    else:
        pw, ph = 1024, min(64, hs)
        pix3 = synth.generate("noise_alpha", W, H, SEED, y0=4 * r0, y1=4 * r0 + ph)[:, :pw].copy()
        pix1 = pix3.copy(); pix1[..., 3] = 255
        want1 = O.compress(O.BC1, pix1, pw, ph, op, threads=4)
        want3 = None if iterative else O.compress(O.BC3, pix3, pw, ph, op, threads=4)
    parity = {}
    g1 = block_rows_of(d_out1.cpu().numpy(), 8, W // 4, 0, ph // 4, pw // 4)
    c1 = parity_counts(0, g1, np.asarray(want1).reshape(-1, 8), pix1, op)
    c3 = (0, 0, 0)
    if not iterative:
        g3 = block_rows_of(d_out3.cpu().numpy(), 16, W // 4, 0, ph // 4, pw // 4)
        c3 = parity_counts(2, g3, np.asarray(want3).reshape(-1, 16), pix3, op)
    tot = sum_over_ranks(list(c1) + list(c3))
    for name, (n, ident, worse) in (("bc1", tot[0:3]), ("bc3", tot[3:6])):
        if n:
            parity[name] = {"blocks_checked": n, "identical": ident, "pct": 100.0 * ident / n, "worse_than_ref": worse}
    parity["oracle"] = "oracle/txp_oracle.c (C port of the reference algorithm)"
    parity["sample"] = (f"top-left {pw}x{ph} of the 8192^2 texture (the cpu_baseline crop)" if do_cpu else
                        f"top {ph} pixel rows x {pw} pixels of every rank's shard, counts summed over ranks")

    # ---- multi-GPU entry points of the library (one process driving N devices): correctness only, outside every timed region --------
    multi_ok = None
    if world > 1:
        barrier()
        if rank == 0:
            try:
                mw, mh = 1024, 1000
                mimg = synth.generate("smooth", mw, mh, 77)
                ok = True
                for fmt, prm in ((T.Format.Bc1, T.Params(T.Algorithm.ClusterFit, PERC, False)), (T.Format.Bc3, T.Params(T.Algorithm.RangeFit, PERC, True)),
                                 (T.Format.Bc5, T.Params())):
                    one = fmt.compress(mimg, mw, mh, prm)
                    many = T.compress_multi(fmt, mimg, mw, mh, prm, n_gpus=world)
                    ok = ok and bool(np.array_equal(one, many))
                    ok = ok and bool(np.array_equal(fmt.decompress(one, mw, mh), T.decompress_multi(fmt, one, mw, mh, n_gpus=world)))
                texs = [(synth.generate("smooth", 256, 252, 900 + t), 256, 252) for t in range(2 * world + 1)]
                lone = [T.Format.Bc3.compress(a, w_, h_) for a, w_, h_ in texs]
                ok = ok and all(np.array_equal(x, y) for x, y in zip(lone, T.compress_batch(T.Format.Bc3, texs, n_gpus=world)))
                back = T.decompress_batch(T.Format.Bc3, [(b_, 256, 252) for b_ in lone], n_gpus=world)
                ok = ok and all(np.array_equal(x, T.Format.Bc3.decompress(b_, 256, 252)) for x, b_ in zip(back, lone))
                multi_ok = bool(ok)
            except Exception as e:                         # reported, not fatal: the timed numbers above stand on their own
                multi_ok = f"error: {e}"
            T.set_device(local)
        barrier()

    # ---- the other BASELINE configurations, device-resident, CUDA events (each record carries its own roofline) ------------------------
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    props = torch.cuda.get_device_properties(local)
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    fp32_peak = props.multi_processor_count * 128 * sm_max * 1e6 / 1e12      # Tflop/s, non-FMA issue
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    fp32_measured = None
    if rank == 0:
        v = ctypes.c_double(0.0)
        if L.txp_measure_fp32_issue(ctypes.byref(v)) == 0:
            fp32_measured = v.value / 1e12

    def hbm_roof(kernel, bytes_per_block, blocks, ms):
        gbs = bytes_per_block * blocks / (ms / 1e3) / 1e9
        return {"kernel": kernel, "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                "bytes_per_block": bytes_per_block, "peak_source": hbm_src, "traffic": None}

    def fp32_roof(kernel, flops_per_block, blocks, ms, note=None):
        tf = flops_per_block * blocks / (ms / 1e3) / 1e12
        r = {"kernel": kernel, "bound": "fp32_issue", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak,
             "flops_per_block": flops_per_block, "traffic": None}
        if fp32_measured:
            r["peak_measured"] = fp32_measured; r["frac_of_measured"] = tf / fp32_measured
        if note:
            r["note"] = note
        return r

    configs = {}
    if not args.no_configs and not iterative:
        # cfg3: BC1 IterativeClusterFit, PERCEPTUAL, this rank's block rows of the same 8192^2 noise_opaque texture
        pit = T.Params(T.Algorithm.IterativeClusterFit, PERC, False)
        ms3 = max_over_ranks([timed(lambda: dev_call(T.Format.Bc1, d_bc1, W, hs, pit, d_out1), reps=2)])[0]
        t0 = time.perf_counter()
        T.Format.Bc1.compress(h_bc1.numpy(), W, hs, pit, output=h_out1.numpy())          # warm
        barrier(); t0 = time.perf_counter()
        T.Format.Bc1.compress(h_bc1.numpy(), W, hs, pit, output=h_out1.numpy())
        barrier()
        e3 = max_over_ranks([1e3 * (time.perf_counter() - t0)])[0]
        # candidate counts of the reference control flow on a sample (the oracle counts them: SURVEY 8(d), App. C)
        sw_, sh_ = 256, min(256, hs)
        spx = synth.generate("noise_opaque", W, H, SEED, y0=4 * r0, y1=4 * r0 + sh_)[:, :sw_].copy()
        want, st = O.compress(O.BC1, spx, sw_, sh_, O.make_params(O.ITERATIVE_CLUSTER_FIT, O.PERCEPTUAL, False), threads=4, want_stats=True)
        gotb = block_rows_of(h_out1.numpy(), 8, W // 4, 0, sh_ // 4, sw_ // 4)
        n3, id3, worse3 = sum_over_ranks(parity_counts(0, gotb, want.reshape(-1, 8), spx, O.make_params(O.ITERATIVE_CLUSTER_FIT, O.PERCEPTUAL, False)))
        fpb = (138.0 * st.cand3 + 159.0 * st.cand4) / max(1, st.blocks)
        if rank == 0:
            configs["cfg3_bc1_iterative_8192"] = {
                "workload": f"BC1 IterativeClusterFit, PERCEPTUAL, 8192^2 noise_opaque seed 3, block-row shards x{world}",
                "ms": ms3, "mpix_s": W * H / (ms3 / 1e3) / 1e6, "e2e_ms": e3, "e2e_mpix_s": W * H / (e3 / 1e3) / 1e6,
                "parity": {"blocks_checked": n3, "identical": id3, "pct": 100.0 * id3 / max(1, n3), "worse_than_ref": worse3,
                           "sample": f"top {sh_} rows x {sw_} px of every rank's shard, host-API output"},
                "roofline": fp32_roof("cluster_setup_sorted_kernel<BC1> + cluster_lane_iter_kernel<BC1> x2", fpb, nblk, ms3,
                                      f"flops/block = 138*C3 + 159*C4 with the oracle's candidate counts on a {sw_}x{sh_} sample "
                                      f"({st.orderings3 / max(1, st.blocks):.2f} + {st.orderings4 / max(1, st.blocks):.2f} orderings per block)")}
        # restore the ClusterFit outputs for nothing below depends on them; keep buffers

        # cfg5: textures 1024^2 `smooth` + full mip chains, BC3 ClusterFit, texture t -> rank t % N, end to end from pinned host memory
        n_tex_total = args.batch_textures
        mine = list(range(rank, n_tex_total, world))
        uniq = min(len(mine), 32)
        if uniq:
            tex_pinned = [torch.from_numpy(synth.generate("smooth", 1024, 1024, 5_000_000 + mine[k]).reshape(-1)).pin_memory() for k in range(uniq)]
            size5 = L.txp_mipchain_compressed_size(2, 1024, 1024)
            ring = min(len(mine), 256)
            out_pinned = torch.empty(ring * size5, dtype=torch.uint8).pin_memory()
            outs_np = out_pinned.numpy()
            texs = [(tex_pinned[k % uniq].numpy(), 1024, 1024) for k in range(len(mine))]
            outs = [outs_np[(k % ring) * size5:(k % ring + 1) * size5] for k in range(len(mine))]
            T.compress_batch_mips(T.Format.Bc3, texs[:ring], T.Params(), n_gpus=1, outputs=outs[:ring])      # warm-up
        barrier(); t0 = time.perf_counter()
        if uniq:
            T.compress_batch_mips(T.Format.Bc3, texs, T.Params(), n_gpus=1, outputs=outs)
        barrier()
        s5 = max_over_ranks([time.perf_counter() - t0])[0]
        ok5 = 1
        if uniq:
            # per-texture bytes equal a lone call on the same texture (checked on this rank's first texture)
            ok5 = int(np.array_equal(outs[0], T.compress_mipchain(T.Format.Bc3, texs[0][0], 1024, 1024, T.Params())))
        ok5 = sum_over_ranks([1 - ok5])[0] == 0
        if rank == 0:
            pix5 = sum(w_ * h_ for w_, h_ in T.mip_levels(1024, 1024))
            configs["cfg5_batch_mips_bc3"] = {
                "workload": f"{n_tex_total} textures 1024^2 smooth + 10 mips, BC3 ClusterFit, texture t -> rank t % {world}, pinned host -> pinned host "
                            f"(mips generated on the device); {uniq} distinct textures per rank repeated, outputs into a ring of 256 pinned buffers",
                "s": s5, "textures_per_s": n_tex_total / s5, "mpix_s": n_tex_total * pix5 / s5 / 1e6,
                "h2d_bytes": n_tex_total * 4 << 20, "d2h_bytes": n_tex_total * L.txp_mipchain_compressed_size(2, 1024, 1024),
                "matches_lone_call": ok5}
            del tex_pinned, out_pinned
    if not args.no_configs and not iterative and world == 1:
        pdef = T.Params()
        # cfg2: BC3 ClusterFit 4096^2 noise_alpha seed 2, default Params
        w2 = 4096
        d2 = torch.from_numpy(synth.generate("noise_alpha", w2, w2, 2).reshape(-1)).cuda()
        o2 = torch.empty((w2 // 4) ** 2 * 16, dtype=torch.uint8, device="cuda")
        ms = timed(lambda: dev_call(T.Format.Bc3, d2, w2, w2, pdef, o2))
        configs["cfg2_bc3_cluster_4096"] = {"workload": "BC3 ClusterFit, default Params, 4096^2 noise_alpha seed 2", "ms": ms, "mpix_s": w2 * w2 / (ms / 1e3) / 1e6,
                                            "roofline": fp32_roof("cluster_setup_sorted_kernel<BC3> + cluster_lane_kernel<BC3>", FLOPS_BC3, (w2 // 4) ** 2, ms)}
        del d2, o2
        # RangeFit: cfg1 (1024^2 noise_opaque seed 1, with the CPU port beside it) and 8192^2
        prf = T.Params(T.Algorithm.RangeFit, PERC, False)
        img1 = synth.generate("noise_opaque", 1024, 1024, 1)
        d1 = torch.from_numpy(img1.reshape(-1)).cuda()
        o1 = torch.empty(256 * 256 * 8, dtype=torch.uint8, device="cuda")
        ms = timed(lambda: dev_call(T.Format.Bc1, d1, 1024, 1024, prf, o1), reps=5)
        t0 = time.perf_counter()
        want = O.compress(O.BC1, img1, 1024, 1024, O.make_params(O.RANGE_FIT, O.PERCEPTUAL, False), threads=host_threads())
        cpu_s = time.perf_counter() - t0
        configs["cfg1_bc1_rangefit_1024"] = {"workload": "BC1 RangeFit, PERCEPTUAL, 1024^2 noise_opaque seed 1", "ms": ms, "mpix_s": 1024 * 1024 / (ms / 1e3) / 1e6,
                                             "cpu_port_mpix_s": 1024 * 1024 / cpu_s / 1e6, "cpu_cores": host_threads(),
                                             "bit_exact_vs_oracle": bool(np.array_equal(o1.cpu().numpy(), want)),
                                             "roofline": hbm_roof("range_encode_kernel<BC1>", 72, 65536, ms)}
        ms = timed(lambda: dev_call(T.Format.Bc1, d_bc1, W, hs, prf, d_out1), reps=5)
        rr = hbm_roof("range_encode_kernel<BC1>", 72, nblk, ms)
        rr["fp32_issue_context"] = {"flops_per_block": 1900, "achieved_tflops": 1900.0 * nblk / (ms / 1e3) / 1e12, "frac": 1900.0 * nblk / (ms / 1e3) / 1e12 / fp32_peak}
        configs["rangefit_bc1_8192"] = {"workload": "BC1 RangeFit, PERCEPTUAL, 8192^2 noise_opaque seed 3", "ms": ms, "mpix_s": W * H / (ms / 1e3) / 1e6, "roofline": rr}
        del d1, o1
        # decode: BC1 / BC3 8192^2 device-resident + BC3 end to end through Format.decompress (pinned host)
        dev_call(T.Format.Bc3, d_bc3, W, hs, params, d_out3)
        dimg = torch.empty(W * hs * 4, dtype=torch.uint8, device="cuda")

        def dev_decode(fmt, d_in):
            _lib.check(L.txp_decompress_device(int(fmt), ctypes.c_void_p(d_in.data_ptr()), W, hs, ctypes.c_void_p(dimg.data_ptr()), dimg.numel(), ctypes.c_void_p(stream)))
        ms = timed(lambda: dev_decode(T.Format.Bc3, d_out3), reps=5)
        h_img = torch.empty(W * hs * 4, dtype=torch.uint8).pin_memory()
        torch.cuda.synchronize()
        h_out3.copy_(d_out3.cpu())
        T.Format.Bc3.decompress(h_out3.numpy(), W, hs, output=h_img.numpy())
        t0 = time.perf_counter()
        T.Format.Bc3.decompress(h_out3.numpy(), W, hs, output=h_img.numpy())
        e_ms = 1e3 * (time.perf_counter() - t0)
        sample_ok = bool(np.array_equal(h_img.numpy()[:W * 64 * 4], O.decompress(O.BC3, h_out3.numpy()[:(W // 4) * 16 * 16], W, 64)))
        configs["decode_bc3_8192"] = {"workload": "BC3 decode of the encoded 8192^2 texture", "ms": ms, "mpix_s": W * H / (ms / 1e3) / 1e6,
                                      "e2e_ms": e_ms, "e2e_mpix_s": W * H / (e_ms / 1e3) / 1e6, "e2e_api": "Format.decompress(pinned host blocks) -> pinned host rgba",
                                      "bit_exact_vs_oracle_first_64_rows": sample_ok, "roofline": hbm_roof("decode_kernel<BC3>", 80, nblk, ms)}
        del dimg, h_img
        # the same BC3 ClusterFit call with PAGEABLE (plain numpy) buffers, as a caller of the reference's &[u8] signature would make it: the library
        # stages every pipeline chunk through pinned memory with a pool of copy threads (N = 1 only: eight ranks would measure the host's memory)
        if world == 1:
            p_in = h_bc3.numpy().copy()
            p_out = np.empty(h_out3.numel(), np.uint8)
            T.Format.Bc3.compress(p_in, W, hs, params, output=p_out)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                T.Format.Bc3.compress(p_in, W, hs, params, output=p_out)
                ts.append(1e3 * (time.perf_counter() - t0))
            t0 = time.perf_counter()
            T.Format.Bc3.compress(h_bc3.numpy(), W, hs, params, output=h_out3.numpy())
            pinned_ms = 1e3 * (time.perf_counter() - t0)
            configs["pageable_bc3_cluster_8192"] = {"workload": "BC3 ClusterFit 8192^2 noise_alpha through Format.compress with pageable numpy buffers (staged by the library)",
                                                    "pageable_ms": sorted(ts)[1], "pinned_ms": pinned_ms, "same_bytes": bool(np.array_equal(p_out, h_out3.numpy()))}
            del p_in, p_out
        # cfg4: BC4 and BC5, 16384^2 r_rg seed 4 (HBM-bound by intent)
        w4 = 16384
        d4 = torch.from_numpy(synth.generate("r_rg", w4, w4, 4).reshape(-1)).cuda()
        for name, fmt, bs in (("bc4", T.Format.Bc4, 8), ("bc5", T.Format.Bc5, 16)):
            o4 = torch.empty((w4 // 4) ** 2 * bs, dtype=torch.uint8, device="cuda")
            ms = timed(lambda: dev_call(fmt, d4, w4, w4, pdef, o4), reps=5)
            got = o4[:(w4 // 4) * 16 * bs].cpu().numpy()                               # first 16 block rows against the oracle
            want = O.compress(int(fmt), synth.generate("r_rg", w4, w4, 4, y0=0, y1=64), w4, 64, O.make_params(), threads=host_threads())
            configs[f"cfg4_{name}_16384"] = {"workload": f"{name.upper()} encode, 16384^2 r_rg uniform byte noise seed 4", "ms": ms, "mpix_s": w4 * w4 / (ms / 1e3) / 1e6,
                                             "bit_exact_vs_oracle_first_64_rows": bool(np.array_equal(got, want)),
                                             "roofline": hbm_roof(f"alpha_lattice_image_kernel<{name.upper()}>", 64 + bs, (w4 // 4) ** 2, ms)}
            del o4
        del d4

    if rank == 0:
        total_pix = (1 if iterative else 2) * W * H             # BC1 (+ BC3) over the whole texture, all ranks
        ms_step = dev_ms / args.steps
        value = total_pix / (ms_step / 1e3) / 1e6
        e2e_val = total_pix / (e2e_ms / args.steps / 1e3) / 1e6
        blocks_rank = nblk                                      # rank 0's launch (max-time rank is within +-1 row)
        bc3_ms = (t1 if iterative else t3) / args.steps         # dominant kernel of the step
        achieved = FLOPS_BC3 * blocks_rank / (bc3_ms / 1e3) / 1e12
        bc3_gbs = 80.0 * blocks_rank / (bc3_ms / 1e3) / 1e9
        line = {
            "metric": METRIC if not iterative else "BC1 IterativeClusterFit Mpix/s (8192^2 synthetic RGBA, BASELINE config 3)", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("BC1 IterativeClusterFit, 8192x8192 synthetic RGBA (noise_opaque, seed 3), PERCEPTUAL weights, sharded by block rows"
                                    if iterative else
                                    "BC1+BC3 ClusterFit, 8192x8192 synthetic RGBA (BC1: noise_opaque, BC3: noise_alpha, seed 3), "
                                    "PERCEPTUAL weights, sharded by block rows"),
                       "blocks_per_format": (W // 4) * (H // 4), "parallelism": f"block-row shards x{world}, no collectives",
                       "l2": "256 MiB flush write before every timed kernel"},
            "per_format": {"bc1_mpix_s": W * H / (t1 / args.steps / 1e3) / 1e6, "bc3_mpix_s": (W * H / (t3 / args.steps / 1e3) / 1e6) if t3 else None,
                           "bc1_ms": t1 / args.steps, "bc3_ms": t3 / args.steps},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": 2 * W * H * 4, "d2h_bytes_per_step": (W // 4) * (H // 4) * 24,
                    "ms_per_step": e2e_ms / args.steps, "api": "Format.compress(pinned host rgba) -> pinned host blocks"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"kernel": "cluster_setup_sorted_kernel<BC3> + cluster_lane_kernel<BC3> (ClusterFit, 16 colours/block)", "bound": "fp32_issue",
                         "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                         "peak_source": f"derived: {props.multi_processor_count} SMs x 128 lanes x {sm_max:.0f} MHz (MEASURED_PEAKS sm_max_mhz), 1 flop per lane-instruction (no FMA contraction allowed)",
                         # measured live (txp_measure_fp32_issue: independent FMUL chains, 24 warps per SM): what the pipe actually issues
                         "peak_measured": fp32_measured, "frac_of_measured": (achieved / fp32_measured) if fp32_measured else None,
                         "flops_per_block": FLOPS_BC3,
                         # `frac` counts the ALGORITHMIC 159 fp32 ops per candidate; the kernel hoists the (i, j)-only
                         # terms and issues 126 fp32 lane-operations per candidate, so pipe utilisation is lower than frac:
                         "ops_issued_per_candidate": 126, "ops_algorithmic_per_candidate": 159,
                         "frac_pipe": NCU_FMA_PIPE_FRAC, "frac_pipe_source": NCU_FMA_PIPE_SOURCE,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one 8192^2 launch pair (ncu --set full), scaled to this rank's blocks
                         "traffic": int(NCU_TRAFFIC_BYTES_8192 * blocks_rank / 4194304), "traffic_source": NCU_TRAFFIC_SOURCE,
                         "algorithmic_bytes": 80 * blocks_rank,
                         "hbm_context": {"achieved_gbs": bc3_gbs, "peak_gbs": hbm_peak, "frac": bc3_gbs / hbm_peak}},
            "paths_agree": same,
            "parity": parity,
        }
        if multi_ok is not None:
            line["multi_matches_single"] = multi_ok
        if configs:
            line["configs"] = configs
        if iterative:
            line["roofline"] = None                             # candidate count is data dependent (orderings per block): no fixed flop figure
            line["e2e"]["h2d_bytes_per_step"] = W * H * 4; line["e2e"]["d2h_bytes_per_step"] = (W // 4) * (H // 4) * 8
        if cpu_line:
            line["cpu_baseline"] = cpu_line
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` sub-records (the other BASELINE configurations)")
    ap.add_argument("--batch-textures", type=int, default=4096, help="textures of BASELINE config 5 (configs.cfg5_batch_mips_bc3)")
    ap.add_argument("--shard-of", type=int, default=1, help="tuning aid (N=1 only): time rank 0's block-row shard of an N-rank run")
    ap.add_argument("--workload", default="cluster", choices=["cluster", "iterative"],
                    help="cluster (default, the BASELINE metric): BC1+BC3 ClusterFit; iterative: BASELINE config 3, BC1 IterativeClusterFit")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
