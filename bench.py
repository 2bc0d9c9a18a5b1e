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


def cpu_reference_time(crop, threads, repeats=1):
    """Times the CPU oracle (port of the reference algorithm) on a crop x crop sample of the workload.
    Returns (Mpix/s over BC1+BC3, seconds)."""
    import numpy as np
    from texpresso_b200 import synth
    from tests import oracle_lib as O                     # allowed here: cpu_baseline / --impl reference leg
    img = synth.generate("noise_alpha", W, H, SEED, y0=0, y1=crop)[:, :crop].copy()
    opaque = img.copy(); opaque[..., 3] = 255
    p = O.make_params(O.CLUSTER_FIT, O.PERCEPTUAL, False)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.compress(O.BC1, opaque, crop, crop, p, threads=threads)
        O.compress(O.BC3, img, crop, crop, p, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
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

    # ---- this rank's shard: block rows [r0, r1) -----------------------------------------------------------
    r0, r1 = T.shard_rows(H, rank, world)
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
    params = T.Params(T.Algorithm.IterativeClusterFit if iterative else T.Algorithm.ClusterFit, T.COLOUR_WEIGHTS_PERCEPTUAL, False)
    cp = params._c()
    stream = torch.cuda.current_stream().cuda_stream

    def dev_encode(fmt, d_in, d_out):
        _lib.check(L.txp_compress_device(int(fmt), ctypes.c_void_p(d_in.data_ptr()), W, hs, ctypes.byref(cp),
                                         ctypes.c_void_p(d_out.data_ptr()), d_out.numel(), ctypes.c_void_p(stream)))

    ev = lambda: torch.cuda.Event(enable_timing=True)

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
    dev_ms = (t1 + t3)
    if dist is not None:
        t = torch.tensor([dev_ms, t1, t3], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, t1, t3 = t.tolist()
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())

    # ---- end to end through the host API (pinned host in, pinned host out) --------------------------------------
    for _ in range(max(1, min(args.warmup, 2))):
        host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    if dist is not None:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item()
    clocks = sampler.stop() if rank == 0 else None

    # correctness guard: device-resident and host-API paths must agree byte for byte
    same = bool(torch.equal(d_out1.cpu(), h_out1) and (iterative or torch.equal(d_out3.cpu(), h_out3)))

    if rank == 0:
        total_pix = (1 if iterative else 2) * W * H             # BC1 (+ BC3) over the whole texture, all ranks
        ms_step = dev_ms / args.steps
        value = total_pix / (ms_step / 1e3) / 1e6
        e2e_val = total_pix / (e2e_ms / args.steps / 1e3) / 1e6
        props = torch.cuda.get_device_properties(local)
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        fp32_peak = props.multi_processor_count * 128 * sm_max * 1e6 / 1e12      # Tflop/s, non-FMA issue
        blocks_rank = nblk                                      # rank 0's launch (max-time rank is within +-1 row)
        bc3_ms = (t1 if iterative else t3) / args.steps         # dominant kernel of the step
        achieved = FLOPS_BC3 * blocks_rank / (bc3_ms / 1e3) / 1e12
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
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
                         "flops_per_block": FLOPS_BC3,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one 8192^2 launch pair (ncu --set full,
                         # profiles/ncu_lane_r01_summary.txt: setup 335.7 MB + 1204 MB, search 1254 MB + 65.3 MB; algorithmic
                         # 268.4 + 67.1 MB -- the rest is the 284 B/block scratch between the two kernels), scaled to this rank's blocks
                         "traffic": int(2859.4e6 * blocks_rank / 4194304),
                         "hbm_context": {"achieved_gbs": bc3_gbs, "peak_gbs": hbm_peak, "frac": bc3_gbs / hbm_peak}},
            "paths_agree": same,
        }
        if iterative:
            line["roofline"] = None                             # candidate count is data dependent (orderings per block): no fixed flop figure
            line["e2e"]["h2d_bytes_per_step"] = W * H * 4; line["e2e"]["d2h_bytes_per_step"] = (W // 4) * (H // 4) * 8
        if world == 1 and not args.no_cpu_baseline and not iterative:
            threads = host_threads()
            crop = 256
            mp, dt = cpu_reference_time(crop, threads)
            while dt < 4.0 and crop < 4096:
                crop *= 2
                mp, dt = cpu_reference_time(crop, threads)
            line["cpu_baseline"] = {"value": mp, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{crop}x{crop} crop of the same 8192^2 workload, BC1+BC3 ClusterFit, {dt:.1f} s",
                                    "note": "C port of the reference algorithm (oracle/txp_oracle.c); the Rust reference cannot be built here"}
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
    ap.add_argument("--workload", default="cluster", choices=["cluster", "iterative"],
                    help="cluster (default, the BASELINE metric): BC1+BC3 ClusterFit; iterative: BASELINE config 3, BC1 IterativeClusterFit")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
