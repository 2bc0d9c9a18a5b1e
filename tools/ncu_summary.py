#!/usr/bin/env python3
"""Key metrics + per-region instruction counts of one kernel in an .ncu-rep (offline).
usage: ncu_summary.py <report.ncu-rep> [units_per_launch] [listing-out.txt]
units_per_launch (e.g. tiles or blocks) scales the per-region executed-instruction counts."""
import csv, io, subprocess, sys
rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d.get("Kernel Name", "")[:100])
    for k in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
              "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
              "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
              "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
              "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]:
        if k in d:
            print(f"  {k:75s} {d[k]}")
    for k, v in d.items():
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            try:
                if float(v) > 0.12:
                    print(f"    stall {k.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {float(v):.2f}")
            except ValueError:
                pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isrc, ie, ist = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
base = None
ins = []
for r in rows[hi + 1:]:
    if len(r) <= ist or not r[ia].startswith("0x"):
        continue
    a = int(r[ia], 16)
    base = a if base is None else base
    ins.append((a - base, r[isrc].strip(), int(r[ie]), int(r[ist])))
tot_s = sum(x[3] for x in ins) or 1
runs = []
for a, s, e, st in ins:
    if runs and abs(runs[-1][2] - e) <= 0.02 * max(e, runs[-1][2], 1):
        runs[-1][1] = a; runs[-1][3] += e; runs[-1][4] += 1; runs[-1][5] += st
    else:
        runs.append([a, a, e, e, 1, st])
print(f"  regions (>= 1 instr per unit, units = {units:g}):")
for r in runs:
    if r[3] / units >= 1:
        print(f"    {r[0]:05x}-{r[1]:05x} n={r[4]:4d} exec/instr={r[2]:9d} per_unit={r[3] / units:7.1f} stall_samples={100.0 * r[5] / tot_s:5.1f}%")
print(f"  total per unit {sum(x[2] for x in ins) / units:.1f}")
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write("\n".join(f"{a:05x} {e:9d} {st:6d}  {s}" for a, s, e, st in ins) + "\n")
