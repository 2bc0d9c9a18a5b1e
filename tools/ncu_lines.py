#!/usr/bin/env python3
"""Join an ncu SASS-level profile (--page source --csv) with nvdisasm line info to attribute executed
warp-instructions to CUDA source lines.  usage: ncu_lines.py <rep> <kernel-substring> <mangled-substr> <blocks>"""
import csv, collections, re, subprocess, sys, pathlib, tempfile

rep, kname, mangled, nb = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
root = pathlib.Path(__file__).resolve().parent.parent
tmp = pathlib.Path(tempfile.mkdtemp())
subprocess.run(["cuobjdump", "-xelf", "all", str(root / "texpresso_b200/libtexpresso_b200.so")], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = next(tmp.glob("*.cubin"))
dis = subprocess.run(["nvdisasm", "--print-line-info", str(cubin)], capture_output=True, text=True).stdout.splitlines()
# line info per instruction offset inside the kernel's .text section
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l)
lines_by_off = {}
cur = None
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (pathlib.Path(m.group(1)).name, int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines_by_off[int(m.group(1), 16)] = (cur, m.group(2))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
ker, hdr, cur = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = r[1]; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if cur and kname in cur and r:
        ker.append(r)
ia, ist = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
base = int(ker[0][0], 16)
agg = collections.defaultdict(lambda: [0.0, 0])
for r in ker:
    off = int(r[0], 16) - base
    src, _ = lines_by_off.get(off, (None, None))
    agg[src][0] += int(r[ia]) / nb
    agg[src][1] += int(r[ist])
tot = sum(v[0] for v in agg.values()); tst = sum(v[1] for v in agg.values())
print(f"total warp-instr/block {tot:.0f}; stall samples {tst}")
srcs = {}
for (key, v) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 60]:
    text = ""
    if key:
        f = root / "texpresso_b200/csrc" / key[0]
        if f.exists():
            if f not in srcs: srcs[f] = f.read_text().splitlines()
            text = srcs[f][key[1] - 1].strip()[:100]
    print(f"{v[0]:8.1f} {100*v[0]/tot:5.1f}%  stall {100*v[1]/max(tst,1):5.1f}%  {key}  {text}")
