#!/usr/bin/env python3
"""List the loops of a kernel in libtexpresso_b200.so with opcode histograms (offline, no GPU).
usage: sass_loops.py <mangled-substring> [min_packed]"""
import re, collections, subprocess, sys, tempfile, pathlib
root = pathlib.Path(__file__).resolve().parent.parent
tmp = pathlib.Path(tempfile.mkdtemp())
subprocess.run(["cuobjdump", "-xelf", "all", str(root / "texpresso_b200/libtexpresso_b200.so")], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = next(tmp.glob("*.cubin"))
out = subprocess.run(["cuobjdump", "-sass", str(cubin)], capture_output=True, text=True).stdout
sec, ins = None, []
for l in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        sec = m.group(1); continue
    if sec and sys.argv[1] in sec:
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
print("kernel instrs", len(ins))
minp = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for a, t in ins:
    m = re.search(r"BRA (0x[0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        tgt = int(m.group(1), 16)
        body = [x for x in ins if tgt <= x[0] <= a]
        c = collections.Counter()
        for _, tt in body:
            op = tt.split()[1] if tt.startswith("@") else tt.split()[0]
            c[op.split(".")[0]] += 1
        npk = c["FFMA2"] + c["FADD2"] + c["FMUL2"]
        if npk >= minp:
            print(hex(tgt), "->", hex(a), "body", len(body), dict(c.most_common(40)))
