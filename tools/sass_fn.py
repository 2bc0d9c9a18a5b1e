#!/usr/bin/env python3
"""Dump the SASS of one kernel of libtexpresso_b200.so (offline).  usage: sass_fn.py <mangled-substring> [out-file]"""
import re, subprocess, sys, tempfile, pathlib
root = pathlib.Path(__file__).resolve().parent.parent
tmp = pathlib.Path(tempfile.mkdtemp())
subprocess.run(["cuobjdump", "-xelf", "all", str(root / "texpresso_b200/libtexpresso_b200.so")], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = next(tmp.glob("*.cubin"))
out = subprocess.run(["cuobjdump", "-sass", str(cubin)], capture_output=True, text=True).stdout
sec, lines = None, []
for l in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        sec = m.group(1); continue
    if sec and sys.argv[1] in sec:
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m: lines.append(f"{m.group(1)}  {m.group(2).strip()}")
text = "\n".join(lines) + "\n"
if len(sys.argv) > 2: pathlib.Path(sys.argv[2]).write_text(text)
else: sys.stdout.write(text)
