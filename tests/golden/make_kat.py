#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors into tests/golden/kat.json.

Source: /root/reference/lib/src/test_data.rs (8 data sets) and lib/src/lib.rs:350-361 (size table),
lib.rs:424-431 (4x6 decode vector).  Run in the build container (the reference is not present on the
GPU box); the JSON is committed.
"""
import json, re, pathlib

SRC = pathlib.Path("/root/reference/lib/src/test_data.rs").read_text()
LIB = pathlib.Path("/root/reference/lib/src/lib.rs").read_text()


def bytes_of(body):
    return [int(x, 16) for x in re.findall(r"0x([0-9A-Fa-f]{2})", re.sub(r"//.*", "", body))]


def const_array(name):
    m = re.search(r"const %s: \[u8; [^\]]*\] = \[(.*?)\];" % name, SRC, re.S)
    return bytes_of(m.group(1))


def expand_single_to_rgb(v):
    return [c for x in v for c in (x, x, x)]


def add_alpha(rgb, alpha):
    out = []
    for i in range(16):
        out += rgb[3 * i:3 * i + 3] + [alpha[i]]
    return out


COLOUR = const_array("COLOUR_BLOCK_RGB")
GRAY = const_array("GRAY_BLOCK_LUMA")
BC3A = const_array("BC3_ALPHA_DECODED")
RAMP = const_array("LINEAR_RAMP")
FF = [0xFF] * 16


def dataset(name):
    m = re.search(r"pub const %s: TestDataSet = TestDataSet \{(.*?)\n\};" % name, SRC, re.S)
    body = m.group(1)
    enc = bytes_of(re.search(r"encoded: &\[(.*?)\],\s*decoded", body, re.S).group(1))
    dec_src = re.search(r"decoded: (.*)", body, re.S).group(1)
    return enc, dec_src


def inline_array(dec_src):
    m = re.search(r"&\[\s*((?:0x[0-9A-Fa-f]{2}[,\s]*(?://[^\n]*\n\s*)?)+)\]", dec_src)
    return bytes_of(m.group(1))


out = {"sets": {}, "source": "reference lib/src/test_data.rs + lib/src/lib.rs tests"}
for name, fmt in [("BC1_GRAY", 0), ("BC1_COLOUR", 0), ("BC2_GRAY", 1), ("BC2_COLOUR", 1), ("BC3_GRAY", 2),
                  ("BC3_COLOUR", 2), ("BC4_GRAY", 3), ("BC5_GRAY", 4)]:
    enc, d = dataset(name)
    if name in ("BC1_GRAY", "BC4_GRAY"):
        dec = add_alpha(expand_single_to_rgb(inline_array(d)), FF)
    elif name == "BC1_COLOUR":
        dec = add_alpha(COLOUR, FF)
    elif name == "BC2_GRAY":
        dec = add_alpha(expand_single_to_rgb(GRAY), RAMP)
    elif name == "BC2_COLOUR":
        dec = add_alpha(COLOUR, RAMP)
    elif name == "BC3_GRAY":
        dec = add_alpha(expand_single_to_rgb(GRAY), BC3A)
    elif name == "BC3_COLOUR":
        dec = add_alpha(COLOUR, BC3A)
    elif name == "BC5_GRAY":
        dec = add_alpha(inline_array(d), FF)
    assert len(dec) == 64 and len(enc) in (8, 16), name
    out["sets"][name] = {"format": fmt, "encoded": enc, "decoded": dec}

# lib.rs:424-431
m = re.search(r"let encoded = \[(.*?)\];", LIB, re.S)
out["decode_4x6"] = {"format": 0, "width": 4, "height": 6, "encoded": bytes_of(m.group(1)), "pixel": [0x7F, 0x7F, 0x7F, 0xFF]}
# lib.rs:350-361
out["sizes"] = [[int(f) - 1, int(w), int(h), int(s)] for f, w, h, s in
                re.findall(r"Format::Bc(\d)\.compressed_size\((\d+), (\d+)\), (\d+)\)", LIB)]
assert len(out["sizes"]) == 10
p = pathlib.Path(__file__).with_name("kat.json")
p.write_text(json.dumps(out, indent=1))
print("wrote", p, {k: len(v["encoded"]) for k, v in out["sets"].items()})
