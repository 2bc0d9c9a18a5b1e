"""Command line mirror of the reference's `texpresso` binary (cli/src/main.rs:48-196), on the CUDA path:

    python -m texpresso_b200.cli compress INFILE -f bc1|bc2|bc3|bc4|bc5 [-p speed|balanced|quality]
                                 [--weigh-colour-by-alpha] [-w R G B] [-o OUT.dds] [--mips]
    python -m texpresso_b200.cli decompress INFILE.dds [-o OUT.png]

Image ingest follows cli/src/image/png.rs:30-71 and jpeg.rs:29-57 (grey -> (l,l,l,255), grey+alpha -> (l,l,l,a),
RGB -> (r,g,b,255), 16-bit stripped to 8) through Pillow.  `--mips` (not in the reference) adds the device-generated
mip chain to the DDS."""
import argparse, pathlib, sys
import numpy as np

from . import Algorithm, Format, Params, COLOUR_WEIGHTS_PERCEPTUAL, compress_mipchain, compress_pixels, mip_levels
from . import dds

PROFILES = {"speed": Algorithm.RangeFit, "balanced": Algorithm.ClusterFit, "quality": Algorithm.IterativeClusterFit}   # main.rs:198-206
FORMATS = {"bc1": Format.Bc1, "bc2": Format.Bc2, "bc3": Format.Bc3, "bc4": Format.Bc4, "bc5": Format.Bc5}


def build_parser():
    ap = argparse.ArgumentParser(prog="texpresso_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    c = sub.add_parser("compress", help="Compress a PNG or JPEG file to DDS")
    c.add_argument("infile", metavar="INFILE")
    c.add_argument("-o", "--output", dest="outfile")
    c.add_argument("-f", "--format", required=True, type=str.lower, choices=sorted(FORMATS))
    c.add_argument("-p", "--profile", default="balanced", type=str.lower, choices=sorted(PROFILES))
    c.add_argument("--weigh-colour-by-alpha", action="store_true")
    c.add_argument("-w", "--weights", type=float, nargs="*", default=[])
    c.add_argument("--mips", action="store_true", help="also encode the full mip chain (extension)")
    d = sub.add_parser("decompress", help="Decompress a DDS file to PNG")
    d.add_argument("infile", metavar="INFILE")
    d.add_argument("-o", "--output", dest="outfile")
    return ap


def read_image(path):
    """-> (rgba uint8 (h, w, 4), w, h)"""
    from PIL import Image
    ext = pathlib.Path(path).suffix.lower()
    if ext not in (".png", ".jpg", ".jpeg"):
        raise SystemExit("Unrecognized image format. Supported formats are PNG and JPEG")      # main.rs:136
    im = Image.open(path)
    if im.mode in ("I;16", "I;16B", "I"):
        im = im.point(lambda v: v >> 8).convert("L")                                             # STRIP_16
    rgba = np.asarray(im.convert("RGBA"), dtype=np.uint8)
    return np.ascontiguousarray(rgba), im.width, im.height


def read_pixels(path):
    """-> (pixels uint8 (h, w, c), w, h) in the decoded file layout: c = 1 (gray), 2 (gray + alpha), 3 (RGB) or 4.  The
    expansion to RGBA8 that cli/src/image/png.rs:47-62 does on the host is left to the device (compress_pixels)."""
    from PIL import Image
    ext = pathlib.Path(path).suffix.lower()
    if ext not in (".png", ".jpg", ".jpeg"):
        raise SystemExit("Unrecognized image format. Supported formats are PNG and JPEG")      # main.rs:136
    im = Image.open(path)
    if im.mode in ("I;16", "I;16B", "I"):
        im = im.point(lambda v: v >> 8).convert("L")                                             # STRIP_16
    if im.mode not in ("L", "LA", "RGB", "RGBA"):
        im = im.convert("RGBA")                                                                  # EXPAND (palette, 1-bit, CMYK ...)
    px = np.asarray(im, dtype=np.uint8).reshape(im.height, im.width, -1)
    return np.ascontiguousarray(px), im.width, im.height


def params_from_args(args):
    if not args.weights:
        w = COLOUR_WEIGHTS_PERCEPTUAL
    elif len(args.weights) == 3:
        w = tuple(args.weights)
    else:
        raise SystemExit("Weights must have 3 values")                                            # main.rs:109
    return Params(PROFILES[args.profile], w, args.weigh_colour_by_alpha)


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.cmd == "compress":
        fmt = FORMATS[args.format]
        out = args.outfile or str(pathlib.Path(pathlib.Path(args.infile).name).with_suffix(".dds"))
        params = params_from_args(args)
        if args.mips:
            rgba, w, h = read_image(args.infile)
            data = compress_mipchain(fmt, rgba, w, h, params)
            dds.write_dds(out, fmt, w, h, data, mip_levels=len(mip_levels(w, h)))
        else:
            pixels, w, h = read_pixels(args.infile)          # 1-4 bytes per pixel; expanded to RGBA8 on the device
            data = compress_pixels(fmt, pixels, w, h, params)
            dds.write_dds(out, fmt, w, h, data)
    else:
        from PIL import Image
        out = args.outfile or str(pathlib.Path(pathlib.Path(args.infile).name).with_suffix(".png"))
        f, w, h, data, _levels = dds.read_dds(args.infile)
        fmt = Format(f)
        need = fmt.compressed_size(w, h)
        px = fmt.decompress(np.frombuffer(data[:need], dtype=np.uint8), w, h)
        Image.fromarray(px.reshape(h, w, 4), "RGBA").save(out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
