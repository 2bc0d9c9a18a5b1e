"""DDS container and CLI mirror.  CPU: header layout, read/write round trip, argument
parsing, image ingest.  GPU: file in -> DDS -> file out equals the oracle's encode/decode of the same pixels."""
import io, pathlib, struct
import numpy as np
import pytest

from texpresso_b200 import dds, cli


def test_header_layout():
    h = dds.header_bytes(2, 640, 480)
    assert len(h) == 148 and h[:4] == b"DDS "
    size, flags, height, width, linear, depth, mips = struct.unpack_from("<7I", h, 4)
    assert (size, height, width, depth, mips) == (124, 480, 640, 0, 0)
    assert flags == 0x1 | 0x2 | 0x4 | 0x1000 | 0x80000 and linear == 160 * 120 * 16
    assert struct.unpack_from("<2I4s", h, 76) == (32, 4, b"DX10")
    assert struct.unpack_from("<I", h, 108)[0] == 0x1000                          # caps = TEXTURE
    assert struct.unpack_from("<5I", h, 128) == (78, 3, 0, 1, 1)                  # BC3_UNorm_sRGB, Texture2D, straight alpha
    assert struct.unpack_from("<5I", dds.header_bytes(0, 4, 4), 128) == (72, 3, 0, 1, 2)   # BC1: premultiplied (main.rs:143-147)
    hm = dds.header_bytes(4, 16, 16, mip_levels=5)
    assert struct.unpack_from("<I", hm, 8)[0] & 0x20000 and struct.unpack_from("<I", hm, 28)[0] == 5


@pytest.mark.parametrize("fmt", range(5))
def test_round_trip(fmt):
    w, h = 20, 12
    n = dds._level_size(fmt, w, h)
    data = bytes(range(256)) * (n // 256 + 1)
    buf = io.BytesIO()
    dds.write_dds(buf, fmt, w, h, data[:n])
    buf.seek(0)
    assert dds.read_dds(buf) == (fmt, w, h, data[:n], 1)
    with pytest.raises(ValueError):
        dds.write_dds(io.BytesIO(), fmt, w, h, data[:n - 1])


def test_legacy_fourcc_and_rejections():
    h = bytearray(dds.header_bytes(0, 8, 8)[:128])
    h[84:88] = b"DXT5"
    assert dds.read_dds(io.BytesIO(bytes(h) + b"\0" * 64))[0] == 2                # main.rs:241-248
    h[84:88] = b"ATI2"
    with pytest.raises(ValueError):
        dds.read_dds(io.BytesIO(bytes(h) + b"\0" * 64))
    bad = bytearray(dds.header_bytes(0, 8, 8)); struct.pack_into("<I", bad, 128, 71)     # BC1_UNorm (non-sRGB): reference panics
    with pytest.raises(ValueError):
        dds.read_dds(io.BytesIO(bytes(bad)))


def test_cli_arguments_and_ingest(tmp_path):
    from PIL import Image
    a = cli.build_parser().parse_args(["compress", "x.png", "-f", "BC3", "-p", "quality", "-w", "1", "1", "1", "--weigh-colour-by-alpha"])
    p = cli.params_from_args(a)
    assert (int(p.algorithm), p.weights, p.weigh_colour_by_alpha) == (2, (1.0, 1.0, 1.0), True)
    assert int(cli.params_from_args(cli.build_parser().parse_args(["compress", "x.png", "-f", "bc1"])).algorithm) == 1
    grey = np.arange(12, dtype=np.uint8).reshape(3, 4) * 20
    Image.fromarray(grey, "L").save(tmp_path / "g.png")
    rgba, w, h = cli.read_image(tmp_path / "g.png")
    assert (w, h) == (4, 3) and np.array_equal(rgba[..., 0], grey) and np.array_equal(rgba[..., 2], grey) and (rgba[..., 3] == 255).all()
    la = np.dstack([grey, 255 - grey])
    Image.fromarray(la, "LA").save(tmp_path / "la.png")
    rgba, _, _ = cli.read_image(tmp_path / "la.png")
    assert np.array_equal(rgba[..., 1], grey) and np.array_equal(rgba[..., 3], 255 - grey)
    # read_pixels keeps the file layout (expanded on the device by compress_pixels); expand_pixels is its host statement
    import texpresso_b200 as T
    px, w, h = cli.read_pixels(tmp_path / "g.png")
    assert px.shape == (3, 4, 1) and np.array_equal(T.expand_pixels(px, w, h), cli.read_image(tmp_path / "g.png")[0])
    px, w, h = cli.read_pixels(tmp_path / "la.png")
    assert px.shape == (3, 4, 2) and np.array_equal(T.expand_pixels(px, w, h), rgba)
    Image.fromarray(np.dstack([grey, grey // 2, 255 - grey]), "RGB").save(tmp_path / "rgb.png")
    px, w, h = cli.read_pixels(tmp_path / "rgb.png")
    assert px.shape == (3, 4, 3) and np.array_equal(T.expand_pixels(px, w, h), cli.read_image(tmp_path / "rgb.png")[0])


@pytest.mark.gpu
@pytest.mark.parametrize("fmt,profile,alg", [("bc1", "speed", 0), ("bc3", "balanced", 1), ("bc5", "quality", 2)])
def test_cli_compress_decompress_files(tmp_path, fmt, profile, alg):
    from PIL import Image
    from texpresso_b200 import synth
    from tests import oracle_lib as O
    w, h = 50, 38
    img = synth.generate("smooth", w, h, seed=5)
    Image.fromarray(img, "RGBA").save(tmp_path / "in.png")
    assert cli.main(["compress", str(tmp_path / "in.png"), "-f", fmt, "-p", profile, "-o", str(tmp_path / "out.dds")]) == 0
    f, fw, fh, data, levels = dds.read_dds(tmp_path / "out.dds")
    want = O.compress(f, img, w, h, O.make_params(alg, O.PERCEPTUAL, False))
    assert (fw, fh, levels) == (w, h, 1) and np.array_equal(np.frombuffer(data, np.uint8), want)
    assert cli.main(["decompress", str(tmp_path / "out.dds"), "-o", str(tmp_path / "back.png")]) == 0
    back = np.asarray(Image.open(tmp_path / "back.png").convert("RGBA"))
    assert np.array_equal(back.reshape(-1), O.decompress(f, want, w, h))
    # mip chain extension
    assert cli.main(["compress", str(tmp_path / "in.png"), "-f", fmt, "-p", profile, "--mips", "-o", str(tmp_path / "m.dds")]) == 0
    f, fw, fh, data, levels = dds.read_dds(tmp_path / "m.dds")
    assert levels == 6 and len(data) == sum(dds._level_size(f, a, b) for a, b in dds.mip_chain_dims(w, h, 6))
