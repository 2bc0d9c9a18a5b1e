"""DDS container read/write for the five BCn formats (reference cli/src/main.rs:143-164,
:174-193, :220-248).

The reference delegates the container to the un-vendored crate `ddsfile = "0.5"` (cli/Cargo.toml); its exact
output cannot be produced here (no Rust toolchain), so this restates the published DDS layout the way
`Dds::new_dxgi` fills it: magic, 124-byte DDS_HEADER (CAPS|HEIGHT|WIDTH|PIXELFORMAT|LINEARSIZE [+MIPMAPCOUNT]),
FourCC "DX10" pixel format, 20-byte DDS_HEADER_DXT10 with the DXGI format, Texture2D, array size 1 and the alpha
mode (premultiplied for BC1, straight otherwise -- main.rs:143-147).  Parity with ddsfile is unpinned.

Reading accepts what the reference accepts: DX10 headers with the five DXGI formats below (Texture2D only) and
legacy FourCC DXT1 / DXT3 / DXT5 (main.rs:230-248).  Unlike the reference this module can also carry mip levels
(the reference writes none, main.rs:153)."""
import struct

MAGIC = 0x20534444  # "DDS "
DDSD_CAPS, DDSD_HEIGHT, DDSD_WIDTH, DDSD_PIXELFORMAT, DDSD_MIPMAPCOUNT, DDSD_LINEARSIZE = 0x1, 0x2, 0x4, 0x1000, 0x20000, 0x80000
DDPF_FOURCC = 0x4
DDSCAPS_COMPLEX, DDSCAPS_TEXTURE, DDSCAPS_MIPMAP = 0x8, 0x1000, 0x400000
D3D10_RESOURCE_DIMENSION_TEXTURE2D = 3
ALPHA_MODE_STRAIGHT, ALPHA_MODE_PREMULTIPLIED = 1, 2

# main.rs:220-228 : Format -> DxgiFormat (sRGB variants for BC1-3)
DXGI_OF_FORMAT = {0: 72, 1: 75, 2: 78, 3: 80, 4: 83}   # BC1_UNorm_sRGB, BC2_UNorm_sRGB, BC3_UNorm_sRGB, BC4_UNorm, BC5_UNorm
FORMAT_OF_DXGI = {v: k for k, v in DXGI_OF_FORMAT.items()}            # main.rs:230-239
FORMAT_OF_FOURCC = {b"DXT1": 0, b"DXT3": 1, b"DXT5": 2}                 # main.rs:241-248
BLOCK_SIZE = {0: 8, 1: 16, 2: 16, 3: 8, 4: 16}


def _level_size(fmt, w, h):
    return ((w + 3) // 4) * ((h + 3) // 4) * BLOCK_SIZE[fmt]


def mip_chain_dims(width, height, levels):
    out, w, h = [], width, height
    for _ in range(levels):
        out.append((w, h))
        w, h = max(1, w // 2), max(1, h // 2)
    return out


def header_bytes(fmt, width, height, mip_levels=None):
    """magic + DDS_HEADER + DDS_HEADER_DXT10 (148 bytes)."""
    fmt = int(fmt)
    flags = DDSD_CAPS | DDSD_HEIGHT | DDSD_WIDTH | DDSD_PIXELFORMAT | DDSD_LINEARSIZE
    caps = DDSCAPS_TEXTURE
    if mip_levels and mip_levels > 1:
        flags |= DDSD_MIPMAPCOUNT
        caps |= DDSCAPS_COMPLEX | DDSCAPS_MIPMAP
    hdr = struct.pack("<I", MAGIC)
    hdr += struct.pack("<7I", 124, flags, height, width, _level_size(fmt, width, height), 0, mip_levels or 0)
    hdr += b"\0" * 44                                                       # reserved1[11]
    hdr += struct.pack("<2I4s5I", 32, DDPF_FOURCC, b"DX10", 0, 0, 0, 0, 0)  # DDS_PIXELFORMAT
    hdr += struct.pack("<5I", caps, 0, 0, 0, 0)                             # caps, caps2, caps3, caps4, reserved2
    alpha_mode = ALPHA_MODE_PREMULTIPLIED if fmt == 0 else ALPHA_MODE_STRAIGHT
    hdr += struct.pack("<5I", DXGI_OF_FORMAT[fmt], D3D10_RESOURCE_DIMENSION_TEXTURE2D, 0, 1, alpha_mode)
    assert len(hdr) == 148
    return hdr


def write_dds(fileobj_or_path, fmt, width, height, data, mip_levels=None):
    data = bytes(data)
    dims = mip_chain_dims(width, height, mip_levels or 1)
    expect = sum(_level_size(int(fmt), w, h) for w, h in dims)
    if len(data) != expect:
        raise ValueError(f"data holds {len(data)} bytes, the {len(dims)} level(s) need {expect}")
    blob = header_bytes(fmt, width, height, mip_levels) + data
    if hasattr(fileobj_or_path, "write"):
        fileobj_or_path.write(blob)
    else:
        with open(fileobj_or_path, "wb") as f:
            f.write(blob)


def read_dds(fileobj_or_path):
    """-> (format, width, height, data bytes of all levels, mip level count)"""
    if hasattr(fileobj_or_path, "read"):
        blob = fileobj_or_path.read()
    else:
        with open(fileobj_or_path, "rb") as f:
            blob = f.read()
    if len(blob) < 128 or struct.unpack_from("<I", blob, 0)[0] != MAGIC:
        raise ValueError("not a DDS file")
    size, flags, height, width, _linear, _depth, mips = struct.unpack_from("<7I", blob, 4)
    if size != 124:
        raise ValueError("bad DDS header size")
    pf_size, pf_flags, fourcc = struct.unpack_from("<2I4s", blob, 76)
    off = 128
    if pf_flags & DDPF_FOURCC and fourcc == b"DX10":
        dxgi, dim, _misc, _array, _misc2 = struct.unpack_from("<5I", blob, 128)
        off = 148
        if dim != D3D10_RESOURCE_DIMENSION_TEXTURE2D:
            raise ValueError("Only images with resource dimension Texture2D are supported")   # main.rs:178-180
        if dxgi not in FORMAT_OF_DXGI:
            raise ValueError("Unsupported DXGI format!")                                         # main.rs:237
        fmt = FORMAT_OF_DXGI[dxgi]
    elif pf_flags & DDPF_FOURCC and fourcc in FORMAT_OF_FOURCC:
        fmt = FORMAT_OF_FOURCC[fourcc]
    else:
        raise ValueError("Unsupported D3D format!")                                              # main.rs:246
    levels = mips if (flags & DDSD_MIPMAPCOUNT and mips > 1) else 1
    return fmt, width, height, blob[off:], levels
