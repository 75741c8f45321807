"""Host-side image files for the output side: the reference writes its screenshots with stb
(saveImage, common/common_host.cpp:2715-2723) from the packed RGBA8 image; this is the dependency-free equivalent."""
from __future__ import annotations

import struct
import zlib

import numpy as np


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def write_png(path: str, rgba8: np.ndarray) -> None:
    """rgba8: [H, W] uint32 packed R | G << 8 | B << 16 | A << 24 (gfx_present_launch / engine.Context.present)"""
    img = np.ascontiguousarray(rgba8, dtype="<u4")
    h, w = img.shape
    rows = img.view(np.uint8).reshape(h, w * 4)
    raw = b"".join(b"\x00" + rows[y].tobytes() for y in range(h))   # filter type 0 on every scanline
    png = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) + \
        _chunk(b"IDAT", zlib.compress(raw, 6)) + _chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(png)


def read_png(path: str) -> np.ndarray:
    """inverse of write_png for 8-bit RGBA files with filter type 0 (round-trip tests)"""
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == (zlib.crc32(tag + body) & 0xFFFFFFFF)
        if tag == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", body[:10])
            assert (depth, ctype) == (8, 6)
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 4 * w)
    assert not raw[:, 0].any()
    return np.ascontiguousarray(raw[:, 1:]).view("<u4").reshape(h, w)
