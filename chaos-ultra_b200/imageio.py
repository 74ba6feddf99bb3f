"""PNG export of a composed frame, the on-disk format next to compose.

Role in the reference: ImageHelpers.saveImageToFile, util/ImageHelpers.java:192-214 (BufferedImage TYPE_INT_ARGB filled
from the RGBA texture, R in the low byte, row 0 = top; GLRenderer.java:281-299 reads the texture back first).
Dependency-free writer (zlib + struct): 8-bit RGBA, non-interlaced.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def encode_png(frame: np.ndarray) -> bytes:
    """frame: (H, W) uint32, R in bits 0-7, G 8-15, B 16-23, A 24-31 (ImageHelpers.java:138-158)"""
    if frame.ndim != 2 or frame.dtype != np.uint32:
        raise ValueError("expected an (H, W) uint32 RGBA frame")
    h, w = frame.shape
    rgba = np.ascontiguousarray(frame).view(np.uint8).reshape(h, w * 4)      # little endian: bytes are R,G,B,A
    raw = np.concatenate([np.zeros((h, 1), dtype=np.uint8), rgba], axis=1).tobytes()   # filter type 0 per scanline
    return (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0)) +
            _chunk(b"IDAT", zlib.compress(raw, 6)) + _chunk(b"IEND", b""))


def save_png(path, frame: np.ndarray) -> None:
    with open(path, "wb") as f:
        f.write(encode_png(frame))


def decode_png_rgba8(data: bytes) -> np.ndarray:
    """inverse of encode_png for files it wrote (filter 0 only); used by the tests"""
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        (n,) = struct.unpack(">I", data[pos:pos + 4])
        tag, body = data[pos + 4:pos + 8], data[pos + 8:pos + 8 + n]
        if tag == b"IHDR":
            w, h = struct.unpack(">II", body[:8])
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + w * 4)
    assert (raw[:, 0] == 0).all()
    return np.ascontiguousarray(raw[:, 1:]).view(np.uint32).reshape(h, w)
