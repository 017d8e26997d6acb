"""Minimal CPTV v2 writer for the decoder tests (the layout of SURVEY.md section 8c, as recovered from the reference's
tests/clips/*.cptv): gzip stream, b"CPTV", version 2, an 'H' section and one 'F' section + payload per frame."""
import gzip
import struct

import numpy as np


def _field(code, data):
    return bytes([len(data)]) + code.encode() + data


def _snake(width, height):
    idx = np.arange(width * height, dtype=np.int64).reshape(height, width)
    idx[1::2] = idx[1::2, ::-1]
    return idx.reshape(-1)


def pack_deltas(deltas, bit_width):
    """Two's-complement ``bit_width``-bit values, MSB first."""
    vals = np.asarray(deltas, dtype=np.int64) & ((1 << bit_width) - 1)
    bits = ((vals[:, None] >> np.arange(bit_width - 1, -1, -1)) & 1).astype(np.uint8).reshape(-1)
    return np.packbits(bits).tobytes()


def write_cptv(path, frames, model=None, background_first=False, force_bit_width=None):
    """frames: (T, H, W) uint16.  Returns the bit width used per frame."""
    frames = np.asarray(frames)
    T, H, W = frames.shape
    snake = _snake(W, H)
    out = bytearray(b"CPTV\x02")
    fields = [_field("X", struct.pack("<I", W)), _field("Y", struct.pack("<I", H)), _field("T", struct.pack("<Q", 1_600_000_000_000_000)),
              _field("C", b"\x01"), _field("Z", b"\x09")]
    if model:
        fields.append(_field("E", model.encode()))
    if background_first:
        fields.append(_field("g", b"\x01"))
    out += b"H" + bytes([len(fields)]) + b"".join(fields)
    prev = np.zeros(W * H, dtype=np.int64)
    widths = []
    for t in range(T):
        cur = frames[t].astype(np.int64).reshape(-1)
        change = (cur - prev)[snake]
        prev = cur
        deltas = np.diff(change)
        need = 1
        if len(deltas):
            lo, hi = int(deltas.min()), int(deltas.max())
            while not (-(1 << (need - 1)) <= lo and hi < (1 << (need - 1))):
                need += 1
        w = max(need, force_bit_width or 1)
        widths.append(w)
        payload = struct.pack("<i", int(change[0])) + pack_deltas(deltas, w)
        ff = [_field("t", struct.pack("<I", 10_000_000 + t * 111)), _field("c", struct.pack("<I", 0)), _field("w", bytes([w])),
              _field("f", struct.pack("<I", len(payload)))]
        if background_first and t == 0:
            ff.append(_field("g", b"\x01"))
        out += b"F" + bytes([len(ff)]) + b"".join(ff) + payload
    with gzip.open(str(path), "wb") as f:
        f.write(bytes(out))
    return widths
