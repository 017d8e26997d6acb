"""CPTV v2 decoder (host side, K0 in SURVEY.md §8a).

Stands in for the third-party ``cptv_rs_python_bindings.CptvReader`` (python-cptv 0.0.8,
Rust; not in the reference tree) that the reference calls at
``src/track/cliptrackextractor.py:108-129,160-165``.  The layout was recovered by
decoding ``tests/clips/*.cptv``:

* the file is one gzip stream: ``b"CPTV"``, a version byte (2), then sections;
* a section is a type char (``H`` header / ``F`` frame) and a field-count byte,
  followed by fields ``len:u8, code:char, data[len]``;
* a frame section is followed by ``f`` payload bytes: an int32 LE start value and then
  ``W*H-1`` two's-complement deltas of ``w`` bits each, packed MSB first.  The deltas are
  cumulatively summed along a boustrophedon scan (odd rows right to left) and the
  result is added to the previous frame.

Decoder parity is pinned end-to-end only (``tests/clips/possum.txt``), as the reference has
no decoder test of its own.
"""
import gzip
import struct

import numpy as np

_MAGIC = b"CPTV"


class CptvHeader:
    """Clip header; attribute names follow what the reference reads from the Rust reader."""

    def __init__(self):
        self.timestamp = None  # microseconds since epoch
        self.x_resolution = 0
        self.y_resolution = 0
        self.compression = 0
        self.device_name = None
        self.model = None
        self.brand = None
        self.firmware_version = None
        self.device_id = None
        self.fps = None
        self.serial_number = None
        self.preview_secs = None
        self.motion_config = None
        self.latitude = None
        self.longitude = None
        self.loc_timestamp = None
        self.altitude = None
        self.accuracy = None
        self.has_background_frame = False
        self.total_frames = None
        self.min_value = None
        self.max_value = None


class CptvFrame:
    """One decoded frame (duck-type consumed by ``ClipTrackExtractor.process_frame``)."""

    __slots__ = (
        "pix",
        "time_on",
        "last_ffc_time",
        "temp_c",
        "last_ffc_temp_c",
        "background_frame",
        "frame_temp_c",
    )

    def __init__(self, pix, time_on, last_ffc_time, temp_c, last_ffc_temp_c, background_frame):
        self.pix = pix
        self.time_on = time_on
        self.last_ffc_time = last_ffc_time
        self.temp_c = temp_c
        self.last_ffc_temp_c = last_ffc_temp_c
        self.background_frame = background_frame
        self.frame_temp_c = temp_c


def _u(fmt, data):
    return struct.unpack("<" + fmt, data)[0]


_HEADER_FIELDS = {
    "T": ("timestamp", lambda d: _u("Q", d)),
    "X": ("x_resolution", lambda d: _u("I", d)),
    "Y": ("y_resolution", lambda d: _u("I", d)),
    "C": ("compression", lambda d: d[0]),
    "D": ("device_name", lambda d: d.decode("utf-8", "replace")),
    "E": ("model", lambda d: d.decode("utf-8", "replace")),
    "B": ("brand", lambda d: d.decode("utf-8", "replace")),
    "V": ("firmware_version", lambda d: d.decode("utf-8", "replace")),
    "N": ("device_id", lambda d: _u("I", d)),
    "Z": ("fps", lambda d: d[0]),
    "I": ("serial_number", lambda d: _u("I", d)),
    "P": ("preview_secs", lambda d: d[0]),
    "M": ("motion_config", lambda d: d.decode("utf-8", "replace")),
    "L": ("latitude", lambda d: _u("f", d)),
    "O": ("longitude", lambda d: _u("f", d)),
    "S": ("loc_timestamp", lambda d: _u("Q", d)),
    "A": ("altitude", lambda d: _u("f", d)),
    "U": ("accuracy", lambda d: _u("f", d)),
    "g": ("has_background_frame", lambda d: d[0] != 0),
    "Q": ("min_value", lambda d: _u("H", d)),
    "K": ("max_value", lambda d: _u("H", d)),
    "J": ("total_frames", lambda d: _u("H", d)),
}


def _snake_index(width, height):
    """Flat pixel index visited at each step of the boustrophedon scan."""
    idx = np.arange(width * height, dtype=np.int64).reshape(height, width)
    idx[1::2] = idx[1::2, ::-1]
    return idx.reshape(-1)


def unpack_deltas(payload, bit_width, count):
    """``count`` signed ``bit_width``-bit integers packed MSB first → int64 array."""
    if bit_width == 8:
        return np.frombuffer(payload, dtype=np.int8, count=count).astype(np.int64)
    if bit_width == 16:
        return np.frombuffer(payload, dtype=">i2", count=count).astype(np.int64)
    nbytes = (count * bit_width + 7) // 8
    bits = np.unpackbits(np.frombuffer(payload, dtype=np.uint8, count=nbytes))[: count * bit_width]
    bits = bits.reshape(count, bit_width).astype(np.int64)
    weights = 1 << np.arange(bit_width - 1, -1, -1, dtype=np.int64)
    vals = bits @ weights
    sign = 1 << (bit_width - 1)
    return np.where(vals >= sign, vals - (1 << bit_width), vals)


class CptvReader:
    """``CptvReader(path).get_header()`` / ``.next_frame()`` → frame or ``None`` at EOF."""

    def __init__(self, path):
        with gzip.open(str(path), "rb") as f:
            self._buf = f.read()
        if self._buf[:4] != _MAGIC:
            raise ValueError("{}: not a CPTV file".format(path))
        self.version = self._buf[4]
        if self.version != 2:
            raise ValueError("{}: unsupported CPTV version {}".format(path, self.version))
        self._pos = 5
        self._header = None
        self._prev = None
        self._snake = None
        self._read_header()

    # -- section parsing -------------------------------------------------------------
    def _read_fields(self, expected):
        buf = self._buf
        if self._pos >= len(buf):
            return None
        kind = chr(buf[self._pos])
        if kind != expected:
            raise ValueError("expected section {!r}, found {!r}".format(expected, kind))
        count = buf[self._pos + 1]
        pos = self._pos + 2
        fields = {}
        for _ in range(count):
            length = buf[pos]
            code = chr(buf[pos + 1])
            fields[code] = buf[pos + 2 : pos + 2 + length]
            pos += 2 + length
        self._pos = pos
        return fields

    def _read_header(self):
        fields = self._read_fields("H")
        header = CptvHeader()
        for code, data in fields.items():
            spec = _HEADER_FIELDS.get(code)
            if spec is not None:
                setattr(header, spec[0], spec[1](data))
        self._header = header
        self._snake = _snake_index(header.x_resolution, header.y_resolution)
        self._prev = np.zeros(header.x_resolution * header.y_resolution, dtype=np.int64)

    def get_header(self):
        return self._header

    def next_frame(self):
        fields = self._read_fields("F")
        if fields is None:
            return None
        width, height = self._header.x_resolution, self._header.y_resolution
        bit_width = fields["w"][0]
        size = _u("I", fields["f"])
        payload = self._buf[self._pos : self._pos + size]
        self._pos += size
        n = width * height
        deltas = np.empty(n, dtype=np.int64)
        deltas[0] = _u("i", payload[:4])
        deltas[1:] = unpack_deltas(payload[4:], bit_width, n - 1)
        change = np.cumsum(deltas)
        cur = self._prev.copy()
        cur[self._snake] += change
        self._prev = cur
        pix = cur.astype(np.uint16).reshape(height, width)
        time_on = _u("I", fields["t"]) if "t" in fields else None
        last_ffc = _u("I", fields["c"]) if "c" in fields else None
        temp_c = _u("f", fields["a"]) if "a" in fields else None
        ffc_temp = _u("f", fields["b"]) if "b" in fields else None
        background = bool(fields["g"][0]) if "g" in fields else False
        return CptvFrame(pix, time_on, last_ffc, temp_c, ffc_temp, background)

    def index_frames(self):
        """Walk every frame section from the start of the file WITHOUT decoding pixels.

        Returns ``(stream bytes, table, frames)``: the inflated stream, one ``(payload_offset, bit_width)`` row per
        frame for ``cpt_cptv_decode`` and the frames' metadata as ``CptvFrame`` objects with ``pix = None``."""
        saved = self._pos
        try:
            self._pos = 5
            self._read_fields("H")
            table, frames = [], []
            while True:
                fields = self._read_fields("F")
                if fields is None:
                    break
                size = _u("I", fields["f"])
                if self._pos + size > len(self._buf):
                    raise ValueError("CPTV frame section at byte {} is truncated".format(self._pos))
                table.append((self._pos, fields["w"][0]))
                self._pos += size
                frames.append(CptvFrame(
                    None, _u("I", fields["t"]) if "t" in fields else None, _u("I", fields["c"]) if "c" in fields else None,
                    _u("f", fields["a"]) if "a" in fields else None, _u("f", fields["b"]) if "b" in fields else None,
                    bool(fields["g"][0]) if "g" in fields else False))
            return self._buf, table, frames
        finally:
            self._pos = saved

    def __iter__(self):
        return self

    def __next__(self):
        frame = self.next_frame()
        if frame is None:
            raise StopIteration
        return frame


def read_clip(path):
    """Decode a whole file → (header, list of frames)."""
    reader = CptvReader(path)
    return reader.get_header(), list(reader)


def decode_clips_device(engine, readers):
    """Decode every frame of several clips on the device (csrc/cptv_kernels.cu).

    ``readers``: ``CptvReader`` objects of clips that share the engine's resolution.  Returns ``(d_frames, clip_first,
    frames)``: a CUDA uint16 tensor ``(total_frames, H, W)`` with the clips back to back, the first frame index of
    every clip (length n + 1) and, per clip, the frames' metadata objects (``pix`` filled by the caller if wanted)."""
    import torch

    from .. import native

    streams, rows, clip_first, frames = [], [], [0], []
    base = 0
    for r in readers:
        h = r.get_header()
        if (h.x_resolution, h.y_resolution) != (engine.width, engine.height):
            raise ValueError("clip resolution {}x{} does not match the engine".format(h.x_resolution, h.y_resolution))
        buf, table, fr = r.index_frames()
        n_px = h.x_resolution * h.y_resolution
        for off, w in table:
            if not 1 <= w <= 24:
                raise ValueError("unsupported CPTV bit width {}".format(w))
            # the payload the device will read must lie inside this clip's stream (a malformed file must not make the
            # unpack kernel read out of bounds)
            if off + 4 + ((n_px - 1) * w + 7) // 8 > len(buf):
                raise ValueError("CPTV frame payload at byte {} runs past the end of the stream".format(off))
            rows.append((base + off, w, 0))
        streams.append(buf)
        base += len(buf)
        clip_first.append(clip_first[-1] + len(table))
        frames.append(fr)
    total = clip_first[-1]
    host = np.frombuffer(b"".join(streams) + b"\0\0\0\0", dtype=np.uint8)  # the last window may read 3 bytes past the end
    d_stream = torch.from_numpy(host.copy()).to(engine.device)
    table = np.array(rows, dtype=native.CPTV_FRAME_DTYPE) if rows else np.zeros(0, native.CPTV_FRAME_DTYPE)
    d_table = torch.from_numpy(table.view(np.uint8).reshape(-1).copy()).to(engine.device)
    d_first = torch.tensor(clip_first, dtype=torch.int32, device=engine.device)
    d_frames = torch.empty((max(total, 1), engine.height, engine.width), dtype=torch.uint16, device=engine.device)
    engine.ctx.use_torch_stream()
    engine.ctx.cptv_decode(d_stream, d_table, total, d_first, len(readers), d_frames)
    return d_frames, clip_first, frames
