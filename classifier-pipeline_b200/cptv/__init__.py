from .reader import CptvReader, CptvFrame, CptvHeader, decode_clips_device, read_clip, unpack_deltas

__all__ = ["CptvReader", "CptvFrame", "CptvHeader", "decode_clips_device", "read_clip", "unpack_deltas"]
