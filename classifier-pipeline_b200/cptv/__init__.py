from .reader import CptvReader, CptvFrame, CptvHeader, read_clip, unpack_deltas

__all__ = ["CptvReader", "CptvFrame", "CptvHeader", "read_clip", "unpack_deltas"]
