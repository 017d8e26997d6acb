"""``Frame`` / ``TrackChannels``: one frame's channels (ml_tools/frame.py:9-362).

Attributes and methods used on the extraction and classifier-input path keep the reference's
names.  Resizing goes through ``imageprocessing.resize_and_pad`` (device kernels); optical flow
is not part of this path (``flow`` is carried but never computed).
"""
import enum

import numpy as np


class TrackChannels(enum.Enum):
    thermal = 0
    filtered = 1
    flow_h = 2
    flow_v = 3
    mask = 4
    flow = 5


class Frame:
    __slots__ = ("thermal", "filtered", "frame_number", "mask", "flow", "flow_clipped", "scaled_thermal",
                 "ffc_affected", "region", "frame_temp_median", "preprocessed")

    def __init__(self, thermal, filtered, frame_number, mask=None, flow=None, flow_clipped=False, scaled_thermal=None,
                 ffc_affected=False, region=None, frame_temp_median=None, preprocessed=False):
        self.thermal = thermal
        self.filtered = filtered
        self.frame_number = frame_number
        self.mask = mask
        self.flow = flow
        self.flow_clipped = flow_clipped
        self.scaled_thermal = scaled_thermal
        self.ffc_affected = ffc_affected
        self.region = region
        self.frame_temp_median = frame_temp_median
        self.preprocessed = preprocessed

    def get_channel(self, channel):
        if channel == TrackChannels.thermal:
            return self.thermal
        if channel == TrackChannels.filtered:
            return self.filtered
        if channel == TrackChannels.flow:
            return self.flow
        if channel == TrackChannels.mask:
            return self.mask
        return None

    def as_array(self, split_flow=True):
        data = [self.thermal]
        if self.filtered is None:
            return np.array(data)
        data.append(self.filtered)
        if self.mask is not None:
            data.append(self.mask)
        return np.asarray(data)

    def normalize(self):
        from .imageprocessing import normalize

        if self.thermal is not None:
            self.thermal, _ = normalize(self.thermal, new_max=255)
        if self.filtered is not None:
            self.filtered, _ = normalize(self.filtered, new_max=255)

    def crop_by_region(self, region, only_thermal=False, out=None):
        """New frame holding ``region.subimage`` views of every channel (frame.py:203-236)."""
        thermal = region.subimage(self.thermal) if self.thermal is not None else None
        filtered = mask = flow = None
        if not only_thermal:
            filtered = region.subimage(self.filtered) if self.filtered is not None else None
            mask = region.subimage(self.mask) if self.mask is not None else None
            flow = region.subimage(self.flow) if self.flow is not None else None
        if out:
            out.thermal, out.filtered, out.mask, out.flow, out.region = thermal, filtered, mask, flow, region
            return out
        frame = Frame(thermal, filtered, self.frame_number, mask=mask, flow_clipped=self.flow_clipped,
                      ffc_affected=self.ffc_affected, region=region)
        frame.flow = flow
        return frame

    def resize_with_aspect(self, dim, crop_rectangle, keep_edge=False, edge_offset=(0, 0, 0, 0), original_region=None):
        """Aspect-preserving resize of every channel into ``dim`` (frame.py:238-297)."""
        from .imageprocessing import INTER_NEAREST, resize_and_pad

        if self.thermal is not None:
            self.thermal = resize_and_pad(self.thermal, dim, self.region, crop_rectangle, keep_edge=keep_edge,
                                          edge_offset=edge_offset, original_region=original_region)
        if self.mask is not None:
            self.mask = resize_and_pad(self.mask, dim, self.region, crop_rectangle, keep_edge=keep_edge, pad=0,
                                       interpolation=INTER_NEAREST, edge_offset=edge_offset)
        if self.filtered is not None:
            self.filtered = resize_and_pad(self.filtered, dim, self.region, crop_rectangle, keep_edge=keep_edge, pad=0,
                                           edge_offset=edge_offset, original_region=original_region)

    def float_arrays(self):
        for name in ("thermal", "mask", "flow", "filtered"):
            value = getattr(self, name)
            if value is not None:
                setattr(self, name, np.float32(value))

    def copy(self):
        def dup(a):
            return None if a is None else a.copy()

        return Frame(dup(self.thermal), dup(self.filtered), self.frame_number, mask=dup(self.mask), flow=dup(self.flow),
                     flow_clipped=self.flow_clipped, ffc_affected=self.ffc_affected,
                     region=None if self.region is None else self.region.copy())

    def flip(self):
        for name in ("thermal", "mask", "flow", "filtered"):
            value = getattr(self, name)
            if value is not None:
                setattr(self, name, np.flip(value, axis=1))

    @property
    def shape(self):
        return self.thermal.shape
