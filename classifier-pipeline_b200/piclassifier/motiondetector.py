"""``WeightedBackground``, ``RunningMean``, ``SlidingWindow`` and the ``MotionDetector`` base
(piclassifier/motiondetector.py:7-248).

``WeightedBackground`` keeps its state (background, per-pixel weight counters, average) in device
memory in the layout the extraction kernel uses, so the streaming extractor can filter frames
against it without a copy; ``process_frame`` is one launch of ``background_step_kernel``.
"""
import logging
from abc import ABC, abstractmethod
from threading import Lock

import numpy as np

from .. import engine as _engine


class SlidingWindow:
    """Ring of the most recent frames with an 'oldest non-FFC' cursor (motiondetector.py:7-94)."""

    def __init__(self, shape, dtype=None):
        self.lock = Lock()
        self.frames = [None] * shape
        self.size = len(self.frames)
        self.last_index = None
        self.oldest_index = None
        self.non_ffc_index = None
        self.ffc = False

    def _after_store(self, ffc):
        if not ffc and self.ffc:
            self.non_ffc_index = self.last_index
        self.ffc = ffc

    def update_current_frame(self, frame, ffc=False):
        with self.lock:
            if self.last_index is None:
                self.oldest_index = self.last_index = 0
                if not ffc:
                    self.non_ffc_index = 0
            if not ffc and self.ffc:
                self.non_ffc_index = self.last_index
            self.frames[self.last_index] = frame
            self.ffc = ffc

    def add(self, frame, ffc=False):
        with self.lock:
            if self.last_index is None:
                self.oldest_index = self.last_index = 0
                self.frames[0] = frame
                if not ffc:
                    self.non_ffc_index = 0
            else:
                nxt = (self.last_index + 1) % self.size
                if nxt == self.oldest_index:
                    if self.oldest_index == self.non_ffc_index and not ffc:
                        self.non_ffc_index = (self.oldest_index + 1) % self.size
                    self.oldest_index = (self.oldest_index + 1) % self.size
                self.frames[nxt] = frame
                self.last_index = nxt
            self._after_store(ffc)

    @property
    def current(self):
        with self.lock:
            return None if self.last_index is None else self.frames[self.last_index]

    @property
    def oldest(self):
        with self.lock:
            return None if self.oldest_index is None else self.frames[self.oldest_index]

    @property
    def oldest_nonffc(self):
        with self.lock:
            return None if self.non_ffc_index is None else self.frames[self.non_ffc_index]

    def get(self, i):
        with self.lock:
            return self.frames[i % self.size]

    def get_frames(self):
        with self.lock:
            if self.last_index is None:
                return []
            out = []
            cur, end = self.oldest_index, (self.last_index + 1) % self.size
            while not out or cur != end:
                out.append(self.frames[cur])
                cur = (cur + 1) % self.size
            return out

    def reset(self):
        with self.lock:
            self.last_index = None
            self.oldest_index = None


class RunningMean:
    """uint32 running sum over up to ``window_size`` frames (motiondetector.py:160-175)."""

    def __init__(self, data, window_size):
        self.running_mean = np.sum(data, axis=0, dtype=np.uint32)
        self.running_mean_frames = len(data)
        self.window_size = window_size

    def add(self, new_data, oldest_data):
        if self.running_mean_frames == self.window_size:
            self.running_mean -= oldest_data
            self.running_mean += new_data
        else:
            self.running_mean = self.running_mean + new_data
            self.running_mean_frames += 1

    def mean(self):
        return self.running_mean / self.running_mean_frames


class WeightedBackground:
    """Per-pixel background that follows the frame down immediately and up only after the frame has
    stayed warmer for long enough (motiondetector.py:178-248).  Device resident."""

    def __init__(self, edge_pixels, crop_rectangle, res_x, res_y, weight_add, init_average=None, device=None,
                 max_frames=65534):
        import torch

        self.edge_pixels = edge_pixels
        self.crop_rectangle = crop_rectangle
        self.res_x, self.res_y = res_x, res_y
        self.weight_add = weight_add
        self.engine = _engine.get_engine(device, res_x, res_y, edge_pixels)
        self.ctx = self.engine.ctx
        self.weight_slot = self.ctx.weight_table(weight_add, max_frames=max_frames)
        self.d_state = torch.zeros((1, self.ctx.state_bytes), dtype=torch.uint8, device=self.engine.device)
        self._d_frame = torch.empty((res_y, res_x), dtype=torch.int32, device=self.engine.device)
        self._initialised = False
        self._cache = None
        self._init_average = init_average

    # ------------------------------------------------------------------ device state
    def invalidate(self, average=None):
        """Call after a kernel has written ``d_state`` (the extractor and the motion detector do);
        ``average`` is the new average when the kernel already returned it (saves a state read)."""
        self._cache = None
        self._initialised = True
        self._known_average = average

    def _read(self):
        if self._cache is None:
            self._cache = self.ctx.state_read(self.d_state, 0)
        return self._cache

    @property
    def initialised(self):
        return self._initialised

    @property
    def _background(self):
        return self.background

    @property
    def background(self):
        if not self._initialised:
            return None
        return self._read()["background"].astype(np.float64)

    @property
    def background_weight(self):
        if not self._initialised:
            return np.zeros((self.res_y - 2 * self.edge_pixels, self.res_x - 2 * self.edge_pixels))
        counts = self._read()["weight_count"]
        table = np.array([self.ctx.weight_value(self.weight_slot, k) for k in range(int(counts.max()) + 1)])
        return table[counts]

    @property
    def average(self):
        if not self._initialised:
            if self._init_average is None:
                raise AttributeError("average")
            return self._init_average
        known = getattr(self, "_known_average", None)
        avg = float(known) if known is not None else float(self._read()["average"])
        # np.average(frame) after the first call, int(round(.)) once the background has changed
        # (motiondetector.py:210,232): integral values are handed back as Python ints
        return int(avg) if avg.is_integer() else avg

    def get_average(self):
        return self.average

    # ------------------------------------------------------------------ update
    def process_frame(self, frame):
        import torch

        frame = np.asarray(frame)
        if frame.shape != (self.res_y, self.res_x):
            raise ValueError("frame shape {} does not match {}x{}".format(frame.shape, self.res_x, self.res_y))
        a = np.int32(frame)
        if a.min() < 0 or a.max() > 65535:
            raise ValueError("WeightedBackground frames must lie in the uint16 range")
        self._d_frame.copy_(torch.from_numpy(np.ascontiguousarray(a)), non_blocking=False)
        self.ctx.use_torch_stream()
        self.ctx.background_process(self.d_state, self._d_frame, self.weight_slot)
        self._cache = None
        self._initialised = True
        self._known_average = None

    def set_background_edges(self):
        """Edges are replicated on the device whenever the background changes; nothing to do."""


class MotionDetector(ABC):
    def __init__(self, thermal_config, headers):
        self.movement_detected = False
        self.use_low_power_mode = thermal_config.recorder.use_low_power_mode
        self.num_frames = 0
        self.rec_window = thermal_config.recorder.rec_window
        self.location_config = thermal_config.location
        self.use_sunrise = self.rec_window.use_sunrise_sunset()
        self.last_sunrise_check = None
        self.location = None
        self.sunrise = None
        self.sunset = None
        self.recording = False
        if self.use_sunrise:
            self.rec_window.set_location(*self.location_config.get_lat_long(use_default=True), self.location_config.altitude)
        logging.info("Recording window %s - %s ", self.rec_window.start.dt, self.rec_window.end.dt)
        self.headers = headers

    @property
    def res_x(self):
        return self.headers.res_x

    @property
    def res_y(self):
        return self.headers.res_y

    def can_record(self):
        return self.rec_window.inside_window() and not self.use_low_power_mode

    @abstractmethod
    def process_frame(self, clipped_frame, received_at=None):
        ...

    @abstractmethod
    def preview_frames(self):
        ...

    @abstractmethod
    def get_recent_frame(self):
        ...

    @abstractmethod
    def disconnected(self):
        ...

    @property
    @abstractmethod
    def calibrating(self):
        ...

    @property
    @abstractmethod
    def background(self):
        ...
