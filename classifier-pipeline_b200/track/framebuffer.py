"""``FrameBuffer``: the clip's frames in host memory (track/framebuffer.py:26-166).

In the batched extractor the per-frame ``filtered`` / ``mask`` arrays are views into the two
buffers copied back from the device once per clip.  The HDF5 spill cache and optical flow are
out of scope.
"""
from threading import Lock

from ..ml_tools.frame import Frame


class FrameBuffer:
    def __init__(self, cptv_name, high_quality_flow=False, cache_to_disk=False, calc_flow=False, keep_frames=True,
                 max_frames=None):
        if cache_to_disk:
            raise NotImplementedError("cache_to_disk (HDF5 frame cache) is not part of the B200 path")
        if calc_flow:
            raise NotImplementedError("optical flow is not part of the B200 path")
        self.cache = None
        self.opt_flow = None
        self.high_quality_flow = high_quality_flow
        self.calc_flow = False
        self.max_frames = max_frames
        self.keep_frames = True if max_frames and max_frames > 0 else keep_frames
        self.prev_frame = None
        self.current_frame = None
        self.current_frame_i = 0
        self.frame_lock = Lock()
        self.reset()

    def reset(self):
        self.frames = []
        self.frames_by_frame_number = {}

    def add_frame(self, thermal, filtered, mask, frame_number, ffc_affected=False):
        self.prev_frame = self.current_frame
        frame = Frame(thermal, filtered, frame_number, mask=mask, ffc_affected=ffc_affected)
        self.current_frame = frame
        if self.keep_frames:
            if self.max_frames and len(self.frames) == self.max_frames:
                with self.frame_lock:
                    del self.frames_by_frame_number[self.frames[0].frame_number]
                    del self.frames[0]
            self.frames.append(frame)
            self.frames_by_frame_number[frame_number] = frame
        return frame

    @property
    def has_flow(self):
        return False

    def get_frame(self, frame_number):
        frame = self.frames_by_frame_number.get(frame_number)
        if frame is not None:
            return frame
        for candidate in (self.prev_frame, self.current_frame):
            if candidate and candidate.frame_number == frame_number:
                return candidate
        return None

    def get_last_x(self, x=25):
        return self.frames[-x:] if len(self.frames) > 0 else None

    def close_cache(self):
        pass

    def remove_cache(self):
        pass

    def __len__(self):
        return len(self.frames)

    def __iter__(self):
        return self

    def __next__(self):
        frame = self.get_frame(self.current_frame_i)
        if frame is None:
            raise StopIteration
        self.current_frame_i += 1
        return frame
