"""``CPTVMotionDetector`` and ``is_affected_by_ffc`` (piclassifier/cptvmotiondetector.py:14-234)."""
import ctypes
from datetime import timedelta

import numpy as np

from .. import native
from ..ml_tools.rectangle import Rectangle
from .motiondetector import MotionDetector, SlidingWindow, WeightedBackground

FFC_PERIOD = timedelta(seconds=9.9)


def is_affected_by_ffc(cptv_frame):
    """A frame taken during / just after a flat-field correction (cptvmotiondetector.py:211-222).

    parity: with integer ``time_on`` / ``last_ffc_time`` (milliseconds from the CPTV decoder) the
    difference is compared with ``FFC_PERIOD.seconds`` == 9, i.e. nine *milliseconds*."""
    if hasattr(cptv_frame, "ffc_status") and cptv_frame.ffc_status in [1, 2]:
        return True
    if cptv_frame.time_on is None or cptv_frame.last_ffc_time is None:
        return False
    if isinstance(cptv_frame.time_on, int):
        return (cptv_frame.time_on - cptv_frame.last_ffc_time) < FFC_PERIOD.seconds
    return (cptv_frame.time_on - cptv_frame.last_ffc_time) < FFC_PERIOD


def since_ffc(cptv_frame):
    if hasattr(cptv_frame, "ffc_status") and cptv_frame.ffc_status in [1, 2]:
        return True
    if cptv_frame.time_on is None or cptv_frame.last_ffc_time is None:
        return False
    return cptv_frame.time_on - cptv_frame.last_ffc_time


class CPTVMotionDetector(MotionDetector):
    """Streaming motion trigger for Lepton frames (cptvmotiondetector.py:14-205).

    Same constructor, attributes and return values as the reference class.  The frame ring, the uint32
    running sum, the weighted background and the motion count live on the device; ``process_frame`` is one
    host->device frame copy, one fused kernel launch (csrc/motion_kernels.cu) and a 32-byte read-back.  The
    ring bookkeeping (``SlidingWindow`` cursors, FFC handling, trigger counting) stays on the host and keeps
    the frame objects for ``preview_frames`` / ``get_recent_frame``."""

    FFC_PERIOD = FFC_PERIOD
    BACKGROUND_WEIGHT_ADD = 0.1
    MEAN_FRAMES = 45

    def __init__(self, thermal_config, dynamic_thresh, headers, detect_after=None, device=None):
        super().__init__(thermal_config, headers)
        self.headers = headers
        if headers.model and headers.model.lower() == "lepton3.5":
            CPTVMotionDetector.BACKGROUND_WEIGHT_ADD = 1  # class attribute, as in the reference
        self.config = thermal_config.motion
        self.location_config = thermal_config.location
        self.num_preview_frames = thermal_config.recorder.preview_secs * headers.fps
        self.compare_gap = self.config.frame_compare_gap + 1
        edge = self.config.edge_pixels
        self.min_frames = thermal_config.recorder.min_secs * headers.fps
        self.max_frames = thermal_config.recorder.max_secs * headers.fps
        if not self.config.one_diff_only:
            self.diff_window = SlidingWindow(self.compare_gap, np.int32)
        self.running_mean = None  # becomes the device running sum (True once started)
        self.thermal_window = SlidingWindow(self.num_preview_frames + 1, "O")
        self.processed = 0
        self.thermal_thresh = 0
        self.crop_rectangle = Rectangle(edge, edge, headers.res_x - 2 * edge, headers.res_y - 2 * edge)
        self._background = WeightedBackground(edge, self.crop_rectangle, self.res_x, self.res_y,
                                              CPTVMotionDetector.BACKGROUND_WEIGHT_ADD, self.config.temp_thresh, device=device)
        self.movement_detected = False
        self.dynamic_thresh = dynamic_thresh
        self.triggered = 0
        self.ffc_affected = False
        self.detect_after = self.thermal_window.size * 2 if detect_after is None else detect_after
        self._ctx = self._background.ctx
        lib = self._ctx.lib
        self._lib = lib
        self._m = lib.cpt_motion_open(self._ctx._h, self.thermal_window.size, self.MEAN_FRAMES,
                                      0 if self.config.one_diff_only else self.compare_gap, self._background.weight_slot)
        if not self._m:
            raise native.NativeError("cpt_motion_open failed: " + lib.cpt_last_error().decode())
        self._result = native.CptMotionResult()
        self._average = None
        self.last_diff = 0

    def __del__(self):
        try:
            if getattr(self, "_m", None):
                self._lib.cpt_motion_close(self._m)
                self._m = None
        except Exception:
            pass

    @property
    def calibrating(self):
        return self.ffc_affected

    def preview_frames(self):
        return self.thermal_window.get_frames()[:-1]

    @property
    def temp_thresh(self):
        return self._background.average

    @property
    def background(self):
        return self._background.background

    def get_recent_frame(self):
        return self.thermal_window.current

    def disconnected(self):
        self.thermal_window.reset()
        if not self.config.one_diff_only:
            self.diff_window.reset()
        self.processed = 0

    @staticmethod
    def _pix(cptv_frame):
        pix = np.ascontiguousarray(cptv_frame.pix)
        if pix.dtype != np.uint16:
            raise TypeError("CPTVMotionDetector frames must be uint16")
        return pix

    def process_frame(self, cptv_frame, force_process=False):
        prev_ffc = self.ffc_affected
        self.ffc_affected = is_affected_by_ffc(cptv_frame)
        if self.can_record() or force_process:
            pix = self._pix(cptv_frame)
            if pix.shape != (self.res_y, self.res_x):
                raise ValueError("frame shape {} does not match {}x{}".format(pix.shape, self.res_x, self.res_y))
            win = self.thermal_window
            win.add(cptv_frame, self.ffc_affected)
            flags = 0
            if self.running_mean is None:
                stored = win.get_frames()[: self.MEAN_FRAMES]
                if len(stored) == 1:
                    flags |= native.MOTION_MEAN | native.MOTION_MEAN_RESTART
                else:
                    # frames were only stored so far (outside the recording window): RunningMean(last_45)
                    native.check(self._lib.cpt_motion_store(self._m, pix.ctypes.data, win.last_index))
                    slots = [(win.oldest_index + i) % win.size for i in range(len(stored))]
                    arr = (ctypes.c_int32 * len(slots))(*slots)
                    native.check(self._lib.cpt_motion_mean_init(self._m, arr, len(slots)))
                self.running_mean = True
            else:
                flags |= native.MOTION_MEAN
            if not self.ffc_affected:
                flags |= native.MOTION_BACKGROUND
            reset = self.ffc_affected or prev_ffc
            detect = not reset and self.processed > self.detect_after
            diff_new = diff_old = -1
            if reset and prev_ffc:
                win.non_ffc_index = win.last_index
            if detect:
                flags |= native.MOTION_DETECT
                if self.config.warmer_only:
                    flags |= native.MOTION_WARMER_ONLY
                if self.config.one_diff_only:
                    flags |= native.MOTION_ONE_DIFF
                else:
                    dw = self.diff_window
                    if self.processed > 2 and dw.non_ffc_index is not None:
                        diff_old = dw.non_ffc_index
                    dw.add(True, self.ffc_affected)
                    diff_new = dw.last_index
            init_avg = self._background._init_average
            native.check(self._lib.cpt_motion_step(
                self._m, pix.ctypes.data, self._background.d_state.data_ptr(), win.last_index,
                -1 if win.oldest_index is None else win.oldest_index,
                -1 if win.non_ffc_index is None else win.non_ffc_index, diff_new, diff_old, flags,
                int(self.config.delta_thresh), float(init_avg if init_avg is not None else 0.0), ctypes.byref(self._result)))
            if flags & native.MOTION_BACKGROUND:
                self._background.invalidate(average=self._result.average)
            if reset:
                self.movement_detected = False
                self.triggered = 0
            elif detect:
                self.last_diff = int(self._result.diff)
                movement = self.last_diff > self.config.count_thresh
                self.triggered = self.triggered + 1 if movement else 0
                self.movement_detected = self.triggered >= self.config.trigger_frames
            self.processed += 1
        else:
            self.thermal_window.update_current_frame(cptv_frame, self.ffc_affected)
            native.check(self._lib.cpt_motion_store(self._m, self._pix(cptv_frame).ctypes.data, self.thermal_window.last_index))
            self.movement_detected = False
        self.num_frames += 1
        return self.movement_detected

    def skip_frame(self):
        return
