"""CPU restatement (numpy) of the streaming motion detector, M1.  TEST INFRASTRUCTURE ONLY.

Restates ``CPTVMotionDetector.process_frame / detect`` (piclassifier/cptvmotiondetector.py:74-205) together
with ``SlidingWindow`` (piclassifier/motiondetector.py:7-94), ``RunningMean`` (:160-175) and
``WeightedBackground`` (:178-248) as one flat state machine over arrays -- no classes shared with the
product code.  Pinned by tests/test_motion_oracle.py against fixtures generated from the reference itself
(tests/golden/motion_*.npz, tests/golden/make_golden_motion.py).
"""
import numpy as np


class Ring:
    """SlidingWindow cursors (motiondetector.py:7-94): newest, oldest, oldest non-FFC."""

    def __init__(self, size):
        self.size = size
        self.items = [None] * size
        self.last = self.oldest = self.nonffc = None
        self.ffc = False

    def add(self, item, ffc):
        if self.last is None:
            self.last = self.oldest = 0
            self.items[0] = item
            if not ffc:
                self.nonffc = 0
        else:
            nxt = (self.last + 1) % self.size
            if nxt == self.oldest:
                if self.oldest == self.nonffc and not ffc:
                    self.nonffc = (self.oldest + 1) % self.size
                self.oldest = (self.oldest + 1) % self.size
            self.items[nxt] = item
            self.last = nxt
        if not ffc and self.ffc:
            self.nonffc = self.last
        self.ffc = ffc

    def replace_current(self, item, ffc):
        if self.last is None:
            self.last = self.oldest = 0
            if not ffc:
                self.nonffc = 0
        if not ffc and self.ffc:
            self.nonffc = self.last
        self.items[self.last] = item
        self.ffc = ffc

    def ordered(self):
        if self.last is None:
            return []
        out, cur, end = [], self.oldest, (self.last + 1) % self.size
        while not out or cur != end:
            out.append(self.items[cur])
            cur = (cur + 1) % self.size
        return out

    def reset(self):
        self.last = self.oldest = None


def weighted_background_step(state, frame, edge, weight_add):
    """WeightedBackground.process_frame (motiondetector.py:197-244).  state: dict(background, weight, average)."""
    a = np.int32(frame[edge:-edge, edge:-edge] if edge else frame)
    if state["background"] is None:
        bg = np.empty(frame.shape)
        bg[edge : frame.shape[0] - edge, edge : frame.shape[1] - edge] = a
        state["background"], state["average"] = bg, np.average(a)
    else:
        bg = state["background"]
        inner = bg[edge : bg.shape[0] - edge, edge : bg.shape[1] - edge]
        keep = inner < a - state["weight"]
        new = np.where(keep, inner, a)
        state["weight"] = np.where(keep, state["weight"] + weight_add, 0)
        if not np.any(new != inner):
            return
        inner[:, :] = new
        state["average"] = int(round(np.average(inner)))
    for i in range(edge):
        bg[i] = bg[edge]
        bg[-i - 1] = bg[-edge - 1]
        bg[:, i] = bg[:, edge]
        bg[:, -i - 1] = bg[:, -1 - edge]


def run_motion(frames, weight_add, temp_thresh, delta_thresh, count_thresh, trigger_frames, edge, warmer_only,
               one_diff_only, frame_compare_gap, preview_frames, detect_after=None, ffc_frames=(), outside_window=(),
               mean_frames=45):
    """Feed ``frames`` through the detector; returns (rows [moved, triggered, temp_thresh, processed], final state)."""
    H, W = frames[0].shape
    window = Ring(preview_frames + 1)
    deltas = None if one_diff_only else Ring(frame_compare_gap + 1)
    bg = dict(background=None, weight=np.zeros((H - 2 * edge, W - 2 * edge)), average=temp_thresh)
    run_sum, run_n = None, 0
    if detect_after is None:
        detect_after = window.size * 2
    processed = triggered = 0
    moved = prev_ffc = False
    rows = []
    for t, pix in enumerate(frames):
        ffc = t in ffc_frames
        if t not in outside_window:
            window.add(pix, ffc)
            oldest = window.items[window.oldest]
            if run_sum is None:
                first = window.ordered()[:mean_frames]
                run_sum, run_n = np.sum(first, axis=0, dtype=np.uint32), len(first)
            elif run_n == mean_frames:
                run_sum -= oldest  # uint32 modular, in place (motiondetector.py:167-169)
                run_sum += pix
            else:
                run_sum = run_sum + pix
                run_n += 1
            if not ffc:
                weighted_background_step(bg, run_sum / run_n, edge, weight_add)
            if ffc or prev_ffc:
                moved, triggered = False, 0
                if prev_ffc:
                    window.nonffc = window.last
            elif processed > detect_after:
                T = bg["average"]
                crop = np.s_[edge : H - edge, edge : W - edge]
                ref = np.clip(window.items[window.nonffc][crop], a_min=T, a_max=None)
                cur = np.clip(np.int32(pix[crop]), a_min=T, a_max=None)
                delta = cur - ref
                if not warmer_only:
                    delta = abs(delta)
                if one_diff_only:
                    count = int(np.count_nonzero(delta > delta_thresh))
                else:
                    delta[delta >= delta_thresh] = delta_thresh
                    count = 0
                    if processed > 2:
                        count = int(np.count_nonzero(deltas.items[deltas.nonffc] + delta == delta_thresh * 2))
                    deltas.add(delta, ffc)
                triggered = triggered + 1 if count > count_thresh else 0
                moved = triggered >= trigger_frames
            processed += 1
        else:
            window.replace_current(pix, ffc)
            moved = False
        prev_ffc = ffc
        rows.append([int(moved), triggered, float(bg["average"]), processed])
    return np.array(rows, dtype=np.float64), dict(background=bg["background"], weight=bg["weight"], running_sum=run_sum, running_frames=run_n)
