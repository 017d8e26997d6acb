#!/usr/bin/env python
"""Golden fixtures for the streaming motion detector (M1), produced by the UNMODIFIED reference
``piclassifier.cptvmotiondetector.CPTVMotionDetector`` (cptvmotiondetector.py:14-205) fed frame by frame.
Build container only (needs /root/reference):

    python tests/golden/make_golden_motion.py      # writes tests/golden/motion_*.npz

Each fixture stores the configuration, the per-frame inputs that are not pixels (ffc flag, inside-window flag)
and, per frame, what the reference returned / held afterwards: movement_detected, triggered, temp_thresh,
processed; plus the final background and running sum.  Pixels come from the committed clips / the seeded
synthetic generator, so they are not stored.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402

from tests.motion_helpers import CASES, Headers, StreamFrame, frames_for, thermal_config  # noqa: E402


def run_case(name):
    ref_harness.setup()
    from piclassifier.cptvmotiondetector import CPTVMotionDetector

    source, model, overrides, preview_secs, detect_after, ffc_frames, outside = CASES[name]
    CPTVMotionDetector.BACKGROUND_WEIGHT_ADD = 0.1  # class attribute mutated by lepton3.5 detectors
    cfg = thermal_config(model, preview_secs, **overrides)
    det = CPTVMotionDetector(cfg, None, Headers(model), detect_after=detect_after)
    rows = []
    for t, pix in enumerate(frames_for(source)):
        cfg.recorder.rec_window.inside = t not in outside
        moved = det.process_frame(StreamFrame(pix, t, ffc=t in ffc_frames))
        rows.append([int(bool(moved)), int(det.triggered), float(det.temp_thresh), int(det.processed)])
    meta = dict(name=name, source=source, model=model, overrides=overrides, preview_secs=preview_secs, detect_after=detect_after,
                ffc_frames=ffc_frames, outside_window=outside, weight_add=float(CPTVMotionDetector.BACKGROUND_WEIGHT_ADD))
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), meta=np.array(json.dumps(meta)), rows=np.array(rows, dtype=np.float64),
        background=np.asarray(det.background, dtype=np.float64), running_sum=np.asarray(det.running_mean.running_mean, dtype=np.uint32),
        running_frames=np.int32(det.running_mean.running_mean_frames),
        background_weight=np.asarray(det._background.background_weight, dtype=np.float64))
    r = np.array(rows)
    print(name, "frames", len(rows), "motion frames", int(r[:, 0].sum()), "first", int(np.argmax(r[:, 0])) if r[:, 0].any() else None,
          "temp_thresh", r[-1, 2])


if __name__ == "__main__":
    for case in CASES:
        run_case(case)
