"""Golden fixtures for the IR side, produced by the UNMODIFIED reference in the build
container: detect_objects_ir / detect_objects (default (15,15) kernel, Otsu) / detect_objects_both
(ml_tools/imageprocessing.py:185-248), IRTrackExtractor.merge_components + rect_distance (track/irtrackextractor.py:324-389,
789-818), get_diff_back_filtered / DiffBackground (track/cliptracker.py:612-668) and an IRMotionDetector trace
(piclassifier/irmotiondetector.py:103-153).  Run:  python tests/golden/make_golden_ir.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import ir_helpers  # noqa: E402
from tests.golden import ref_harness  # noqa: E402


def main():
    ref_harness.setup()
    import types

    for name in ("portalocker", "astral"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["astral"].Location = object
    import cv2
    from ml_tools import imageprocessing as ip
    from piclassifier.irmotiondetector import IRMotionDetector
    from track import cliptracker as ct
    from track.irtrackextractor import IRTrackExtractor, rect_distance

    out = {}
    # ---- detect_objects_ir on background-subtracted 640x480 frames, then the rectangle merging
    merger = IRTrackExtractor.__new__(IRTrackExtractor)
    merger.scale = None
    for seed in range(4):
        img = ir_helpers.ir_filtered(seed)
        n, labels, stats = ip.detect_objects_ir(img, threshold=0)
        out["ir{}_n".format(seed)] = n
        out["ir{}_labels".format(seed)] = labels.astype(np.int32)
        out["ir{}_stats".format(seed)] = stats
        merged = merger.merge_components(list(stats[1:].copy()))
        out["ir{}_merged".format(seed)] = np.array(merged).reshape(-1, 5)
    a, b = np.array([10, 10, 20, 30, 1]), np.array([50, 70, 5, 5, 1])
    out["rect_distance"] = np.array([rect_distance(a, b), rect_distance(b, a), rect_distance(a, np.array([15, 60, 5, 5, 1]))])
    # ---- detect_objects with its default kernel, with Otsu; detect_objects_both
    for seed, shape in enumerate([(120, 160), (120, 160), (480, 640)]):
        img = ir_helpers.thermal_like(10 + seed, shape)
        for otsu in (False, True):
            n, labels, stats, cents = ip.detect_objects(img, otsus=otsu, threshold=70)
            key = "det{}_{}".format(seed, "otsu" if otsu else "thr")
            out[key + "_labels"], out[key + "_stats"], out[key + "_cents"] = labels.astype(np.int32), stats, cents
        n, labels, stats = ip.detect_objects_both(ir_helpers.thermal_like(20 + seed, shape), img, threshold=70)
        out["both{}_labels".format(seed)], out["both{}_stats".format(seed)] = labels.astype(np.int32), stats
    # ---- get_diff_back_filtered / DiffBackground
    rng = np.random.default_rng(5)
    back = rng.integers(40, 90, (480, 640)).astype(np.uint8)
    frame = back.astype(np.int16) + rng.integers(-6, 7, back.shape)
    frame[100:160, 200:300] += 60
    frame = np.clip(frame, 0, 255).astype(np.uint8)
    out["diff_filtered"] = ct.get_diff_back_filtered(back, frame, 15)
    db = ct.DiffBackground(15)
    db.set_background(back, frames=1)
    db.update_background(frame)
    out["diff_background"] = db.background
    # ---- IRMotionDetector trace
    cfg, headers = ir_helpers.ir_config()
    det = IRMotionDetector(cfg, headers)
    flags, trig = [], []
    for f in ir_helpers.ir_video(3):
        flags.append(bool(det.process_frame(f)))
        trig.append(det.triggered)
    out["motion_flags"] = np.array(flags)
    out["motion_triggered"] = np.array(trig)
    out["gray_check"] = cv2.cvtColor(ir_helpers.ir_video(3, frames=1)[0], cv2.COLOR_BGR2GRAY)[::16, ::16]
    np.savez_compressed(os.path.join(HERE, "ir.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
    print("motion frames", np.nonzero(out["motion_flags"])[0])


if __name__ == "__main__":
    main()
