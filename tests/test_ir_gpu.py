"""The IR side on the device against fixtures the unmodified reference produced (tests/golden/make_golden_ir.py):
detect_objects_ir on 640x480 frames, detect_objects with its default (15,15) kernel and with Otsu, detect_objects_both,
get_diff_back_filtered / DiffBackground, and an IRMotionDetector trace.  Integer outputs bit-exact."""
import os

import numpy as np
import pytest

from tests import helpers, ir_helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(helpers.GOLDEN, "ir.npz"))


@pytest.mark.parametrize("seed", range(4))
def test_detect_objects_ir_640x480(gold, seed):
    from classifier_pipeline_b200.ml_tools.imageprocessing import detect_objects_ir
    from classifier_pipeline_b200.track.irtrackextractor import detect_ir_regions

    img = ir_helpers.ir_filtered(seed)
    n, labels, stats = detect_objects_ir(img, threshold=0)
    assert n == int(gold["ir{}_n".format(seed)])
    assert np.array_equal(labels, gold["ir{}_labels".format(seed)])
    assert np.array_equal(stats, gold["ir{}_stats".format(seed)])
    mask, merged = detect_ir_regions(img)
    assert np.array_equal(np.array(merged).reshape(-1, 5), gold["ir{}_merged".format(seed)])


@pytest.mark.parametrize("seed,shape", [(0, (120, 160)), (1, (120, 160)), (2, (480, 640))])
def test_detect_objects_default_kernel_and_otsu(gold, seed, shape):
    from classifier_pipeline_b200.ml_tools.imageprocessing import detect_objects, detect_objects_both

    img = ir_helpers.thermal_like(10 + seed, shape)
    for otsu in (False, True):
        n, labels, stats, cents = detect_objects(img, otsus=otsu, threshold=70)
        key = "det{}_{}".format(seed, "otsu" if otsu else "thr")
        assert np.array_equal(labels, gold[key + "_labels"]), key
        assert np.array_equal(stats, gold[key + "_stats"]), key
        assert np.array_equal(cents[1:], gold[key + "_cents"][1:]), key
    n, labels, stats = detect_objects_both(ir_helpers.thermal_like(20 + seed, shape), img, threshold=70)
    assert np.array_equal(labels, gold["both{}_labels".format(seed)])
    assert np.array_equal(stats, gold["both{}_stats".format(seed)])


def test_diff_background(gold):
    from classifier_pipeline_b200.track.cliptracker import DiffBackground, get_diff_back_filtered

    rng = np.random.default_rng(5)
    back = rng.integers(40, 90, (480, 640)).astype(np.uint8)
    frame = back.astype(np.int16) + rng.integers(-6, 7, back.shape)
    frame[100:160, 200:300] += 60
    frame = np.clip(frame, 0, 255).astype(np.uint8)
    np.testing.assert_allclose(get_diff_back_filtered(back, frame, 15), gold["diff_filtered"], rtol=1e-6, atol=1e-4)
    db = DiffBackground(15)
    db.set_background(back, frames=1)
    db.update_background(frame)
    np.testing.assert_allclose(db.background, gold["diff_background"], rtol=1e-6, atol=1e-4)


def test_ir_motion_detector_trace(gold):
    """Grey conversion, frame difference, erosion and counts on the device; OpenCV's MOG2 (third party) on the host on both
    sides: the motion flag and the trigger counter of every frame equal the reference's."""
    import cv2
    from classifier_pipeline_b200.piclassifier.irmotiondetector import IRMotionDetector

    cfg, headers = ir_helpers.ir_config()
    det = IRMotionDetector(cfg, headers)
    frames = ir_helpers.ir_video(3)
    flags, trig = [], []
    for f in frames:
        flags.append(bool(det.process_frame(f)))
        trig.append(det.triggered)
    assert np.array_equal(np.array(flags), gold["motion_flags"])
    assert np.array_equal(np.array(trig), gold["motion_triggered"])
    assert gold["motion_flags"].any() and not gold["motion_flags"][:100].any()
    # the device's grey conversion is cv2's
    gray = det._dev.gray(frames[0], 0)
    assert np.array_equal(gray[::16, ::16], gold["gray_check"])
    assert np.array_equal(gray, cv2.cvtColor(frames[0], cv2.COLOR_BGR2GRAY))


def test_ir_erosion_counts_match_cv2():
    """The eroded-pixel counts of cpt_ir_motion_detect == cv2.erode + count for both box sizes, image borders included."""
    import cv2
    from classifier_pipeline_b200 import engine, native

    rng = np.random.default_rng(11)
    dev = native.IrMotion(engine.get_engine().ctx, ir_helpers.W, ir_helpers.H, 2)
    a = rng.integers(0, 100, (ir_helpers.H, ir_helpers.W, 3), dtype=np.uint8)
    b = a.copy()
    b[:40, :50] += 120           # a block touching the corner
    b[200:260, 300:420] += 120
    b[470:, 600:] += 120         # too small to survive a 15 x 15 box away from the border, but the border does not constrain
    ga, gb = dev.gray(a, 0), dev.gray(b, 1)
    mask = (rng.random((ir_helpers.H, ir_helpers.W)) > 0.02).astype(np.uint8) * 255
    for k in (15, 10):
        diff, cnt = dev.detect(1, 0, 12, k, mask=mask)
        delta = cv2.threshold(cv2.absdiff(ga, gb), 12, 255, cv2.THRESH_BINARY)[1]
        assert diff == int((cv2.erode(delta, np.ones((k, k), "uint8")) > 0).sum()), k
        assert cnt == int((cv2.erode(mask, np.ones((k, k), "uint8")) > 0).sum()), k
        assert diff > 0 and cnt > 0
