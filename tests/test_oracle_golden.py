"""Pin the C oracle to the reference: every fixture under tests/golden was produced by the
unmodified Python reference (tests/golden/make_golden.py).  Bit-exact for integer/byte work."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests import helpers

ALL = list(helpers.REAL) + list(helpers.SYNTH)


@pytest.mark.parametrize("name", ALL)
def test_extract_matches_reference(name):
    d, meta = helpers.load_golden(name)
    init, tracked = helpers.clip_input(name)
    assert helpers.crc_rows([tracked])[0] == np.uint32(meta["input_crc"])
    p = orc.make_params(
        background_thresh=meta["background_thresh"], weight_add=meta["weight_add"], denoise=meta["denoise"], max_comp=64, calc_stats=True
    )
    o = orc.extract_clip(tracked, init, p)
    T = len(tracked)
    assert T == meta["frames"]
    # K7 background recurrence: every per-frame background, the running average, final weights
    assert np.array_equal(o["bg"], helpers.golden_background(d))
    assert np.array_equal(o["avg"], d["avg"])
    assert np.array_equal(o["final_bg"], d["bg_final"])
    assert np.array_equal(o["final_weight"], d["weight_final"])
    assert o["final_avg"] == float(d["avg_final"])
    # K1 filtered = thermal - background
    assert np.array_equal(o["filtered"], (tracked.astype(np.int64) - helpers.golden_background(d)).astype(np.float32))
    # K2 (+K3) normalised uint8 image and mapped threshold
    assert np.array_equal(o["thresh"].astype(np.float64), d["thresh"])
    assert np.array_equal(o["norm"][:, 0].astype(np.float64), d["norm_max"])
    assert np.array_equal(o["norm"][:, 1].astype(np.float64), d["norm_min"])
    assert np.array_equal(helpers.crc_rows(o["u"]), d["u_crc"])
    assert np.array_equal(o["u"][:: 8], d["u_sub"])
    # K4/K5 label image, stats, centroids
    assert np.array_equal(o["labels"], d["labels"])
    assert np.array_equal(o["ncomp"] + 1, d["ncomp"])
    for t, (gs, gc) in enumerate(helpers.golden_components(d)):
        s, c = orc.stats_centroids_from_comp(o["comp"][t, : o["ncomp"][t]])
        assert np.array_equal(s, gs), t
        assert np.array_equal(c, gc), t
    # K8 frame statistics
    for i, key in enumerate(("fs_min", "fs_max", "fs_median", "fs_mean")):
        assert np.array_equal(o["fstats"][:, i], d[key])
    assert o["fstats"][:, 4].sum() == float(d["filtered_sum"])
    # K6 per-region variance (tolerance class: the reference reduces in fp32)
    for row in d["regions"]:
        t, rid, var = int(row[0]), int(row[7]), row[6]
        assert o["var"][t, rid] == pytest.approx(var, rel=2e-4, abs=1e-4)


def test_nlm_small_images_edges():
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (9, 11), dtype=np.uint8)
    out = orc.nlm_denoise(img)
    assert out.shape == img.shape
    cv2 = pytest.importorskip("cv2")
    assert np.array_equal(out, cv2.fastNlMeansDenoising(img, None))


def test_cc_label_order_random_masks():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    for i in range(60):
        h, w = int(rng.integers(1, 40)), int(rng.integers(1, 50))
        mask = (rng.random((h, w)) < rng.uniform(0.05, 0.7)).astype(np.uint8) * 255
        n, labels, comp = orc.cc8(mask)
        n2, l2, s2, c2 = cv2.connectedComponentsWithStats(mask)
        assert n + 1 == n2
        assert np.array_equal(labels, l2)
        s, c = orc.stats_centroids_from_comp(comp)
        assert np.array_equal(s, s2[1:])
        assert np.array_equal(c, c2[1:])


def test_blur_threshold_close_vs_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for shape in ((120, 160), (7, 5), (1, 9), (3, 1), (2, 2)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        assert np.array_equal(orc.blur5(img), cv2.GaussianBlur(img, (5, 5), 0))
        for th in (12.75, 0.3, 254.2, 255.0, 300.0):
            _, m = cv2.threshold(cv2.GaussianBlur(img, (5, 5), 0), th, 255, cv2.THRESH_BINARY)
            m = cv2.morphologyEx(m, cv2.MORPH_CLOSE, (5, 5))
            ours = orc.threshold_close(orc.blur5(img), th)
            assert np.array_equal(ours * 255, m), (shape, th)
