"""Host logic of the IR tracker against fixtures the unmodified reference produced (tests/golden/make_golden_ir.py):
rectangle merging and the gap metric (track/irtrackextractor.py:324-389,789-818)."""
import os

import numpy as np

from tests import helpers


def _gold():
    return np.load(os.path.join(helpers.GOLDEN, "ir.npz"))


def test_merge_components_matches_reference():
    from classifier_pipeline_b200.track.irtrackextractor import merge_components

    g = _gold()
    for seed in range(4):
        stats = g["ir{}_stats".format(seed)]
        merged = merge_components(list(stats[1:].copy()))
        assert np.array_equal(np.array(merged).reshape(-1, 5), g["ir{}_merged".format(seed)]), seed
    assert any(len(g["ir{}_merged".format(s)]) < len(g["ir{}_stats".format(s)]) - 1 for s in range(4))  # something did merge


def test_rect_distance_matches_reference():
    from classifier_pipeline_b200.track.irtrackextractor import rect_distance

    a, b = np.array([10, 10, 20, 30, 1]), np.array([50, 70, 5, 5, 1])
    got = np.array([rect_distance(a, b), rect_distance(b, a), rect_distance(a, np.array([15, 60, 5, 5, 1]))])
    assert np.array_equal(got, _gold()["rect_distance"])
