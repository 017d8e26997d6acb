"""The integer keep-test table (cpt_build_weight_table) reproduces the reference's literal fp64
comparison `background < frame - background_weight` (piclassifier/motiondetector.py:214-218)
for every count k and every integer (frame, background) pair that can matter."""
import ctypes

import numpy as np
import pytest


def build(weight_add, n):
    import __graft_entry__ as g

    g.build()
    from classifier_pipeline_b200 import native

    lib = native.load()
    thr = np.zeros(n, np.uint32)
    w = np.zeros(n, np.float64)
    native.check(lib.cpt_build_weight_table(float(weight_add), n, thr.ctypes.data, w.ctypes.data))
    return thr, w


def literal_weights(weight_add, n):
    w = np.zeros(n, np.float64)
    acc = np.float64(0.0)
    for k in range(n):
        w[k] = acc
        acc = acc + np.float64(weight_add)
    return w


@pytest.mark.parametrize("weight_add", [0.1, 1.0, 0.05, 0.3, 1.0 / 3.0, 2.5, 0.0, 1000.0])
def test_table_equals_fp64_comparison(weight_add):
    n = 3000
    thr, w = build(weight_add, n)
    assert np.array_equal(w, literal_weights(weight_add, n))
    t = (thr & 0xFFFF).astype(np.int64)
    bound = (thr >> 16).astype(np.int64)
    rng = np.random.default_rng(0)
    # all k, backgrounds across every binade (incl. 0 and powers of two +-1), d around the threshold
    bs = np.unique(np.concatenate([[0, 1, 2, 3], 2 ** np.arange(1, 16), 2 ** np.arange(1, 16) - 1, 2 ** np.arange(1, 16) + 1,
                                   rng.integers(0, 65536, 40)]))
    for b in bs:
        for dd in (-2, -1, 0, 1, 2):
            d = np.ceil(w).astype(np.int64) + dd
            a = b + d
            ok = (a >= 0) & (a <= 65535)
            literal = np.float64(b) < (a.astype(np.float64) - w)
            table = (d >= t) | ((d == t - 1) & (b < bound))
            ok &= t < 65535  # clamped entries (w_k beyond any pixel difference) are excluded by the host check
            assert np.array_equal(literal[ok], table[ok]), (weight_add, int(b), dd)
            assert np.array_equal(table, d >= t - (b < bound))


def test_known_rounding_trap():
    """k=10 with weight_add 0.1: w = 0.9999999999999999; frame - background == 1 is NOT kept
    although 1 > w, because fl(frame - w) rounds to background."""
    thr, w = build(0.1, 32)
    assert w[10] == 0.9999999999999999
    t, bound = int(thr[10] & 0xFFFF), int(thr[10] >> 16)
    assert t == 2 and bound == 1  # d == 1 keeps only when background == 0
