"""The reference-facing Python API on the GPU: ClipTrackExtractor.parse_clip / process_frame,
WeightedBackground, Clip / Track / Region, against the reference's golden fixtures."""
import os

import numpy as np
import pytest

from tests import helpers
from tests.tracking_helpers import MemReader, assert_tracks_match_golden

pytestmark = pytest.mark.gpu


def _extractor(**kw):
    from classifier_pipeline_b200.config import Config
    from classifier_pipeline_b200.track.cliptrackextractor import ClipTrackExtractor

    config = Config.get_defaults()
    config.tracking["thermal"].denoise = False
    return ClipTrackExtractor(config.tracking, False, cache_to_disk=False, **kw), config


def _check_frames_and_state(ext, clip, d, tracked):
    bg = helpers.golden_background(d)
    T = len(tracked)
    assert len(clip.frame_buffer.frames) == T
    for t in (0, 1, T // 2, T - 1):
        f = clip.frame_buffer.frames[t]
        assert np.array_equal(f.thermal, tracked[t])
        assert np.array_equal(f.filtered, (tracked[t].astype(np.int64) - bg[t]).astype(np.float32))
        assert np.array_equal(f.mask, d["labels"][t])
    assert np.array_equal(ext.background_alg.background, d["bg_final"].astype(np.float64))
    assert np.array_equal(ext.background_alg.background_weight, d["weight_final"])
    assert ext.background_alg.average == float(d["avg_final"])
    s = clip.stats
    assert np.array_equal(np.array(s.frame_stats_min, np.float64), d["fs_min"])
    assert np.array_equal(np.array(s.frame_stats_max, np.float64), d["fs_max"])
    assert np.array_equal(np.array(s.frame_stats_median, np.float64), d["fs_median"])
    np.testing.assert_allclose(np.array(s.frame_stats_mean, np.float64), d["fs_mean"], rtol=1e-12)
    assert float(s.filtered_sum) == float(d["filtered_sum"])


@pytest.mark.parametrize("name", ["possum", "hedgehog"])
def test_parse_clip_real_clips(name):
    from classifier_pipeline_b200.track.clip import Clip

    d, meta = helpers.load_golden(name + "_raw")
    _, tracked = helpers.clip_input(name + "_raw")
    ext, config = _extractor()
    clip = Clip(config.tracking["thermal"], os.path.join(helpers.GOLDEN, "clips", name + ".cptv"))
    assert ext.parse_clip(clip) is True
    assert ext.tracking_time is not None
    assert_tracks_match_golden(clip, meta, d)
    _check_frames_and_state(ext, clip, d, tracked)


def test_parse_clips_batch_matches_single(monkeypatch):
    """Several clips in one launch == the same clips one by one (clips are independent)."""
    from classifier_pipeline_b200.synthetic import make_clip
    from classifier_pipeline_b200.track.clip import Clip

    names = ["synth0_raw", "synth1_raw", "synth2_raw", "synth3_raw"]
    pix = {}
    for n in names:
        index, frames = helpers.SYNTH[n]
        pix[n] = make_clip(index, frames=frames)
    ext, config = _extractor()
    ext.reader_factory = lambda path: MemReader(pix[path][0], pix[path][1])
    clips = [Clip(config.tracking["thermal"], n) for n in names]
    ext.parse_clips(clips)
    for n, clip in zip(names, clips):
        d, meta = helpers.load_golden(n)
        assert_tracks_match_golden(clip, meta, d)


def test_streaming_process_frame_matches_parse_clip():
    """process_frame one frame at a time (one launch each, state resumed on the device) == parse_clip."""
    from classifier_pipeline_b200.cptv import CptvReader
    from classifier_pipeline_b200.track.clip import Clip
    from classifier_pipeline_b200.track.track import Track

    d, meta = helpers.load_golden("possum_raw")
    _, tracked = helpers.clip_input("possum_raw")
    path = os.path.join(helpers.GOLDEN, "clips", "possum.cptv")
    ext, config = _extractor()
    clip = Clip(config.tracking["thermal"], path)
    ext.init_clip(clip)
    Track._track_id = 1
    reader = CptvReader(path)
    reader.get_header()
    for frame in iter(reader.next_frame, None):
        if frame.background_frame:
            continue
        ext.process_frame(clip, frame)
    ext.apply_track_filtering(clip)
    clip.stats.completed()
    assert_tracks_match_golden(clip, meta, d)
    _check_frames_and_state(ext, clip, d, tracked)


def test_weighted_background_object_matches_oracle():
    """WeightedBackground.process_frame on the device == the reference recurrence (oracle) for both weight_add values."""
    from classifier_pipeline_b200.ml_tools.rectangle import Rectangle
    from classifier_pipeline_b200.piclassifier.motiondetector import WeightedBackground
    from oracle import oracle as orc

    rng = np.random.default_rng(5)
    for weight_add in (0.1, 1):
        wb = WeightedBackground(1, Rectangle(1, 1, 158, 118), 160, 120, weight_add)
        ob = orc.Background(160, 120, 1, weight_add)
        base = rng.integers(2900, 3100, size=(120, 160)).astype(np.float64)
        for t in range(40):
            frame = base + rng.integers(-3, 4, size=(120, 160)) + (t % 7 == 0) * rng.integers(-40, 40, size=(120, 160)) + 0.5
            wb.process_frame(frame)
            ob.process(np.int32(frame))
            bg, w, avg = ob.get()
            assert np.array_equal(wb.background, bg.astype(np.float64)), t
            assert np.array_equal(wb.background_weight, w), t
            assert wb.average == avg, t


def test_parse_clip_default_config_reproduces_reference_possum_json():
    """The reference's own regression: tests/clips/possum.txt is extract.py's output for possum.cptv with the DEFAULT
    config (denoise on).  parse_clip on the device reproduces every track position and both tracking scores."""
    import json

    from classifier_pipeline_b200.config import Config
    from classifier_pipeline_b200.ml_tools.tools import CustomJSONEncoder
    from classifier_pipeline_b200.track.clip import Clip
    from classifier_pipeline_b200.track.cliptrackextractor import ClipTrackExtractor

    config = Config.get_defaults()
    assert config.tracking["thermal"].denoise is True
    ext = ClipTrackExtractor(config.tracking, False, cache_to_disk=False)
    clip = Clip(config.tracking["thermal"], os.path.join(helpers.GOLDEN, "clips", "possum.cptv"))
    ext.parse_clip(clip)
    d, meta = helpers.load_golden("possum_nlm")
    assert_tracks_match_golden(clip, meta, d)
    gold = json.load(open(os.path.join(helpers.GOLDEN, "clips", "possum.txt")))
    got = json.loads(json.dumps(clip.get_metadata(), cls=CustomJSONEncoder))
    assert len(got["tracks"]) == len(gold["tracks"]) == 2
    for a, b in zip(got["tracks"], gold["tracks"]):
        for k in ("id", "start_s", "end_s", "num_frames", "frame_start", "frame_end"):
            assert a[k] == b[k], k
        assert a["tracking_score"] == pytest.approx(b["tracking_score"], rel=1e-6)
        assert len(a["positions"]) == len(b["positions"])
        for p, q in zip(a["positions"], b["positions"]):
            for k in q:
                if k == "pixel_variance":
                    assert p[k] == pytest.approx(q[k], abs=0.011)
                else:
                    assert p[k] == q[k], k
