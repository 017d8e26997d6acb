"""Host-side tracker (H1 region filter, H2 matching / Kalman, H3 track filtering, O1 objects) against the
reference's golden tracks.  No GPU: the per-frame component lists come from the C oracle, exactly the
record layout the kernel emits."""
import numpy as np
import pytest

from tests import helpers
from tests.tracking_helpers import MemReader, assert_tracks_match_golden, oracle_result

RAW = ["possum_raw", "hedgehog_raw", "synth0_raw", "synth1_raw", "synth2_raw", "synth3_raw"]


def _host_extractor(monkeypatch):
    """A ClipTrackExtractor whose device calls are stubbed out (the GPU test exercises the real ones)."""
    from classifier_pipeline_b200.config import Config
    from classifier_pipeline_b200.track import cliptrackextractor as cte

    class NoBackground:
        weight_add = 0.1
        initialised = True

        def process_frame(self, frame):
            pass

    config = Config.get_defaults()
    config.tracking["thermal"].denoise = False
    ext = cte.ClipTrackExtractor(config.tracking, False, cache_to_disk=False, calc_stats=False)
    monkeypatch.setattr(ext, "_new_background", lambda clip: NoBackground())
    return ext, config


@pytest.mark.parametrize("name", RAW)
def test_tracks_match_reference(name, monkeypatch):
    from classifier_pipeline_b200.track.clip import Clip
    from classifier_pipeline_b200.track.track import Track
    from oracle import oracle as orc

    d, meta = helpers.load_golden(name)
    init, tracked = helpers.clip_input(name)
    T = len(tracked)
    p = orc.make_params(background_thresh=meta["background_thresh"], weight_add=meta["weight_add"], max_comp=64)
    o = orc.extract_clip(tracked, init, p)
    ext, config = _host_extractor(monkeypatch)
    all_frames = np.concatenate([init[None], tracked]) if name in helpers.REAL and len(d["labels"]) != len(tracked) + 0 else tracked
    ext.reader_factory = lambda path: MemReader(tracked if name not in helpers.REAL else _real_frames(name), meta["camera_model"])
    clip = Clip(config.tracking["thermal"], name)
    ext.init_clip(clip)
    res = oracle_result(o, T)
    Track._track_id = 1
    reader = ext.reader_factory(name)
    frames = [f for f in iter(reader.next_frame, None)][-T:]
    for t, frame in enumerate(frames):
        ext._consume_frame(clip, frame, res, t)
    ext.apply_track_filtering(clip)
    assert_tracks_match_golden(clip, meta, d)


def _real_frames(name):
    init, tracked = helpers.clip_input(name)
    return tracked


def test_possum_matches_reference_repo_golden_json(monkeypatch):
    """tests/clips/possum.txt of the reference (its own extract.py output, default config): 2 tracks,
    every position identical, tracking_score equal."""
    import json
    import os

    from classifier_pipeline_b200.track.clip import Clip
    from classifier_pipeline_b200.track.track import Track
    from oracle import oracle as orc

    gold = json.load(open(os.path.join(helpers.GOLDEN, "clips", "possum.txt")))
    d, meta = helpers.load_golden("possum_nlm")  # possum.txt was written with the default config: denoise on
    init, tracked = helpers.clip_input("possum_nlm")
    o = orc.extract_clip(tracked, init, orc.make_params(background_thresh=meta["background_thresh"], weight_add=meta["weight_add"],
                                                        max_comp=64, denoise=True))
    ext, config = _host_extractor(monkeypatch)
    ext.reader_factory = lambda path: MemReader(tracked, meta["camera_model"])
    clip = Clip(config.tracking["thermal"], "possum")
    ext.init_clip(clip)
    res = oracle_result(o, len(tracked))
    Track._track_id = 1
    reader = ext.reader_factory("possum")
    for t, frame in enumerate(iter(reader.next_frame, None)):
        ext._consume_frame(clip, frame, res, t)
    ext.apply_track_filtering(clip)
    from classifier_pipeline_b200.ml_tools.tools import CustomJSONEncoder

    got = json.loads(json.dumps(clip.get_metadata(), cls=CustomJSONEncoder))
    assert got["camera_model"] == gold["camera_model"] and got["background_thresh"] == gold["background_thresh"]
    assert len(got["tracks"]) == len(gold["tracks"]) == 2
    for a, b in zip(got["tracks"], gold["tracks"]):
        for k in ("id", "tracker_version", "start_s", "end_s", "num_frames", "frame_start", "frame_end"):
            assert a[k] == b[k], k
        assert a["tracking_score"] == pytest.approx(b["tracking_score"], rel=1e-6)
        assert len(a["positions"]) == len(b["positions"])
        for p, q in zip(a["positions"], b["positions"]):
            assert list(p.keys()) == list(q.keys())
            for k in q:
                if k == "pixel_variance":
                    assert p[k] == pytest.approx(q[k], abs=0.011)
                else:
                    assert p[k] == q[k], k
    # the thumbnail the reference's extract.py chose for each track (classify/thumbnail.py: mass, contour points of the
    # region's mask, warmth against the frame median): same frame, same contour count, same score
    from classifier_pipeline_b200.track.trackextractor import get_metadata

    ext._tracking_time = 0.0
    full = json.loads(json.dumps(get_metadata(None, "possum", None, clip, ext, save=False), cls=CustomJSONEncoder))
    for a, b in zip(full["tracks"], gold["tracks"]):
        ta, tb = a["thumbnail"], b["thumbnail"]
        assert ta["contours"] == tb["contours"] and ta["median_diff"] == tb["median_diff"] and ta["score"] == tb["score"]
        for k in ("x", "y", "width", "height", "mass", "frame_number", "blank", "in_trap"):
            assert ta["region"][k] == tb["region"][k], k
    assert full["algorithm"]["tracker_version"] == gold["algorithm"]["tracker_version"]
