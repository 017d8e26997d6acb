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
        ext.process_frame(clip, frame, update_background=True)  # (_track_clip's background update fused into the launch)
    ext.apply_track_filtering(clip)
    clip.stats.completed()
    assert_tracks_match_golden(clip, meta, d)
    _check_frames_and_state(ext, clip, d, tracked)


def test_streaming_process_frame_with_denoise_matches_the_reference():
    """TrackingConfig.denoise=True (the reference's default) one frame at a time: the NLM, mask, component and variance passes
    follow every launch, with the previous frame's outputs kept beside the current ones (CPT_CLIP_PREV_IN_OUTPUT); the tracks
    are those the unmodified reference found with its default configuration."""
    from classifier_pipeline_b200.config import Config
    from classifier_pipeline_b200.cptv import CptvReader
    from classifier_pipeline_b200.track.clip import Clip
    from classifier_pipeline_b200.track.cliptrackextractor import ClipTrackExtractor
    from classifier_pipeline_b200.track.track import Track

    d, meta = helpers.load_golden("possum_nlm")
    path = os.path.join(helpers.GOLDEN, "clips", "possum.cptv")
    config = Config.get_defaults()
    assert config.tracking["thermal"].denoise
    ext = ClipTrackExtractor(config.tracking, False, cache_to_disk=False)
    clip = Clip(config.tracking["thermal"], path)
    ext.init_clip(clip)
    Track._track_id = 1
    reader = CptvReader(path)
    reader.get_header()
    for frame in iter(reader.next_frame, None):
        if frame.background_frame:
            continue
        ext.process_frame(clip, frame, update_background=True)
    ext.apply_track_filtering(clip)
    clip.stats.completed()
    assert_tracks_match_golden(clip, meta, d)


@pytest.mark.parametrize("tag", ["default", "tdn", "nodiff"])
def test_interpreter_preprocess_segments_matches_the_reference(tag):
    """The reference-facing call chain -- parse_clip, Track.get_segments, Interpreter.preprocess_segments / get_limits
    (interpreter.py:315-474) -- with the default HyperParams and with thermal_diff_norm / diff_norm=False, against what the
    unmodified reference produced for the same tracks and frames."""
    import json

    from classifier_pipeline_b200.ml_tools.interpreter import HyperParams, Interpreter
    from classifier_pipeline_b200.track.clip import Clip
    from tests.test_preprocess_oracle import ATOL, OPTION_SETS, RTOL

    path = os.path.join(helpers.GOLDEN, "clips", "possum.cptv")
    ext, config = _extractor()
    clip = Clip(config.tracking["thermal"], path)
    ext.parse_clip(clip)
    tracks = list(clip.tracks) + [t for _, t in clip.filtered_tracks if len(t) >= 8]
    opts = {} if tag == "default" else OPTION_SETS[tag]
    pre = np.load(os.path.join(helpers.GOLDEN, "pre_possum.npz" if tag == "default" else "pre_possum_opts.npz"))
    meta = json.loads(str(pre["meta"]))
    interp = Interpreter(HyperParams(opts), seed=1234)
    assert meta["tracks"]
    for entry in meta["tracks"]:
        ti = entry["index"]
        track = tracks[ti]
        assert track.get_id() == entry["id"]
        seg_frames = [pre["t{}_seg{}".format(ti, s)] for s in range(entry["segments"])]
        segments = track.get_segments(25, segment_frames=seg_frames)
        used, data, masses = interp.preprocess_segments(clip, track, segments)
        expected = pre["t{}_out".format(ti) if tag == "default" else "t{}_out_{}".format(ti, tag)]
        np.testing.assert_allclose(data, expected, rtol=RTOL, atol=ATOL)
        thermal_limits, filtered_limits = interp.get_limits(clip, track)
        if tag == "default":
            assert [float(filtered_limits[0]), float(filtered_limits[1])] == entry["filtered_limits"] and thermal_limits is None
        else:
            want_t, want_f = entry[tag + "_thermal_limits"], entry[tag + "_filtered_limits"]
            assert (thermal_limits is None) == (want_t is None) and (filtered_limits is None) == (want_f is None)
            if want_t is not None:
                assert [float(thermal_limits[0]), float(thermal_limits[1])] == want_t
            if want_f is not None:
                assert [float(filtered_limits[0]), float(filtered_limits[1])] == want_f
        if tag == "default":
            # any channel list over thermal / filtered is a selection of the pair (preprocess.py:169-189)
            three = Interpreter(HyperParams(channels=["thermal", "filtered", "filtered"]), seed=1234)
            _, data3, _ = three.preprocess_segments(clip, track, segments)
            assert data3.shape == expected.shape[:-1] + (3,)
            assert np.array_equal(data3[..., :2], data) and np.array_equal(data3[..., 2], data[..., 1])
            # a preprocess_fn the kernel does not know is applied on the host, the built-in one on the device: same numbers
            from classifier_pipeline_b200.ml_tools.interpreter import inc3_preprocess

            _, on_device, _ = Interpreter(HyperParams(), preprocess_fn=inc3_preprocess, seed=1234).preprocess_segments(clip, track, segments)
            _, on_host, _ = Interpreter(HyperParams(), preprocess_fn=lambda x: x / np.float32(127.5) - np.float32(1.0), seed=1234).preprocess_segments(clip, track, segments)
            np.testing.assert_allclose(on_device, on_host, rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(on_device, expected / np.float32(127.5) - np.float32(1.0), rtol=RTOL, atol=1e-4)


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


def test_process_frame_leaves_the_background_to_the_caller():
    """As in the reference, process_frame never updates the background: a caller that drives background_alg itself the way
    _track_clip does (mean of the last 45 thermal frames after every frame, cliptrackextractor.py:167-176) gets parse_clip's
    tracks, and without those calls the background stays the initial frame."""
    from classifier_pipeline_b200.cptv import CptvReader
    from classifier_pipeline_b200.track.clip import Clip
    from classifier_pipeline_b200.track.track import Track

    d, meta = helpers.load_golden("possum_raw")
    path = os.path.join(helpers.GOLDEN, "clips", "possum.cptv")
    ext, config = _extractor()
    clip = Clip(config.tracking["thermal"], path)
    ext.init_clip(clip)
    first_bg = ext.background_alg.background.copy()
    Track._track_id = 1
    reader = CptvReader(path)
    reader.get_header()
    n = 0
    for frame in iter(reader.next_frame, None):
        if frame.background_frame:
            continue
        ext.process_frame(clip, frame)
        if n == 3:
            assert np.array_equal(ext.background_alg.background, first_bg)  # untouched so far
        if n >= 3:
            last_avg = np.mean([f.thermal for f in clip.frame_buffer.get_last_x(x=45)], axis=0)
            ext.background_alg.process_frame(last_avg)
        n += 1
    assert not np.array_equal(ext.background_alg.background, first_bg)


def test_frames_with_many_components_do_not_abort():
    """A frame with far more components than any sane scene (speckle after a temperature step) is tracked like any other:
    the API engines keep every component the label image can number, and the region list equals the oracle's."""
    from classifier_pipeline_b200.synthetic import make_clip
    from classifier_pipeline_b200.track.clip import Clip
    from oracle import oracle as orc

    pix, model = make_clip(0, frames=30)
    pix = pix.copy()
    speck = np.zeros((120, 160), np.uint16)
    speck[4:116:10, 4:156:10] = 300  # 12 x 16 isolated warm spots
    speck[5:116:10, 4:156:10] = 300
    pix[20] += speck
    ext, config = _extractor()
    ext.reader_factory = lambda path: MemReader(pix, model)
    clip = Clip(config.tracking["thermal"], "speckle")
    ext.parse_clips([clip])
    o = orc.extract_clip(pix, pix[0], orc.make_params(background_thresh=clip.background_thresh, weight_add=0.1, max_comp=255))
    assert 64 < o["ncomp"][20] <= 255
    assert len(clip.frame_buffer.frames) == 30
    mask = clip.frame_buffer.frames[20].mask
    assert int(mask.max()) == min(int(o["ncomp"][20]), 255)
    assert np.array_equal(mask.astype(np.uint8), o["labels"][20])
    assert len(clip.region_history) == 30
