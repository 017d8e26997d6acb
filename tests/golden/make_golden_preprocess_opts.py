#!/usr/bin/env python
"""Golden fixtures for the NON-DEFAULT normalisation options of the classifier-input preprocessing path, produced by the
UNMODIFIED reference: ``Interpreter.preprocess_segments`` (ml_tools/interpreter.py:315-474) with
``thermal_diff_norm=True`` (thermal normalised by the track-wide limits of ``frame.thermal - median``) and with
``diff_norm=False`` (both channels normalised per tile) on the tracks of ``possum.cptv``.  Build container only:

    python tests/golden/make_golden_preprocess_opts.py      # writes tests/golden/pre_possum_opts.npz
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402
from make_golden import run_reference  # noqa: E402
from make_golden_preprocess import segment_frames_for  # noqa: E402

OPTIONS = {"tdn": dict(thermal_diff_norm=True), "nodiff": dict(diff_norm=False), "tdn_nodiff": dict(thermal_diff_norm=True, diff_norm=False)}


def main():
    ref_harness.setup()
    from ml_tools.hyperparams import HyperParams
    from ml_tools.interpreter import Interpreter
    import ml_tools.preprocess as pp

    class StubInterpreter(Interpreter):
        def __init__(self, opts):
            self.params = HyperParams()
            for k, v in opts.items():
                self.params[k] = v
            self.preprocess_fn = None
            self.seed = None
            self.labels = []

        def shape(self):
            return None

        def predict(self, frames):
            return None

    config, ext, clip, rec = run_reference(os.path.join(HERE, "clips", "possum.cptv"), denoise=False)
    rng = np.random.default_rng(7)
    tracks = list(clip.tracks) + [t for _, t in clip.filtered_tracks if len(t) >= 8]
    arrays = dict(crop=np.array([clip.crop_rectangle.x, clip.crop_rectangle.y, clip.crop_rectangle.width, clip.crop_rectangle.height]))
    meta = dict(name="possum", options=OPTIONS, tracks=[])
    orig_pm = pp.preprocess_movement

    def seeded(*a, **k):
        k["seed"] = 1234
        return orig_pm(*a, **k)

    for ti, track in enumerate(tracks):
        seg_frames = segment_frames_for(track, rng)
        if not seg_frames:
            continue
        regions = np.array([[r.frame_number, r.x, r.y, r.width, r.height, int(r.blank), r.mass] for r in track.bounds_history], np.int32)
        arrays["t{}_regions".format(ti)] = regions
        for si, fr in enumerate(seg_frames):
            arrays["t{}_seg{}".format(ti, si)] = np.asarray(fr, np.int32)
        entry = dict(index=ti, id=track.get_id(), segments=len(seg_frames))
        for tag, opts in OPTIONS.items():
            interp = StubInterpreter(opts)
            np.random.seed(11)
            segments = track.get_segments(segment_width=25, segment_frames=seg_frames)
            pp.preprocess_movement = seeded
            try:
                used, data, masses = interp.preprocess_segments(clip, track, segments)
            finally:
                pp.preprocess_movement = orig_pm
            thermal_limits, filtered_limits = interp.get_limits(clip, track)
            arrays["t{}_out_{}".format(ti, tag)] = np.float32(data)
            entry[tag + "_thermal_limits"] = None if thermal_limits is None else [float(thermal_limits[0]), float(thermal_limits[1])]
            entry[tag + "_filtered_limits"] = None if filtered_limits is None else [float(filtered_limits[0]), float(filtered_limits[1])]
        meta["tracks"].append(entry)
    path = os.path.join(HERE, "pre_possum_opts.npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), **arrays)
    print("tracks", len(meta["tracks"]), "size kB", os.path.getsize(path) // 1000)
    print(json.dumps(meta["tracks"][0])[:400])


if __name__ == "__main__":
    main()
