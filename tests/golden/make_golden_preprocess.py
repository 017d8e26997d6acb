#!/usr/bin/env python
"""Golden fixtures for the classifier-input preprocessing path (K9/K10), produced by the UNMODIFIED
reference: ``Interpreter.preprocess_segments`` (ml_tools/interpreter.py:365-474) with default
``HyperParams`` on the tracks the reference extracts (denoise off) from its two test clips and
from seeded synthetic clips.  Build container only (needs /root/reference):

    python tests/golden/make_golden_preprocess.py      # writes tests/golden/pre_*.npz
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


def segment_frames_for(track, rng, n_segments=2):
    """Deterministic frame choices: one evenly spread 25-frame segment (with repeats if the track is short)
    and one random 25-subset; frame numbers are absolute and sorted, blank regions excluded."""
    usable = [r.frame_number for r in track.bounds_history if not r.blank and r.width > 0 and r.height > 0]
    usable = np.array(usable)
    out = []
    if len(usable) == 0:
        return out
    idx = np.round(np.linspace(0, len(usable) - 1, 25)).astype(int)
    out.append(usable[idx])
    if len(usable) >= 25 and n_segments > 1:
        out.append(np.sort(rng.choice(usable, 25, replace=False)))
    elif n_segments > 1:
        out.append(usable[: min(len(usable), 12)])  # short segment: preprocess_movement pads with seeded samples
    return out


def pack(name, source, out_dir):
    ref_harness.setup()
    from ml_tools.hyperparams import HyperParams
    from ml_tools.interpreter import Interpreter

    class StubInterpreter(Interpreter):
        def __init__(self):
            self.params = HyperParams()
            self.preprocess_fn = None
            self.seed = None
            self.labels = []

        def shape(self):
            return None

        def predict(self, frames):
            return None

    config, ext, clip, rec = run_reference(source, denoise=False)
    frames = clip.frame_buffer.frames
    thermal = np.stack([f.thermal for f in frames])
    filtered = np.stack([np.float32(f.filtered) for f in frames])
    interp = StubInterpreter()
    rng = np.random.default_rng(7)
    tracks = list(clip.tracks) + [t for _, t in clip.filtered_tracks if len(t) >= 8]
    arrays = dict(crop=np.array([clip.crop_rectangle.x, clip.crop_rectangle.y, clip.crop_rectangle.width, clip.crop_rectangle.height]))
    meta = dict(name=name, tracks=[])
    for ti, track in enumerate(tracks):
        seg_frames = segment_frames_for(track, rng)
        if not seg_frames:
            continue
        np.random.seed(11)
        segments = track.get_segments(segment_width=25, segment_frames=seg_frames)
        regions = np.array([[r.frame_number, r.x, r.y, r.width, r.height, int(r.blank), r.mass] for r in track.bounds_history], np.int32)
        import ml_tools.preprocess as pp

        # preprocess_movement pads short segments with a seeded generator: pin the seed
        orig_pm = pp.preprocess_movement

        def seeded(*a, **k):
            k["seed"] = 1234
            return orig_pm(*a, **k)

        pp.preprocess_movement = seeded
        try:
            used, data, masses = interp.preprocess_segments(clip, track, segments)
        finally:
            pp.preprocess_movement = orig_pm
        thermal_limits, filtered_limits = interp.get_limits(clip, track)
        arrays["t{}_regions".format(ti)] = regions
        arrays["t{}_out".format(ti)] = np.float32(data)
        for si, fr in enumerate(seg_frames):
            arrays["t{}_seg{}".format(ti, si)] = np.asarray(fr, np.int32)
        meta["tracks"].append(dict(index=ti, id=track.get_id(), segments=len(seg_frames), start_frame=int(track.start_frame),
                                   filtered_limits=[float(filtered_limits[0]), float(filtered_limits[1])]))
    np.savez_compressed(os.path.join(out_dir, "pre_" + name + ".npz"), meta=np.array(json.dumps(meta)),
                        **arrays)
    print(name, "tracks", len(meta["tracks"]), "size kB", os.path.getsize(os.path.join(out_dir, "pre_" + name + ".npz")) // 1000)


def main():
    from classifier_pipeline_b200.synthetic import make_clip

    for clip_name in ("possum", "hedgehog"):
        pack(clip_name, os.path.join(HERE, "clips", clip_name + ".cptv"), HERE)
    pix, model = make_clip(1, frames=120)
    ref_harness.register_memory_clip("synthetic-1-120", pix, model)
    pack("synth1", "synthetic-1-120", HERE)


if __name__ == "__main__":
    main()
