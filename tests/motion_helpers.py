"""Shared by the motion-detector fixtures generator and tests: duck-typed thermal config, headers, frames
and the fixture cases (tests/golden/motion_*.npz)."""
import json
import os
import types

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

MOTION_DEFAULTS = {
    # config/thermalconfig.py:82-104
    "lepton3": dict(temp_thresh=2750, delta_thresh=50, count_thresh=3, frame_compare_gap=45, one_diff_only=True,
                    trigger_frames=2, edge_pixels=1, warmer_only=True),
    "lepton3.5": dict(temp_thresh=28000, delta_thresh=150, count_thresh=3, frame_compare_gap=45, one_diff_only=True,
                      trigger_frames=2, edge_pixels=1, warmer_only=True),
}


class Window:
    def __init__(self):
        self.inside = True
        self.start = types.SimpleNamespace(dt="00:00")
        self.end = types.SimpleNamespace(dt="00:00")

    def use_sunrise_sunset(self):
        return False

    def inside_window(self):
        return self.inside


def thermal_config(model, preview_secs=5, **overrides):
    motion = dict(MOTION_DEFAULTS[model])
    motion.update(overrides)
    return types.SimpleNamespace(
        motion=types.SimpleNamespace(**motion),
        recorder=types.SimpleNamespace(use_low_power_mode=False, rec_window=Window(), preview_secs=preview_secs, min_secs=5, max_secs=600),
        location=types.SimpleNamespace(),
    )


class Headers:
    def __init__(self, model, res_x=160, res_y=120, fps=9):
        self.model, self.res_x, self.res_y, self.fps = model, res_x, res_y, fps


class StreamFrame:
    def __init__(self, pix, t, ffc=False):
        self.pix = pix
        self.time_on = 10_000_000 + t * 111
        self.last_ffc_time = 0
        if ffc:
            self.ffc_status = 1


CASES = {
    # name: (source, model, config overrides, preview_secs, detect_after, ffc frames, outside-window frames)
    "motion_possum": ("possum", "lepton3", {}, 5, 0, [], []),
    "motion_synth1": ("synth:1:200", "lepton3.5", {}, 5, None, [60, 61, 62], []),
    "motion_synth0_twodiff": ("synth:0:160", "lepton3", dict(one_diff_only=False, warmer_only=False, frame_compare_gap=8, delta_thresh=30), 3, 0,
                              [40, 41], list(range(0, 4))),
    "motion_synth2_short": ("synth:2:120", "lepton3", dict(delta_thresh=20, count_thresh=1), 2, 5, [], list(range(30, 36))),
}


def frames_for(source):
    from classifier_pipeline_b200.synthetic import make_clip

    if source.startswith("synth:"):
        _, index, n = source.split(":")
        pix, _ = make_clip(int(index), frames=int(n))
        return list(pix)
    from classifier_pipeline_b200.cptv import CptvReader

    reader = CptvReader(os.path.join(HERE, "clips", source + ".cptv"))
    reader.get_header()
    out = []
    while True:
        f = reader.next_frame()
        if f is None:
            return out
        out.append(f.pix)




def load_fixture(name):
    d = np.load(os.path.join(HERE, name + ".npz"))
    return d, json.loads(str(d["meta"]))
