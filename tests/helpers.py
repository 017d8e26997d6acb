"""Shared helpers for the parity tests (fixtures -> arrays)."""
import json
import os
import zlib

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

REAL = ("possum_raw", "hedgehog_raw", "possum_nlm", "hedgehog_nlm")
SYNTH = {"synth0_raw": (0, 120), "synth1_raw": (1, 120), "synth2_raw": (2, 100), "synth3_raw": (3, 100), "synth4_nlm": (4, 48)}


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(d["meta"]))
    return d, meta


def golden_background(d):
    """Per-frame background used for frame t, rebuilt from first + deltas."""
    first = d["bg_first"].astype(np.int32)
    return np.concatenate([first[None], first[None] + np.cumsum(d["bg_delta"], axis=0)])


def clip_input(name):
    """(init_frame, tracked frames uint16 (T,H,W)) for a golden fixture."""
    from classifier_pipeline_b200.cptv import CptvReader
    from classifier_pipeline_b200.synthetic import make_clip

    if name in SYNTH:
        index, frames = SYNTH[name]
        pix, _ = make_clip(index, frames=frames)
        return pix[0], pix
    reader = CptvReader(os.path.join(GOLDEN, "clips", name.split("_")[0] + ".cptv"))
    reader.get_header()
    frames = []
    while True:
        f = reader.next_frame()
        if f is None:
            break
        frames.append(f)
    init = frames[0].pix
    tracked = np.stack([f.pix for f in frames if not f.background_frame])
    return init, tracked


def crc_rows(a):
    return np.array([zlib.crc32(np.ascontiguousarray(x).tobytes()) for x in a], dtype=np.uint32)


def golden_components(d):
    """List per frame of (stats (n,5) int32, centroids (n,2) f64) without the background row."""
    out, idx = [], 0
    for n in d["ncomp"]:
        out.append((d["stats"][idx + 1 : idx + n], d["centroids"][idx + 1 : idx + n]))
        idx += n
    return out
