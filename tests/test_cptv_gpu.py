"""CPTV v2 decode on the device (cpt_cptv_decode) == the host decoder, for the reference's clips and for
synthetic streams of every bit width; parse_clip on files goes through it."""
import os

import numpy as np
import pytest

from tests import helpers
from tests.cptv_writer import write_cptv
from tests.test_cptv import random_walk_clip

pytestmark = pytest.mark.gpu


def _device_decode(paths):
    from classifier_pipeline_b200 import engine
    from classifier_pipeline_b200.cptv import CptvReader, decode_clips_device

    eng = engine.get_engine()
    d_frames, first, frames = decode_clips_device(eng, [CptvReader(p) for p in paths])
    return d_frames.cpu().numpy(), first, frames


def test_reference_clips_match_host_decoder():
    from classifier_pipeline_b200.cptv import read_clip

    paths = [os.path.join(helpers.GOLDEN, "clips", n + ".cptv") for n in ("possum", "hedgehog")]
    pix, first, meta = _device_decode(paths)
    for i, p in enumerate(paths):
        _, frames = read_clip(p)
        assert first[i + 1] - first[i] == len(frames)
        assert np.array_equal(pix[first[i] : first[i + 1]], np.array([f.pix for f in frames]))
        assert [m.background_frame for m in meta[i]] == [f.background_frame for f in frames]
        assert [m.time_on for m in meta[i]] == [f.time_on for f in frames]


def test_every_bit_width(tmp_path):
    paths, clips = [], []
    for k, (step, force) in enumerate([(0, None), (1, None), (3, None), (40, None), (700, None), (20000, None), (2, 13), (5, 16), (1, 9)]):
        rng = np.random.default_rng(100 + k)
        frames = random_walk_clip(rng, 5 + k % 3, step)
        path = tmp_path / "c{}.cptv".format(k)
        write_cptv(path, frames, force_bit_width=force)
        paths.append(path)
        clips.append(frames)
    pix, first, _ = _device_decode(paths)
    for i, frames in enumerate(clips):
        assert np.array_equal(pix[first[i] : first[i + 1]], frames), i
