"""CPTV v2 decoder (K0): writer -> host reader round trips over every bit width, and the reference's two clips."""
import os

import numpy as np
import pytest

from tests import helpers
from tests.cptv_writer import write_cptv


def random_walk_clip(rng, T, step, H=120, W=160):
    base = rng.integers(1000, 60000, size=(H, W)).astype(np.int64)
    frames = [base]
    for _ in range(T - 1):
        frames.append(np.clip(frames[-1] + rng.integers(-step, step + 1, size=(H, W)), 0, 65535))
    return np.array(frames, dtype=np.uint16)


@pytest.mark.parametrize("step,force", [(0, None), (1, None), (3, None), (40, None), (700, None), (20000, None), (2, 13), (5, 16)])
def test_writer_reader_round_trip(tmp_path, step, force):
    from classifier_pipeline_b200.cptv import CptvReader

    rng = np.random.default_rng(step + 1)
    frames = random_walk_clip(rng, 6, step)
    path = tmp_path / "clip.cptv"
    widths = write_cptv(path, frames, model="lepton3", force_bit_width=force)
    reader = CptvReader(path)
    assert reader.get_header().model == "lepton3"
    got = np.array([f.pix for f in reader])
    assert np.array_equal(got, frames)
    assert max(widths) <= 17


def test_reference_clips_decode():
    from classifier_pipeline_b200.cptv import read_clip

    header, frames = read_clip(os.path.join(helpers.GOLDEN, "clips", "possum.cptv"))
    assert (header.x_resolution, header.y_resolution, header.model) == (160, 120, "lepton3")
    assert len(frames) == 161 and frames[0].background_frame and not frames[1].background_frame
    header, frames = read_clip(os.path.join(helpers.GOLDEN, "clips", "hedgehog.cptv"))
    assert len(frames) == 119 and not frames[0].background_frame
