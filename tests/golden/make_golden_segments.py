#!/usr/bin/env python
"""Golden fixture for the frame selection (S1): the UNMODIFIED reference ``ml_tools.datasetstructures.get_segments``
(datasetstructures.py:972-1301) on seeded region lists, for the segment types this repo builds.  Build container only:

    python tests/golden/make_golden_segments.py      # writes tests/golden/segments.json
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402
from tests.segment_helpers import CASES, make_regions  # noqa: E402


def main():
    ref_harness.setup()
    from ml_tools.datasetstructures import SegmentType, get_segments

    out = []
    for case in CASES:
        regions = make_regions(case["n"], case["start"], case["region_seed"], case.get("blank_every"), case.get("zero_mass_every"))
        np.random.seed(case["np_seed"])
        segs, stats = get_segments("clip", 1, case["start"], np.array(regions, dtype=object), segment_width=25,
                                   segment_frame_spacing=case.get("spacing", 9), segment_types=[SegmentType[case["type"]]],
                                   max_segments=case.get("max_segments"), min_segments=case.get("min_segments"),
                                   ffc_frames=case.get("ffc_frames", []), dont_filter=case.get("dont_filter", False),
                                   seed=case["seed"], repeats=case.get("repeats", 1))
        out.append(dict(case=case, frames=[[int(f) for f in s.frame_indices] for s in segs], mass=[int(s.mass) for s in segs],
                        weight=[float(s.weight) for s in segs], stats=stats))
        print(case["type"], case["n"], "->", len(segs), "segments")
    json.dump(out, open(os.path.join(HERE, "segments.json"), "w"))


if __name__ == "__main__":
    main()
