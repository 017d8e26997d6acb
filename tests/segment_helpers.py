"""Seeded region lists + cases shared by the segment-selection fixture generator and its test."""
import types

import numpy as np

CASES = [
    dict(type="ALL_RANDOM_MASKED", n=30, start=5, region_seed=1, seed=11, np_seed=3),
    dict(type="ALL_RANDOM_MASKED", n=90, start=11, region_seed=2, seed=12, np_seed=4, blank_every=7),
    dict(type="ALL_RANDOM_MASKED", n=200, start=0, region_seed=3, seed=13, np_seed=5, max_segments=3, zero_mass_every=9),
    dict(type="ALL_RANDOM_MASKED", n=8, start=2, region_seed=4, seed=14, np_seed=6, min_segments=1),
    dict(type="ALL_RANDOM", n=120, start=40, region_seed=5, seed=15, np_seed=7, blank_every=5),
    dict(type="ALL_RANDOM", n=45, start=0, region_seed=6, seed=16, np_seed=8, ffc_frames=[3, 4, 5], min_segments=1),
    dict(type="ALL_RANDOM_NOMIN", n=60, start=9, region_seed=7, seed=17, np_seed=9),
    dict(type="ALL_SECTIONS", n=150, start=20, region_seed=8, seed=18, np_seed=10),
    dict(type="ALL_SEQUENTIAL", n=70, start=0, region_seed=9, seed=19, np_seed=11, spacing=12),
    dict(type="IMPORTANT_RANDOM", n=50, start=0, region_seed=11, seed=21, np_seed=13, dont_filter=True),
]


def make_regions(n, start, seed, blank_every=None, zero_mass_every=None):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        blank = bool(blank_every and i % blank_every == blank_every - 1)
        mass = 0 if (zero_mass_every and i % zero_mass_every == 0) else int(rng.integers(1, 300))
        out.append(types.SimpleNamespace(frame_number=start + i, mass=mass, blank=blank, width=int(rng.integers(1, 40)),
                                         height=int(rng.integers(1, 40))))
    return out
