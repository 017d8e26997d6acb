"""Frame selection for classification: ``SegmentType``, ``SegmentHeader`` and ``get_segments``
(ml_tools/datasetstructures.py:25-35, 771-846, 972-1301 of the reference).

Host-side index bookkeeping only: it decides WHICH 25 frames of a track form a segment; the pixels
are produced by the preprocessing kernels.  The random draws are made in the reference's order
(``default_rng(seed)`` for shuffles / padding and the global ``np.random.shuffle`` of the masked
variant), so a seeded run selects the same frames.  Supported segment types: ALL_RANDOM_MASKED (the
default), ALL_RANDOM, ALL_RANDOM_NOMIN, IMPORTANT_RANDOM, ALL_SECTIONS, ALL_SEQUENTIAL and IMPORTANT_SEQUENTIAL; the training-only ELONGATION and
TOP_SEQUENTIAL types are not part of this path and TOP_RANDOM is broken in the reference itself.
"""
import logging
from enum import Enum

import numpy as np


class SegmentType(Enum):
    IMPORTANT_RANDOM = 0
    ALL_RANDOM = 1
    IMPORTANT_SEQUENTIAL = 2
    ALL_SEQUENTIAL = 3
    TOP_SEQUENTIAL = 4
    ALL_SECTIONS = 5
    TOP_RANDOM = 6
    ALL_RANDOM_NOMIN = 7
    ALL_RANDOM_MASKED = 8
    ELONGATION = 9


_RANDOM_TYPES = (SegmentType.IMPORTANT_RANDOM, SegmentType.ALL_RANDOM, SegmentType.ALL_RANDOM_NOMIN, SegmentType.TOP_RANDOM,
                 SegmentType.ALL_RANDOM_MASKED, None)


class SegmentHeader:
    """The frames of one classifier input (datasetstructures.py:771-846)."""

    def __init__(self, clip_id, track_id, start_frame, frames, weight, mass, label, regions, frame_indices=None,
                 movement_data=None, best_mass=False, top_mass=False, start_time=None, camera=None, location=None,
                 station_id=None, rec_time=None, source_file=None, filtered=False, track_median_mass=None):
        self.label = label
        self.filtered = filtered
        self.rec_time = rec_time
        self.location = location
        self.station_id = station_id
        self.movement_data = movement_data
        self.top_mass = top_mass
        self.best_mass = best_mass
        self.clip_id = clip_id
        self.track_id = track_id
        self.frame_numbers = np.uint16(frame_indices)
        self.start_time = start_time
        self.regions = regions
        self.start_frame = start_frame
        self.frames = np.uint16(frames)
        self.weight = np.float16(weight)
        self._mass = np.uint16(mass)
        self.camera = camera
        self._source_file = source_file
        self._track_median_mass = track_median_mass

    @property
    def track_median_mass(self):
        return self._track_median_mass

    @property
    def source_file(self):
        return self._source_file

    @property
    def mass(self):
        return self._mass

    @property
    def sample_weight(self):
        return self.weight

    @property
    def track_bounds(self):
        return self.regions

    @property
    def frame_indices(self):
        return self.frame_numbers

    def __repr__(self):
        return "SegmentHeader(track {} frames {})".format(self.track_id, list(self.frame_numbers))


def _usable_frames(regions, ffc_frames, skip_ffc, frame_min_mass, has_no_mass):
    keep = []
    for r in regions:
        if not (has_no_mass or r.mass > 0):
            continue
        if ffc_frames is not None and skip_ffc and r.frame_number in ffc_frames:
            continue
        if r.blank or r.width <= 0 or r.height <= 0:
            continue
        if not has_no_mass and frame_min_mass is not None and r.mass < frame_min_mass:
            continue
        keep.append(r.frame_number)
    return keep


# labels whose tracks keep their false-positive frames (datasetstructures.py:22)
FP_LABELS = ["other", "unidentified", "rain", "false-positive", "water", "insect"]


def get_segments(clip_id, track_id, start_frame, regions, segment_width=25, segment_frame_spacing=9, label=None,
                 segment_min_mass=None, ffc_frames=[], lower_mass=0, repeats=1, min_frames=None,
                 segment_types=[SegmentType.ALL_RANDOM_MASKED], max_segments=None, location=None, station_id=None,
                 camera=None, rec_time=None, source_file=None, dont_filter=False, skip_ffc=True, frame_min_mass=None,
                 fp_frames=None, repeat_frame_indices=True, min_segments=None, seed=None):
    """Returns (segments, filtered_stats) like the reference."""
    if min_frames is None:
        min_frames = segment_width / 4.0
    regions = np.asarray(regions, dtype=object)
    mass_history = np.uint16([r.mass for r in regions])
    stats = {"segment_mass": 0, "too short": 0}
    has_no_mass = np.sum(mass_history) == 0
    segments = []
    for segment_type in segment_types:
        if segment_type in (SegmentType.ELONGATION, SegmentType.TOP_SEQUENTIAL):
            raise NotImplementedError("{} segments are a training-time selection outside this path".format(segment_type))
        if segment_type == SegmentType.TOP_RANDOM:
            # the reference sorts the frames into a Python list and then fails on `frames - start_frame`
            # (datasetstructures.py:1120-1128, 1247): there is no behaviour to reproduce
            raise NotImplementedError("TOP_RANDOM segments raise TypeError in the reference")
        min_mass = None if segment_type == SegmentType.ALL_RANDOM_NOMIN else segment_min_mass
        usable = _usable_frames(regions, ffc_frames, skip_ffc, frame_min_mass, has_no_mass)
        if fp_frames is not None and label not in FP_LABELS:  # datasetstructures.py:1028
            usable = [f for f in usable if f not in fp_frames]
        if not usable:
            logging.warning("Nothing to load for %s - %s", clip_id, track_id)
            return [], stats
        usable = np.array(usable)
        min_mass = 1 if min_mass is None else min(min_mass, np.median(mass_history[usable - start_frame]))
        rng = np.random.default_rng(seed=seed)
        if len(usable) < min_frames and not min_segments:
            stats["too short"] += 1
            continue
        count = int(max(1, len(usable) // segment_frame_spacing))
        mask_length = 25
        if max_segments is not None and segment_type != SegmentType.ALL_SECTIONS:
            count = min(max_segments, count)
            mask_length = max(mask_length, len(usable) // count)
        masked = segment_type == SegmentType.ALL_RANDOM_MASKED
        randomised = segment_type in _RANDOM_TYPES
        pool = usable
        for _ in range(repeats):
            if masked:
                positions = np.arange(len(regions))
                frame_of = positions + start_frame
                available = np.full(len(regions), False)
                available[usable - start_frame] = True
            if not masked or len(usable) < 40:
                pool = usable.copy()
                if randomised:
                    rng.shuffle(pool)
            for i in range(count):
                if masked:
                    if len(usable) < 40:
                        pool = positions[available]
                    else:
                        window = available.copy()
                        window[i * mask_length : (i + 1) * mask_length] = False
                        pool = np.uint32(positions[window])
                        np.random.shuffle(pool)  # the reference uses the global generator here
                if len(pool) == 0 or min_segments is None or len(segments) >= min_segments:
                    if (len(pool) < segment_width / 2.0 and len(segments) > 0) or len(pool) < segment_width / 4:
                        break
                if segment_type == SegmentType.ALL_SECTIONS:
                    section = pool[: int(segment_width * 2.2)]
                    picks = rng.choice(len(section), min(segment_width, len(section)), replace=False)
                    frames = section[picks]
                    pool = pool[segment_width:]
                elif masked:
                    picks = pool[:segment_width]
                    available[picks] = False
                    frames = frame_of[picks]
                elif randomised:
                    frames = pool[:segment_width]
                    pool = pool[segment_width:]
                else:
                    lo = i * segment_frame_spacing
                    frames = pool[lo : min(len(pool), lo + segment_width)]
                short = segment_width - len(frames)
                if short > 0:
                    frames = np.concatenate([frames, rng.choice(frames, min(short, len(frames)), replace=False)])
                frames.sort()
                rel = frames - start_frame
                seg_mass = np.sum(mass_history[rel])
                avg_mass = seg_mass / len(rel)
                filtered = False
                if min_mass and avg_mass < min_mass:
                    if not dont_filter:
                        stats["segment_mass"] += 1
                        continue
                    filtered = True
                seg_regions = regions[rel]
                weight = 0.75 if avg_mass < 50 else (1 if avg_mass < 100 else 1.2)
                if repeat_frame_indices and len(frames) < segment_width:
                    frames = sorted(list(frames) + list(rng.choice(frames, segment_width - len(frames))))
                segments.append(SegmentHeader(clip_id, track_id, start_frame=start_frame, frames=segment_width, weight=weight,
                                              mass=seg_mass, label=label, regions=seg_regions, frame_indices=frames,
                                              camera=camera, location=location, station_id=station_id, rec_time=rec_time,
                                              source_file=source_file, filtered=filtered))
    return segments, stats
