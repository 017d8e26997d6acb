"""``Track`` and ``RegionTracker``: host-side track bookkeeping (track/track.py:35-1031).

BASELINE north_star keeps track-to-region matching and the Kalman bookkeeping on the host; the
device hands over compact region lists per frame.  Class, attribute and method names follow
the reference so that callers (and the tracks JSON) are unchanged.  Behaviour that looks odd
but is load-bearing for parity is kept and marked ``# parity:``.
"""
import logging
import math
from collections import namedtuple

import numpy as np

from ..ml_tools.tools import eucl_distance_sq
from .kalman import Kalman
from .region import Region

TrackMovementStatistics = namedtuple(
    "TrackMovementStatistics",
    "movement max_offset score average_mass median_mass delta_std region_jitter jitter_smaller jitter_bigger "
    "blank_percent frames_moved mass_std, average_velocity",
)
TrackMovementStatistics.__new__.__defaults__ = (0,) * len(TrackMovementStatistics._fields)


class RegionTracker:
    """Per-track matcher state (track.py:35-323)."""

    MIN_KALMAN_FRAMES = 18
    MASS_CHANGE_PERCENT = 0.55
    BASE_DISTANCE_CHANGE = 11250
    MIN_MASS_CHANGE = 20 * 4
    RESTRICT_MASS_AFTER = 1.5
    MAX_DISTANCE = 30752
    BASE_VELOCITY = 8
    VELOCITY_MULTIPLIER = 10

    def __init__(self, id, tracking_config, crop_rectangle=None):
        self.track_id = id
        self.clear_run = 0
        self.kalman_tracker = Kalman()
        self._frames_since_target_seen = 0
        self.frames = 0
        self._blank_frames = 0
        self._last_bound = None
        self.crop_rectangle = crop_rectangle
        self._tracking = False
        self.type = tracking_config.type
        p = tracking_config.params
        self.min_mass_change = p.get("min_mass_change", RegionTracker.MIN_MASS_CHANGE)
        self.max_distance = p.get("max_distance", RegionTracker.MAX_DISTANCE)
        self.base_distance_change = p.get("base_distance_change", RegionTracker.BASE_DISTANCE_CHANGE)
        self.restrict_mass_after = p.get("restrict_mass_after", RegionTracker.RESTRICT_MASS_AFTER)
        self.mass_change_percent = p.get("mass_change_percent", RegionTracker.MASS_CHANGE_PERCENT)
        self.velocity_multiplier = p.get("velocity_multiplier", RegionTracker.VELOCITY_MULTIPLIER)
        self.base_velocity = p.get("base_velocity", RegionTracker.BASE_VELOCITY)
        self.max_blanks = p.get("max_blanks", 18)
        self.predicted_mid = (0.0, 0.0)

    tracking = property(lambda self: self._tracking)
    last_bound = property(lambda self: self._last_bound)
    blank_frames = property(lambda self: self._blank_frames)
    frames_since_target_seen = property(lambda self: self._frames_since_target_seen)
    nonblank_frames = property(lambda self: self.frames - self._blank_frames)

    def get_size_change(self, current_area, region):
        return abs(region.area - current_area) / (current_area + 50)

    def match(self, regions, track):
        """Score every region this track may continue into: list of (score, track, region)."""
        scores = []
        avg_mass = track.average_mass()
        avg_area = track.average_area()
        max_distance = self.get_max_distance_change(track)[0]
        max_mass_change = self.get_max_mass_change_percent(track, avg_mass)
        for region in regions:
            tl, _, br = self.last_bound.average_distance(region)
            # parity: the reference tests the *builtin* ``type`` against "thermal"/"ir" (track.py:140,185),
            # so every tracker scores with the mean of the two corner distances
            distance = (tl + br) / 2
            if max_mass_change and abs(avg_mass - region.mass) > max_mass_change:
                continue
            if max_distance is not None and distance > max_distance:
                continue
            if self.get_size_change(avg_area, region) > get_max_size_change(track, region):
                continue
            scores.append((distance, track, region))
        return scores

    def add_region(self, region):
        self.frames += 1
        if region.blank:
            self._blank_frames += 1
            self._frames_since_target_seen += 1
            limit = min(2 * (self.frames - self._frames_since_target_seen), self.max_blanks)
            self._tracking = self._frames_since_target_seen < limit
        else:
            if self._frames_since_target_seen != 0:
                self.clear_run = 0
            self.clear_run += 1
            self._tracking = True
            self.kalman_tracker.correct(region)
            self._frames_since_target_seen = 0
        prediction = self.kalman_tracker.predict()
        self.predicted_mid = (prediction[0][0], prediction[1][0])
        self._last_bound = region

    def predicted_velocity(self):
        if self.last_bound is None or self.nonblank_frames <= RegionTracker.MIN_KALMAN_FRAMES:
            return (0, 0)
        return (self.predicted_mid[0] - self.last_bound.centroid[0], self.predicted_mid[1] - self.last_bound.centroid[1])

    def add_blank_frame(self):
        """A placeholder region: Kalman-predicted once the track is old enough, else the last box."""
        last = self.last_bound
        if self.frames - RegionTracker.MIN_KALMAN_FRAMES - self._frames_since_target_seen * 2 > 0:
            region = Region(
                int(self.predicted_mid[0] - last.width / 2.0), int(self.predicted_mid[1] - last.height / 2.0),
                last.width, last.height, centroid=[self.predicted_mid[0], self.predicted_mid[1]],
            )
            if self.crop_rectangle:
                region.crop(self.crop_rectangle)
        else:
            region = last.copy()
        region.blank = True
        region.mass = 0
        region.pixel_variance = 0
        region.frame_number = last.frame_number + 1
        self.add_region(region)
        return region

    def get_max_distance_change(self, track):
        vx, vy = track.velocity
        if len(track) == 1:
            vx = vy = self.base_velocity
        vx *= self.velocity_multiplier
        vy *= self.velocity_multiplier
        velocity_distance = vx * vx + vy * vy
        px, py = track.predicted_velocity()
        allowed = self.base_distance_change + max(velocity_distance, px * px + py * py)
        return [allowed, None, allowed]

    def get_max_mass_change_percent(self, track, average_mass):
        if self.mass_change_percent is None or len(track) <= self.restrict_mass_after * track.fps:
            return None
        percent = self.mass_change_percent
        if np.sum(np.abs(track.velocity)) > 5:
            percent = percent + 0.1
        return max(self.min_mass_change, average_mass * percent)


def get_max_size_change(track, region):
    """Allowed relative area change for a match (track.py:326-341)."""
    last_on_border = track.last_bound.is_along_border
    crossing = (region.is_along_border and not last_on_border) or last_on_border  # exiting or entering
    fast = np.sum(np.abs(track.velocity)) > 10
    if crossing:
        return 6 if fast else 2
    percent = 2 if len(track) < 5 else 1.5
    return percent * 2 if fast else percent


class ThumbInfo:
    def __init__(self, track_id):
        self.points = -1
        self.region = None
        self.thumb = None
        self.thumb_frame = None
        self.last_frame_check = None
        self.predicted_tag = None
        self.predicted_confidence = None
        self.track_id = track_id

    def score(self):
        score = self.points
        unit = 100000
        if self.predicted_tag is not None:
            if self.predicted_tag != "false-positive":
                score += 1000 * unit
                confidence = self.predicted_confidence if self.predicted_confidence > 80 else 0
            else:
                confidence = 100 - self.predicted_confidence
            score += confidence * unit
        return score

    def to_metadata(self):
        return {"region": self.region.meta_dictionary(), "contours": self.points, "score": round(self.score())}


class Track:
    """Bounds of one tracked object over time (track.py:372-1031)."""

    _track_id = 1  # parity: class-global id counter, reset by every new Clip (clip.py:57-59)
    JITTER_THRESHOLD = 0.3
    MIN_JITTER_CHANGE = 5

    def __init__(self, clip_id, id=None, fps=9, tracking_config=None, crop_rectangle=None, tracker_version=None):
        if not id:
            self._id = Track._track_id
            Track._track_id += 1
        else:
            self._id = id
        self.clip_id = clip_id
        self.in_trap = False
        self.received_at = None
        self.trap_reported = False
        self.trigger_frame = None
        self.direction = 0
        self.trap_tag = None
        self.start_frame = None
        self.start_s = None
        self.end_s = None
        self.fps = fps
        self.current_frame_num = None
        self.frame_list = []
        self.bounds_history = []
        self.vel_x = []
        self.vel_y = []
        self.tag = "unknown"
        self.prev_frame_num = None
        self.confidence = None
        self.max_novelty = None
        self.avg_novelty = None
        self.from_metadata = False
        self.tags = None
        self.predictions = None
        self.predicted_class = None
        self.predicted_tag = None
        self.predicted_confidence = None
        self.all_class_confidences = None
        self.prediction_classes = None
        self.crop_rectangle = crop_rectangle
        self.tracker_version = tracker_version
        self.tracker = self.get_tracker(tracking_config) if tracking_config is not None else None
        self.thumb_info = None
        self.score = None
        self.stats = None

    def get_tracker(self, tracking_config):
        if tracking_config.tracker == "RegionTracker":
            return RegionTracker(self.get_id(), tracking_config, self.crop_rectangle)
        raise Exception(f"Cant find for tracker {tracking_config.tracker}")

    @classmethod
    def from_region(cls, clip, region, tracker_version=None, tracking_config=None):
        track = cls(clip.get_id(), fps=clip.frames_per_second, tracker_version=tracker_version,
                    crop_rectangle=clip.crop_rectangle, tracking_config=tracking_config)
        track.start_frame = region.frame_number
        track.start_s = region.frame_number / float(clip.frames_per_second)
        track.add_region(region)
        return track

    # ------------------------------------------------------------------ simple views
    def get_id(self):
        return self._id

    @property
    def blank_frames(self):
        return 0 if self.tracker is None else self.tracker.blank_frames

    @property
    def tracking(self):
        return self.tracker.tracking

    @property
    def frames_since_target_seen(self):
        return self.tracker.frames_since_target_seen

    @property
    def end_frame(self):
        return self.bounds_history[-1].frame_number if self.bounds_history else self.start_frame

    @property
    def nonblank_frames(self):
        return self.end_frame + 1 - self.start_frame - self.blank_frames

    @property
    def frames(self):
        return self.end_frame + 1 - self.start_frame

    @property
    def last_mass(self):
        return self.bounds_history[-1].mass

    @property
    def velocity(self):
        return self.vel_x[-1], self.vel_y[-1]

    @property
    def last_bound(self):
        return self.bounds_history[-1]

    def __len__(self):
        return len(self.bounds_history)

    def __repr__(self):
        return "Track: {} frames# {}".format(self.get_id(), len(self))

    # ------------------------------------------------------------------ growth
    def match(self, regions):
        return self.tracker.match(regions, self)

    def predicted_velocity(self):
        return self.tracker.predicted_velocity()

    def add_region(self, region):
        if self.prev_frame_num and region.frame_number:
            for _ in range(region.frame_number - self.prev_frame_num - 1):
                self.add_blank_frame()
        self.tracker.add_region(region)
        self.bounds_history.append(region)
        self.prev_frame_num = region.frame_number
        self.update_velocity()

    def add_blank_frame(self):
        region = self.tracker.add_blank_frame()
        self.bounds_history.append(region)
        self.prev_frame_num = region.frame_number
        self.update_velocity()

    def update_velocity(self):
        if len(self.bounds_history) >= 2:
            a, b = self.bounds_history[-2].centroid, self.bounds_history[-1].centroid
            self.vel_x.append(b[0] - a[0])
            self.vel_y.append(b[1] - a[1])
        else:
            self.vel_x.append(0)
            self.vel_y.append(0)

    def crop_regions(self):
        if self.crop_rectangle is None:
            logging.info("No crop rectangle to crop with")
            return
        for region in self.bounds_history:
            region.crop(self.crop_rectangle)

    def _recent_average(self, value):
        """Mean of ``value(bound)`` over the last five non-blank bounds (track.py:707-735)."""
        total = count = 0
        for bound in reversed(self.bounds_history):
            if not bound.blank:
                total += value(bound)
                count += 1
                if count == 5:
                    break
        return total / count if count else 0

    def average_area(self):
        return self._recent_average(lambda b: b.area)

    def average_mass(self):
        return self._recent_average(lambda b: b.mass)

    # ------------------------------------------------------------------ end-of-clip
    def trim(self):
        """Drop near-empty frames from both ends (track.py:877-905)."""
        masses = [int(b.mass) for b in self.bounds_history]
        cutoff = max(0.005 * np.median(masses), 2)
        start = 0
        while start < len(self) and masses[start] <= cutoff:
            start += 1
        end = len(self) - 1
        while end > 0 and masses[end] <= cutoff:
            if self.tracker and self.frames_since_target_seen > 0:
                self.tracker._frames_since_target_seen -= 1
                self.tracker._blank_frames -= 1
            end -= 1
        if end < start:
            self.bounds_history = []
            self.vel_x = []
            self.vel_y = []
            if self.tracker:
                self.tracker._blank_frames = 0
        else:
            self.start_frame += start
            self.bounds_history = self.bounds_history[start : end + 1]
            self.vel_x = self.vel_x[start : end + 1]
            self.vel_y = self.vel_y[start : end + 1]
        self.start_s = self.start_frame / float(self.fps)

    def set_end_s(self, fps):
        self.end_s = self.start_s if len(self) == 0 else (self.end_frame + 1) / fps

    def calculate_stats(self):
        """Movement / jitter / delta statistics and the tracking score (track.py:737-840)."""
        if len(self) <= 1:
            self.stats = TrackMovementStatistics()
            return
        bounds = self.bounds_history
        non_blank = [b for b in bounds if not b.blank]
        mass_history = [int(b.mass) for b in non_blank]
        variance_history = [b.pixel_variance for b in non_blank if b.pixel_variance]
        movement = 0
        max_offset = 0
        frames_moved = 0
        avg_vel = 0
        origin = bounds[0].mid
        for i, (vx, vy) in enumerate(zip(self.vel_x, self.vel_y)):
            region = bounds[i]
            if not region.blank:
                avg_vel += abs(vx) + abs(vy)
            if i == 0 or region.blank or bounds[i - 1].blank:
                continue
            if region.has_moved(bounds[i - 1]) or region.is_along_border:
                movement += (vx**2 + vy**2) ** 0.5
                max_offset = max(max_offset, eucl_distance_sq(origin, region.mid))
                frames_moved += 1
        avg_vel = avg_vel / len(mass_history)
        max_offset = math.sqrt(max_offset)
        delta_std = float(np.mean(variance_history)) ** 0.5
        jitter_bigger = jitter_smaller = 0
        for prev, bound in zip(bounds, bounds[1:]):
            if prev.is_along_border or bound.is_along_border:
                continue
            dh = bound.height - prev.height
            dw = prev.width - bound.width  # parity: width difference has the opposite sign (track.py:786)
            if abs(dh) > max(Track.MIN_JITTER_CHANGE, prev.height * Track.JITTER_THRESHOLD):
                if dh > 0:
                    jitter_bigger += 1
                else:
                    jitter_smaller += 1
            elif abs(dw) > max(Track.MIN_JITTER_CHANGE, prev.width * Track.JITTER_THRESHOLD):
                if dw > 0:
                    jitter_bigger += 1
                else:
                    jitter_smaller += 1
        jitter_percent = int(round(100 * (jitter_bigger + jitter_smaller) / float(self.frames)))
        blank_percent = int(round(100.0 * self.blank_frames / self.frames))
        score = (min((movement**0.5) + max_offset, 100) + min(delta_std * 25.0, 100) + (100 - jitter_percent)
                 + (100 - blank_percent))
        self.stats = TrackMovementStatistics(
            movement=float(movement), max_offset=float(max_offset), average_mass=float(np.mean(mass_history)),
            median_mass=float(np.median(mass_history)), delta_std=float(delta_std), score=float(score),
            region_jitter=jitter_percent, jitter_bigger=jitter_bigger, jitter_smaller=jitter_smaller,
            blank_percent=blank_percent, frames_moved=frames_moved, mass_std=float(np.std(mass_history)),
            average_velocity=float(avg_vel),
        )

    def smooth(self, frame_bounds):
        """Three-frame moving average of the box size around each centroid (track.py:842-875)."""
        if len(self.bounds_history) == 0:
            return
        hist = self.bounds_history
        hist[1]  # parity: a one-region track raises IndexError in the reference (track.py:853)
        smoothed = []
        last = len(hist) - 1
        for i, cur in enumerate(hist):
            prev, nxt = hist[max(0, i - 1)], hist[min(last, i + 1)]
            w = (prev.width + cur.width + nxt.width) / 3
            h = (prev.height + cur.height + nxt.height) / 3
            box = Region(int(cur.centroid[0] - w / 2), int(cur.centroid[1] - h / 2), int(w), int(h))
            box.crop(frame_bounds)
            smoothed.append(box)
        self.bounds_history = smoothed

    def get_overlap_ratio(self, other_track, threshold=0.05):
        if len(self) == 0 or len(other_track) == 0:
            return 0.0
        first = max(self.start_frame, other_track.start_frame)
        last = min(self.end_frame, other_track.end_frame)
        overlapped = 0
        for pos in range(first, last + 1):
            i, j = pos - self.start_frame, pos - other_track.start_frame
            if 0 <= i < len(self) and 0 <= j < len(other_track):
                ours = self.bounds_history[i]
                if ours.area == 0:
                    continue
                if ours.overlap_area(other_track.bounds_history[j]) / ours.area >= threshold:
                    overlapped += 1
        return overlapped / len(self)

    def update_trapped_state(self):
        if self.in_trap:
            return True
        if len(self.bounds_history) < 2:
            return False
        self.in_trap = all(r.in_trap for r in self.bounds_history[-2:])
        return self.in_trap

    # ------------------------------------------------------------------ segments (classifier input selection)
    def get_segments(self, segment_width, segment_frame_spacing=9, repeats=1, min_frames=0, segment_frames=None,
                     segment_types=None, from_last=None, max_segments=None, ffc_frames=None, dont_filter=False,
                     filter_by_fp=False, min_segments=1, seed=None):
        """Frame-number groups to classify (track.py:480-545).  Explicit ``segment_frames`` are honoured
        exactly; otherwise ``segments.get_segments`` draws them with a seeded generator."""
        from ..ml_tools.segments import SegmentHeader, get_segments

        if from_last is not None:
            if from_last == 0:
                return []
            regions = np.array(self.bounds_history[-from_last:], dtype=object)
            start_frame = regions[0].frame_number
        else:
            start_frame = self.start_frame
            regions = np.array(self.bounds_history, dtype=object)
        if segment_frames is not None:
            masses = np.uint16([r.mass for r in regions])
            out = []
            for frames in segment_frames:
                rel = np.asarray(frames) - self.start_frame
                out.append(SegmentHeader(self.clip_id, self._id, start_frame=start_frame, frames=len(frames), weight=1,
                                         mass=np.sum(masses[rel]), label=None, regions=regions[rel], frame_indices=frames))
            return out
        from ..ml_tools.segments import SegmentType

        segments, _ = get_segments(self.clip_id, self._id, start_frame, segment_frame_spacing=segment_frame_spacing,
                                   segment_width=segment_width, regions=regions, ffc_frames=ffc_frames, repeats=repeats,
                                   min_frames=min_frames,
                                   segment_types=[SegmentType.ALL_RANDOM] if segment_types is None else segment_types,
                                   max_segments=max_segments, dont_filter=dont_filter, min_segments=min_segments, seed=seed)
        return segments

    # ------------------------------------------------------------------ (de)serialisation
    def start_and_end_in_secs(self):
        if self.end_s is None:
            self.end_s = self.start_s if len(self) == 0 else (self.end_frame + 1) / self.fps
        return (self.start_s, self.end_s)

    def get_metadata(self, predictions_per_model=None):
        """Key order and rounding of track.py:1001-1031."""
        start_s, end_s = self.start_and_end_in_secs()
        info = {"id": self.get_id()}
        if self.in_trap:
            info["trap_triggered"] = self.in_trap
            info["trigger_frame"] = self.trigger_frame
            if self.trap_tag is not None:
                info["trap_tag"] = self.trap_tag
        info["tracker_version"] = self.tracker_version
        info["start_s"] = round(start_s, 2)
        info["end_s"] = round(end_s, 2)
        info["num_frames"] = len(self)
        info["frame_start"] = self.start_frame
        info["frame_end"] = self.end_frame
        info["positions"] = self.bounds_history
        if self.thumb_info is not None:
            info["thumbnail"] = self.thumb_info.to_metadata()
        info["tracking_score"] = 0 if self.stats is None else self.stats.score
        predictions = []
        for model_id, model_predictions in (predictions_per_model or {}).items():
            prediction = model_predictions.prediction_for(self.get_id())
            if prediction is None:
                continue
            meta = prediction.get_metadata(model_predictions.thresholds)
            meta["model_id"] = model_id
            predictions.append(meta)
        info["predictions"] = predictions
        return info

    def load_track_meta(self, track_meta, frames_per_second, tag_precedence=None, min_confidence=0.8):
        """Rebuild a track from its JSON (track.py:571-638)."""
        self.tracker_version = track_meta.get("tracker_version", "unknown")
        self.from_metadata = True
        self._id = track_meta["id"]
        extra = track_meta.get("data", track_meta)
        if "start_s" in extra:
            self.start_s, self.end_s = extra["start_s"], extra["end_s"]
        else:
            self.start_s, self.end_s = extra["start"], extra["end"]
        self.fps = frames_per_second
        self.tags = track_meta.get("tags")
        tag = Track.get_best_human_tag(self.tags, tag_precedence, min_confidence)
        if tag:
            self.tag = tag["what"]
            self.confidence = tag["confidence"]
        self.stats = TrackMovementStatistics(score=track_meta.get("tracking_score", 0))
        positions = track_meta.get("positions")
        if not positions:
            return False
        self.bounds_history = []
        self.frame_list = []
        for i, position in enumerate(positions):
            if isinstance(position, list):
                region = Region.region_from_array(position[1])
                if region.frame_number is None:
                    region.frame_number = round(position[0] * frames_per_second)
            else:
                region = Region.region_from_json(position)
                if region.frame_number is None:
                    if "frameTime" not in position:
                        raise Exception("No frame number info for track")
                    region.frame_number = position["frameTime"] * 9 if i == 0 else self.bounds_history[0].frame_number + i
            if self.start_frame is None:
                self.start_frame = region.frame_number
            self.bounds_history.append(region)
            self.frame_list.append(region.frame_number)
        self.current_frame_num = 0
        return True

    @classmethod
    def get_best_human_tag(cls, track_tags, tag_precedence, min_confidence=-1):
        if track_tags is None:
            return None
        candidates = [t for t in track_tags if not t.get("automatic", False) and t.get("confidence") >= min_confidence]
        if not candidates:
            return None
        precedence = tag_precedence or {}
        default = precedence.get("default", 100)
        best_tag, best_rank = None, None
        for candidate in candidates:
            rank = cls.tag_ranking(candidate, precedence, default)
            if best_tag and rank == best_rank:
                if is_conflicting_tag(best_tag, candidate):
                    best_tag = None
                elif len(candidate.get("path")) > len(best_tag.get("path")):
                    best_tag = candidate
            elif best_rank is None or rank < best_rank:
                best_rank, best_tag = rank, candidate
        return best_tag

    @staticmethod
    def tag_ranking(track_tag, precedence, default_prec):
        return precedence.get(track_tag.get("what"), default_prec) + 1 - track_tag.get("confidence", 0)


def is_conflicting_tag(tag_one, tag_two):
    a, b = tag_one.get("path"), tag_two.get("path")
    return tag_one["what"] != tag_two["what"] and not (a in b or b in a)
