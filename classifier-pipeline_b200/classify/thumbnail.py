"""Thumbnail choice for tracks and trackless clips (classify/thumbnail.py:13-200): the frame of a track whose region has
the most mass, the most contour points and the warmest animal wins.

Everything here is per-track bookkeeping on regions the device extraction produced: the label image of each frame
(``Frame.mask``) is already on the host with the clip.  The number of contour points comes from OpenCV's
``findContours(..., CHAIN_APPROX_TC89_L1)`` -- the Teh-Chin approximation is third-party arithmetic that is used as-is
(cv2 must be importable), like the Kalman filter of the matcher.
"""
import logging
from collections import namedtuple

import numpy as np

from ..ml_tools import tools
from ..ml_tools.imageprocessing import normalize
from ..track.region import Region

Stat = namedtuple("Stat", "region contours median_diff")


def _cv2():
    try:
        import cv2
    except ImportError as e:  # pragma: no cover
        raise ImportError("thumbnail scoring counts contour points with cv2.findContours (third party); install OpenCV") from e
    return cv2


def contour_points(sub_mask):
    """Points of the largest external contour of a region's mask, Teh-Chin L1 approximation (thumbnail.py:95-108).
    None when there is no contour."""
    cv2 = _cv2()
    contours, _ = cv2.findContours(np.uint8(sub_mask), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_TC89_L1)
    if len(contours) == 0:
        return None
    return max(len(c) for c in contours)


def best_trackless_thumb(clip):
    """Choose a frame for clips without any track (thumbnail.py:13-60)."""
    best_region = None
    THUMBNAIL_SIZE = 64
    for regions in clip.region_history:
        for region in regions:
            if best_region is None or region.mass > best_region.mass:
                best_region = region
    if best_region is not None:
        return best_region
    best_frame_i = np.argmax(clip.stats.frame_stats_mean)
    best_frame = clip.frame_buffer.get_frame(best_frame_i).thermal
    frame_height, frame_width = best_frame.shape
    best_filtered = best_frame - clip.background
    # the means of every 64x64 window at once (summed-area tables); the choice below walks them in the reference's order
    def window_means(img):
        sat = np.zeros((frame_height + 1, frame_width + 1), np.float64)
        sat[1:, 1:] = np.cumsum(np.cumsum(np.float64(img), axis=0), axis=1)
        s = THUMBNAIL_SIZE
        return (sat[s:, s:] - sat[:-s, s:] - sat[s:, :-s] + sat[:-s, :-s]) / (s * s)

    t_means, f_means = window_means(best_frame), window_means(best_filtered)
    best_region = None
    for y in range(frame_height - THUMBNAIL_SIZE):
        for x in range(frame_width - THUMBNAIL_SIZE):
            thermal_sum, filtered_sum = t_means[y, x], f_means[y, x]
            if best_region is None:
                best_region = ((x, y), filtered_sum, thermal_sum)
            elif best_region[1] > 0:
                if best_region[1] < filtered_sum:
                    best_region = ((x, y), thermal_sum, filtered_sum)
            elif best_region[2] < thermal_sum:
                best_region = ((x, y), thermal_sum, filtered_sum)
    centroid = (best_region[0][0] + THUMBNAIL_SIZE // 2, best_region[0][1] + THUMBNAIL_SIZE // 2)
    return Region(best_region[0][0], best_region[0][1], THUMBNAIL_SIZE, THUMBNAIL_SIZE, frame_number=best_frame_i, centroid=centroid)


def get_track_thumb_stats(clip, track):
    max_mass = 0
    max_median_diff = 0
    min_median_diff = 0
    max_contour = 0
    stats = []
    for region in track.bounds_history:
        if region.blank or region.mass == 0:
            continue
        frame = clip.frame_buffer.get_frame(region.frame_number)
        if frame is None:
            continue
        if frame.mask is None:
            logging.info("Doing contours by filtered")
            contour_image, _ = normalize(frame.filtered, new_max=255)
        else:
            contour_image = frame.mask
        points = contour_points(region.subimage(contour_image))
        if points is None:
            continue
        if points > max_contour:
            max_contour = points
        # the thermal values of the pixels that are considered animal, against the frame's median
        sub_mask = region.subimage(contour_image) > 0
        masked_thermal = region.subimage(frame.thermal)[sub_mask]
        median_diff = np.median(masked_thermal) - np.median(frame.thermal)
        if region.mass > max_mass:
            max_mass = region.mass
        if median_diff > max_median_diff:
            max_median_diff = median_diff
        if median_diff < min_median_diff:
            min_median_diff = median_diff
        stats.append(Stat(region, points, median_diff))
    return stats, max_mass, max_median_diff, min_median_diff, max_contour


def get_thumbnail_info(clip, track):
    stats, max_mass, max_median_diff, min_median_diff, max_contour = get_track_thumb_stats(clip, track)
    if len(stats) == 0:
        if len(track.bounds_history) == 0:
            return None, 0
        return Stat(track.bounds_history[0], 0, 0), 0
    scored_frames = sorted(stats, key=lambda s: score(s, max_mass, max_median_diff, min_median_diff, max_contour), reverse=True)
    best_score = score(scored_frames[0], max_mass, max_median_diff, min_median_diff, max_contour)
    return scored_frames[0], best_score


def score(stat, max_mass, max_median_diff, min_median_diff, max_contour):
    region = stat.region
    mass_percent = region.mass / max_mass * 40      # mass out of 40
    pts = stat.contours / max_contour * 50          # contours out of 50
    centroid_mid = tools.eucl_distance_sq(region.centroid, region.mid) ** 0.5 * 2
    if max_median_diff == 0:
        diff = 0
        if min_median_diff != 0:
            diff = (stat.median_diff + abs(min_median_diff)) / abs(min_median_diff) * 40
    else:
        diff = stat.median_diff / max_median_diff * 40
    total = mass_percent + pts + diff - centroid_mid
    # prefer frames not on the border
    if region.x <= 1 or region.y <= 1 or region.bottom >= 119 or region.right >= 159:
        total = total - 1000
    return total
