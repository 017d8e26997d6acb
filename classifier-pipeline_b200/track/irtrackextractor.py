"""The detection half of the IR (640x480) tracker (track/irtrackextractor.py:324-389,391-492,789-818): components of the
background-subtracted frame through ``detect_objects_ir`` on the device, then the rectangle merging that glues the
fragments of one animal together.  (The reference's ``IRTrackExtractor._process_frame`` itself cannot run as written --
with ``DO_SALIENCY = False`` it evaluates ``np.amin(None)`` -- so only the pieces with defined behaviour are mirrored.)"""
from ..ml_tools.imageprocessing import detect_objects_ir
from ..ml_tools.tools import eucl_distance_sq


def rect_distance(r_a, r_b):
    """Gap between two [x, y, w, h, ...] rectangles: 0 along an axis on which they overlap (irtrackextractor.py:789-818)."""
    x_1 = x_2 = y_1 = y_2 = 0
    if r_a[2] + r_b[2] > max(r_a[0] + r_a[2], r_b[2] + r_b[0]) - min(r_a[0], r_b[0]):
        pass
    elif r_a[0] < r_b[0]:
        x_1, x_2 = r_a[0] + r_a[2], r_b[0]
    else:
        x_1, x_2 = r_b[0] + r_b[2], r_a[0]
    if r_a[3] + r_b[3] > max(r_a[1] + r_a[3], r_b[1] + r_b[3]) - min(r_a[1], r_b[1]):
        pass
    elif r_a[1] < r_b[1]:
        y_1, y_2 = r_a[1] + r_a[3], r_b[1]
    else:
        y_1, y_2 = r_b[1] + r_b[3], r_a[1]
    return eucl_distance_sq((x_1, y_1), (x_2, y_2)) ** 0.5


def merge_components(rectangles, scale=None):
    """Merge stats rows [x, y, w, h, area] that overlap or lie within MAX_GAP of each other, largest first, until nothing
    merges any more (irtrackextractor.py:324-389; the reference's arithmetic is kept as written, including the bottom edge
    it derives from the x extent)."""
    min_mass, min_size, max_gap = 10 * 4, 16, 40
    if scale:
        min_mass = int(min_mass * scale)
        min_size = int(min_size * scale)
        max_gap *= scale
    rectangles = [r for r in rectangles if r[4] > min_mass or (r[2] > min_size and r[3] > min_size)]
    rectangles = sorted(rectangles, key=lambda s: s[4], reverse=True)
    rectangles = [(r, r.copy()) for r in rectangles]  # merge on the original rectangle, not on the grown one
    rect_i = 0
    while rect_i < len(rectangles):
        rect, merged_r = rectangles[rect_i]
        merged = False
        index = 0
        while index < len(rectangles):
            within = False
            r_2 = rectangles[index][0]
            if r_2[0] == rect[0]:
                index += 1
                continue
            if r_2[2] + rect[2] > max(r_2[0] + r_2[2], rect[2] + rect[0]) - min(r_2[0], rect[0]):
                within = r_2[3] + rect[3] > max(r_2[1] + r_2[3], rect[1] + rect[3]) - min(r_2[1], rect[1])
            if rect_distance(rect, r_2) < max_gap or within:
                cur_right = merged_r[0] + merged_r[2]
                merged_r[0] = min(merged_r[0], r_2[0])
                merged_r[1] = min(merged_r[1], r_2[1])
                merged_r[2] = max(cur_right, r_2[0] + r_2[2])
                merged_r[3] = max(merged_r[1] + merged_r[3], r_2[1] + r_2[3])
                merged_r[2] -= merged_r[0]
                merged_r[3] -= merged_r[1]
                merged_r[4] += r_2[4]
                merged = True
                del rectangles[index]
            else:
                index += 1
        rect_i = 0 if merged else rect_i + 1
    return [rect[1] for rect in rectangles]


def detect_ir_regions(filtered, scale=None):
    """What ``IRTrackExtractor._process_frame`` does with a background-subtracted frame (irtrackextractor.py:453-455):
    ``detect_objects_ir(filtered, threshold=0)`` -> stats rows without the background -> ``merge_components``.
    Returns (labels, merged rectangles)."""
    _, mask, component_details = detect_objects_ir(filtered, threshold=0)
    return mask, merge_components(list(component_details[1:]), scale=scale)
