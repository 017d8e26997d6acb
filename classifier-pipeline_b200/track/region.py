"""``Region``: a ``Rectangle`` with mass, centroid and per-frame bookkeeping (track/region.py:27-228)."""
import numpy as np

from ..ml_tools.rectangle import Rectangle
from ..ml_tools.tools import eucl_distance_sq


class Region(Rectangle):
    __slots__ = ("centroid", "mass", "frame_number", "pixel_variance", "id", "was_cropped", "blank",
                 "is_along_border", "in_trap")

    def __init__(self, x, y, width, height, centroid=None, mass=0, frame_number=0, pixel_variance=0, id=0,
                 was_cropped=False, blank=False, is_along_border=False, in_trap=False):
        super().__init__(x, y, width, height)
        self.centroid = centroid
        self.mass = mass
        self.frame_number = frame_number
        self.pixel_variance = pixel_variance
        self.id = id
        self.was_cropped = was_cropped
        self.blank = blank
        self.is_along_border = is_along_border
        self.in_trap = in_trap

    # ------------------------------------------------------------------ constructors
    @staticmethod
    def from_ltwh(left, top, width, height):
        return Region(left, top, width, height, centroid=None)

    @staticmethod
    def from_ltrb(left, top, right, bottom):
        return Region(left, top, right - left, bottom - top)

    @classmethod
    def region_from_array(cls, bounds):
        """[left, top, right, bottom, frame_number?, mass?, blank?] (region.py:70-99)."""
        width = np.uint8(max(int(bounds[2]) - bounds[0], 0))
        height = np.uint8(max(int(bounds[3]) - bounds[1], 0))
        frame_number = np.uint16(bounds[4]) if len(bounds) > 4 and bounds[4] is not None else None
        mass = bounds[5] if len(bounds) > 5 else 0
        blank = bounds[6] == 1 if len(bounds) > 6 else False
        centroid = [int(bounds[0] + width / 2), int(bounds[1] + height / 2)]
        return cls(bounds[0], bounds[1], width, height, frame_number=frame_number, mass=mass, blank=blank, centroid=centroid)

    @classmethod
    def region_from_json(cls, data):
        """Position dictionaries of the tracks JSON (region.py:101-127)."""
        frame = data.get("frame_number")
        if frame is None:
            frame = data.get("frameNumber")
        if frame is None:
            frame = data.get("order")
        centroid = data.get("centroid")
        if centroid is None:
            centroid = [int(data["x"] + data["width"] / 2), int(data["y"] + data["height"] / 2)]
        mass = data.get("mass", 0)
        return cls(data["x"], data["y"], data["width"], data["height"], frame_number=frame,
                   mass=0 if mass is None else mass, blank=data.get("blank", False),
                   pixel_variance=data.get("pixel_variance", 0), centroid=centroid)

    def copy(self):
        # in_trap is not carried over (region.py:163-177)
        return Region(self.x, self.y, self.width, self.height, self.centroid, self.mass, self.frame_number,
                      self.pixel_variance, self.id, self.was_cropped, self.blank, self.is_along_border)

    def to_array(self):
        return np.uint16([self.left, self.top, self.right, self.bottom, self.frame_number, self.mass, 1 if self.blank else 0])

    # ------------------------------------------------------------------ behaviour
    def rescale(self, factor):
        self.x = int(self.x * factor)
        self.y = int(self.y * factor)
        self.width = int(self.width * factor)
        self.height = int(self.height * factor)
        self.mass = self.mass * (factor**2)

    def has_moved(self, other):
        """Shifted (both edges of an axis moved), not merely grown (region.py:134-140)."""
        return (self.x != other.x and self.right != other.right) or (self.y != other.y and self.bottom != other.bottom)

    def set_is_along_border(self, bounds, edge=0):
        # compares against bounds.width / bounds.height, not right / bottom: region.py:154-161
        self.is_along_border = (
            self.was_cropped
            or self.x <= bounds.x + edge
            or self.y <= bounds.y + edge
            or self.right >= bounds.width - edge
            or self.bottom >= bounds.height - edge
        )

    def average_distance(self, other):
        """Squared distances between top-left corners, centres and bottom-right corners (region.py:179-212)."""
        return [
            eucl_distance_sq((int(other.x), int(other.y)), (self.x, self.y)),
            eucl_distance_sq((int(other.mid_x), int(other.mid_y)), (self.mid_x, self.mid_y)),
            eucl_distance_sq((other.right, other.bottom), (self.right, self.bottom)),
        ]

    def on_height_edge(self, crop_region):
        return self.top == crop_region.top or self.bottom == crop_region.bottom

    def on_width_edge(self, crop_region):
        return self.left == crop_region.left or self.right == crop_region.right

    def meta_dictionary(self):
        """Keys and order of the reference's JSON positions (rectangle.py:164-177 applied to Region)."""
        var = self.pixel_variance
        return {
            "x": self.x, "y": self.y, "width": self.width, "height": self.height, "mass": self.mass,
            "frame_number": self.frame_number, "pixel_variance": round(var, 2) if var is not None else 0,
            "blank": self.blank, "in_trap": self.in_trap,
        }
