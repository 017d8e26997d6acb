"""``Rectangle``: top-left + width/height box with the reference's mutation semantics.

Mirrors ml_tools/rectangle.py:6-177 of the reference (same attribute and method names).  The
reference builds the class with ``attrs``; this is a plain ``__slots__`` class.  Equality and
hashing are by identity (``attr.s(eq=False)`` there), which the tracker relies on when it
keeps regions in sets.
"""
import math


class Rectangle:
    __slots__ = ("x", "y", "width", "height")

    def __init__(self, x, y, width, height):
        self.x = x
        self.y = y
        self.width = width
        self.height = height

    # ------------------------------------------------------------------ constructors / views
    @staticmethod
    def from_ltrb(left, top, right, bottom):
        return Rectangle(left, top, right - left, bottom - top)

    def to_ltrb(self):
        return [self.x, self.y, self.x + self.width, self.y + self.height]

    def to_ltwh(self):
        return [self.x, self.y, self.width, self.height]

    def copy(self):
        return Rectangle(self.x, self.y, self.width, self.height)

    # ------------------------------------------------------------------ derived values
    @property
    def left(self):
        return self.x

    @property
    def top(self):
        return self.y

    @property
    def right(self):
        return self.x + self.width

    @property
    def bottom(self):
        return self.y + self.height

    # moving the left/top edge keeps the opposite edge where it is (rectangle.py:62-80)
    @left.setter
    def left(self, value):
        self.width = self.x + self.width - value
        self.x = value

    @top.setter
    def top(self, value):
        self.height = self.y + self.height - value
        self.y = value

    @right.setter
    def right(self, value):
        self.width = value - self.x

    @bottom.setter
    def bottom(self, value):
        self.height = value - self.y

    @property
    def mid_x(self):
        return self.x + self.width / 2

    @property
    def mid_y(self):
        return self.y + self.height / 2

    @property
    def mid(self):
        return (self.mid_x, self.mid_y)

    @property
    def elongation(self):
        return max(self.width, self.height) / min(self.width, self.height)

    @property
    def area(self):
        return int(self.width) * self.height

    # ------------------------------------------------------------------ geometry
    def overlap_area(self, other):
        dx = min(self.right, other.right) - max(self.left, other.left)
        dy = min(self.bottom, other.bottom) - max(self.top, other.top)
        return max(0, dx) * max(0, dy)

    def crop(self, bounds):
        """Clamp every edge into ``bounds`` (rectangle.py:91-96)."""
        left = min(bounds.right, max(self.left, bounds.left))
        top = min(bounds.bottom, max(self.top, bounds.top))
        right = max(bounds.left, min(self.right, bounds.right))
        bottom = max(bounds.top, min(self.bottom, bounds.bottom))
        self.x, self.y = left, top
        self.width, self.height = right - left, bottom - top

    def subimage(self, image):
        return image[self.y : self.y + self.height, self.x : self.x + self.width]

    def enlarge(self, border, max=None):
        """Grow by ``border`` on every side, then clamp into ``max`` (rectangle.py:138-146)."""
        self.x -= border
        self.y -= border
        self.width += 2 * border
        self.height += 2 * border
        if max:
            self.crop(max)

    def enlarge_even(self, width_enlarge, height_enlarge, crop):
        """Grow, then shrink symmetrically by the larger overshoot per axis (rectangle.py:105-136)."""
        self.x -= width_enlarge
        self.width += 2 * width_enlarge
        self.y -= height_enlarge
        self.height += 2 * height_enlarge

        def overshoot(amount, limit):
            return min(max(0, amount), limit)

        dw = max(overshoot(crop.left - self.left, crop.width), overshoot(self.right - crop.right, crop.width))
        self.x += dw
        self.width -= 2 * dw
        dh = max(overshoot(self.bottom - crop.bottom, crop.height), overshoot(crop.top - self.top, crop.height))
        self.y += dh
        self.height -= 2 * dh

    def enlarge_for_rotation(self, crop_rectangle, final_dim=32, extra_needed=13):
        """rectangle.py:182-199: pad so that a rotation augment of the resized tile has no empty corners."""
        scale = min(final_dim / self.width, final_dim / self.height)
        extra = extra_needed / scale
        grow_w = grow_h = math.ceil(extra / 2)
        if self.width > self.height:
            grow_h = math.ceil((extra + (self.width - self.height)) / 2)
        else:
            grow_w = math.ceil((extra + (self.height - self.width)) / 2)
        self.enlarge_even(grow_w, grow_h, crop=crop_rectangle)

    def contains(self, x, y):
        # the reference's (inverted) vertical test is kept: rectangle.py:148-150
        return self.left <= x and self.right >= x and self.top >= y and self.bottom <= y

    def __repr__(self):
        return "(x{0},y{1},x2{2},y2{3})".format(self.left, self.top, self.right, self.bottom)

    def __str__(self):
        return "<(x{0},y{1})-h{2}xw{3}>".format(self.x, self.y, self.height, self.width)

    def meta_dictionary(self):
        """JSON form (rectangle.py:164-177); ``Region`` adds its own fields."""
        return {"x": self.x, "y": self.y, "width": self.width, "height": self.height}
