"""Small host helpers the tracking code shares (ml_tools/tools.py:42-61,180-186 of the reference)."""
import datetime
import json
from enum import Enum
from pathlib import Path

import numpy as np

from .rectangle import Rectangle


def eucl_distance_sq(first, second):
    dx = first[0] - second[0]
    dy = first[1] - second[1]
    return dx * dx + dy * dy


class CustomJSONEncoder(json.JSONEncoder):
    """numpy scalars / arrays, datetimes, paths, enums and Rectangles -> JSON."""

    def default(self, obj):
        if isinstance(obj, np.integer):
            return int(obj)
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.bool_):
            return bool(obj)
        if isinstance(obj, np.ndarray):
            return list(obj)
        if isinstance(obj, datetime.datetime):
            return obj.isoformat()
        if isinstance(obj, Rectangle):
            return obj.meta_dictionary()
        if isinstance(obj, Path):
            return str(obj)
        if isinstance(obj, Enum):
            return str(obj.name)
        return super().default(obj)


def load_clip_metadata(filename):
    """Loads the metadata file of a clip (tools.py:90-104)."""
    with open(filename, "r") as t:
        meta = json.load(t)
    if meta.get("recordingDateTime"):
        from datetime import datetime

        meta["recordingDateTime"] = datetime.fromisoformat(meta["recordingDateTime"].replace("Z", "+00:00"))
    if meta.get("tracks") is None and meta.get("Tracks"):
        meta["tracks"] = meta["Tracks"]
    return meta
