"""Per-process registry of device engines: one ``BatchExtractor`` (native context, stream, weight
tables) per (device, geometry).  Everything in the host-side mirror of the reference API that needs
the GPU goes through ``get_engine``; there is no CPU fallback behind it."""
import threading

from .batch import BatchExtractor

_lock = threading.Lock()
_engines = {}
DEFAULT_DEVICE = 0
# region slots per frame of the API engines: every component the uint8 label image can number (CPT_MAX_COMPONENTS).  The
# reference iterates every stats row (cliptracker.py:263-365); a noisy frame (FFC, lens cap) must not abort a batch.
API_MAX_REGIONS = 255


def set_default_device(device):
    global DEFAULT_DEVICE
    DEFAULT_DEVICE = int(device)


def get_engine(device=None, width=160, height=120, edge_pixels=1, max_regions=API_MAX_REGIONS):
    device = DEFAULT_DEVICE if device is None else int(device)
    key = (device, width, height, edge_pixels, max_regions)
    with _lock:
        eng = _engines.get(key)
        if eng is None:
            eng = BatchExtractor(device=device, width=width, height=height, edge_pixels=edge_pixels, max_regions=max_regions)
            _engines[key] = eng
        return eng
