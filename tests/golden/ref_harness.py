"""Import the Python reference from /root/reference/src (THIS container only).

Used by ``make_golden.py`` to generate the committed fixtures and by a few ``not gpu``
tests that re-validate the oracle against the live reference when it is present.  The
GPU box has no /root/reference: nothing under ``-m gpu``, ``smoke()`` or ``bench.py``
imports this file.

Recipe: stub the modules the reference imports but this image
lacks (pytz, h5py, timezonefinder, matplotlib, and the Rust CPTV reader, which is
replaced by this repo's decoder plus an in-memory registry for synthetic clips).
"""
import os
import sys
import types
import zoneinfo

REFERENCE_SRC = "/root/reference/src"
REPO_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

_MEMORY_CLIPS = {}


def available():
    return os.path.isdir(REFERENCE_SRC)


class _MemHeader:
    def __init__(self, width, height, model):
        self.x_resolution = width
        self.y_resolution = height
        self.model = model
        self.brand = "flir"
        self.timestamp = 1_600_000_000_000_000
        self.fps = 9


class _MemFrame:
    def __init__(self, pix, t):
        self.pix = pix
        # far from the last FFC so that is_affected_by_ffc is False
        self.time_on = 10_000_000 + t * 111
        self.last_ffc_time = 0
        self.temp_c = 20.0
        self.last_ffc_temp_c = 20.0
        self.background_frame = False


class _MemReader:
    def __init__(self, key):
        self.pix, self.model = _MEMORY_CLIPS[key]
        self.i = 0

    def get_header(self):
        return _MemHeader(self.pix.shape[2], self.pix.shape[1], self.model)

    def next_frame(self):
        if self.i >= len(self.pix):
            return None
        f = _MemFrame(self.pix[self.i], self.i)
        self.i += 1
        return f


def register_memory_clip(key, pix, model):
    _MEMORY_CLIPS[key] = (pix, model)


def _reader_factory(path):
    if path in _MEMORY_CLIPS:
        return _MemReader(path)
    if REPO_ROOT not in sys.path:
        sys.path.insert(0, REPO_ROOT)
    from classifier_pipeline_b200.cptv import CptvReader

    return CptvReader(path)


def setup():
    """Install stubs and put the reference on sys.path.  Idempotent."""
    if not available():
        raise RuntimeError("reference tree not present")
    if "cptv_rs_python_bindings" in sys.modules and getattr(sys.modules["cptv_rs_python_bindings"], "_graft_stub", False):
        return
    pytz = types.ModuleType("pytz")
    pytz.timezone = zoneinfo.ZoneInfo
    pytz.utc = zoneinfo.ZoneInfo("UTC")
    sys.modules.setdefault("pytz", pytz)
    for name in ("h5py", "timezonefinder"):
        sys.modules.setdefault(name, types.ModuleType(name))
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)
    rs = types.ModuleType("cptv_rs_python_bindings")
    rs.CptvReader = _reader_factory
    rs._graft_stub = True
    sys.modules["cptv_rs_python_bindings"] = rs
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
