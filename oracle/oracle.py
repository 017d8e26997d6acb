"""ctypes front-end of the C oracle (``oracle/cptrack_oracle.c``).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs import this
module; the product package never does.
"""
import ctypes
import os

import numpy as np

from . import build as _build

_lib = None


class OrcParams(ctypes.Structure):
    _fields_ = [
        ("W", ctypes.c_int32),
        ("H", ctypes.c_int32),
        ("edge", ctypes.c_int32),
        ("background_thresh", ctypes.c_int32),
        ("weight_add", ctypes.c_double),
        ("denoise", ctypes.c_int32),
        ("update_background", ctypes.c_int32),
        ("max_comp", ctypes.c_int32),
        ("calc_stats", ctypes.c_int32),
    ]


def lib():
    global _lib
    if _lib is None:
        path = _build.LIB
        if not os.path.exists(path) or any(
            os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(path) for s in _build.SOURCES
        ):
            path = _build.build()
        _lib = ctypes.CDLL(path)
        _lib.orc_box_variance.restype = ctypes.c_double
        _lib.orc_background_create.restype = ctypes.c_void_p
    return _lib


def _p(a, ctype=None):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.c_void_p)


def normalise_frame(pix, bg, bg_average, background_thresh):
    """K2.  Returns (u8 image, threshold fp32, max, min, avg_change)."""
    pix = np.ascontiguousarray(pix, dtype=np.uint16)
    bg = np.ascontiguousarray(bg, dtype=np.int32)
    u = np.empty(pix.shape, np.uint8)
    th, mx, mn = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
    ac = ctypes.c_int32()
    lib().orc_normalise_frame(
        _p(pix), _p(bg), ctypes.c_double(bg_average), ctypes.c_int(pix.size), ctypes.c_int(int(background_thresh)),
        _p(u), ctypes.byref(th), ctypes.byref(mx), ctypes.byref(mn), ctypes.byref(ac),
    )
    return u, np.float32(th.value), np.float32(mx.value), np.float32(mn.value), ac.value


def blur5(u):
    u = np.ascontiguousarray(u, dtype=np.uint8)
    out = np.empty_like(u)
    lib().orc_blur5(_p(u), ctypes.c_int(u.shape[1]), ctypes.c_int(u.shape[0]), _p(out))
    return out


def threshold_close(blurred, thresh):
    blurred = np.ascontiguousarray(blurred, dtype=np.uint8)
    out = np.empty_like(blurred)
    lib().orc_threshold_close(
        _p(blurred), ctypes.c_int(blurred.shape[1]), ctypes.c_int(blurred.shape[0]), ctypes.c_float(float(thresh)), _p(out)
    )
    return out


def cc8(mask, max_comp=4096):
    """K5.  Returns (n_components, labels int32, comp int32 (n,8): l,t,w,h,area,sumx,sumy,key)."""
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    labels = np.empty(mask.shape, np.int32)
    comp = np.zeros((max_comp, 8), np.int32)
    n = lib().orc_cc8(_p(mask), ctypes.c_int(mask.shape[1]), ctypes.c_int(mask.shape[0]), _p(labels), _p(comp), ctypes.c_int(max_comp))
    return n, labels, comp[: min(n, max_comp)].copy()


def detect_objects(u8, thresh, max_comp=4096):
    """blur -> threshold -> close -> CC, OpenCV-compatible tuple (n+1, labels, stats, centroids)."""
    mask = threshold_close(blur5(u8), thresh)
    n, labels, comp = cc8(mask, max_comp)
    return n, labels, comp


def stats_centroids_from_comp(comp):
    """cv2-style stats rows (no background row) and double centroids."""
    stats = comp[:, :5].astype(np.int32)
    cents = np.stack([comp[:, 5] / comp[:, 4].astype(np.float64), comp[:, 6] / comp[:, 4].astype(np.float64)], axis=1)
    return stats, cents


def nlm_denoise(u8):
    u8 = np.ascontiguousarray(u8, dtype=np.uint8)
    out = np.empty_like(u8)
    lib().orc_nlm_denoise(_p(u8), ctypes.c_int(u8.shape[1]), ctypes.c_int(u8.shape[0]), _p(out))
    return out


class Background:
    """K7 WeightedBackground restatement (stateful)."""

    def __init__(self, W, H, edge, weight_add):
        self.W, self.H, self.edge = W, H, edge
        self._h = ctypes.c_void_p(lib().orc_background_create(W, H, edge, ctypes.c_double(weight_add)))

    def process(self, frame_i32):
        f = np.ascontiguousarray(frame_i32, dtype=np.int32)
        lib().orc_background_process(self._h, _p(f))

    def get(self):
        bg = np.empty((self.H, self.W), np.int32)
        w = np.empty((self.H - 2 * self.edge, self.W - 2 * self.edge), np.float64)
        avg = ctypes.c_double()
        lib().orc_background_get(self._h, _p(bg), _p(w), ctypes.byref(avg))
        return bg, w, avg.value

    def __del__(self):
        try:
            lib().orc_background_destroy(self._h)
        except Exception:
            pass


def frame_stats(pix, filtered=None):
    pix = np.ascontiguousarray(pix, dtype=np.uint16)
    out = np.zeros(5, np.float64)
    f = None if filtered is None else np.ascontiguousarray(filtered, dtype=np.float32)
    lib().orc_frame_stats(_p(pix), _p(f), ctypes.c_int(pix.size), _p(out))
    return out


def make_params(W=160, H=120, edge=1, background_thresh=20, weight_add=0.1, denoise=False, update_background=True,
                max_comp=64, calc_stats=False):
    return OrcParams(W, H, edge, int(background_thresh), float(weight_add), int(denoise), int(update_background),
                     int(max_comp), int(calc_stats))


def extract_clip(frames, init_frame=None, params=None, want=("filtered", "labels", "u", "bg", "fstats")):
    """Run the whole-clip restatement.  Returns a dict of numpy arrays."""
    frames = np.ascontiguousarray(frames, dtype=np.uint16)
    T, H, W = frames.shape
    p = params or make_params(W=W, H=H)
    init = frames[0] if init_frame is None else np.ascontiguousarray(init_frame, dtype=np.uint16)
    mc = p.max_comp
    out = dict(
        ncomp=np.zeros(T, np.int32),
        comp=np.zeros((T, mc, 8), np.int32),
        var=np.zeros((T, mc), np.float64),
        thresh=np.zeros(T, np.float32),
        norm=np.zeros((T, 2), np.float32),
        avg=np.zeros(T, np.float64),
        final_bg=np.zeros((H, W), np.int32),
        final_weight=np.zeros((H - 2 * p.edge, W - 2 * p.edge), np.float64),
    )
    opt = dict(
        filtered=np.zeros((T, H, W), np.float32) if "filtered" in want else None,
        labels=np.zeros((T, H, W), np.uint8) if "labels" in want else None,
        u=np.zeros((T, H, W), np.uint8) if "u" in want else None,
        bg=np.zeros((T, H, W), np.int32) if "bg" in want else None,
        fstats=np.zeros((T, 5), np.float64) if ("fstats" in want and p.calc_stats) else None,
    )
    favg = ctypes.c_double()
    lib().orc_extract_clip(
        _p(frames), ctypes.c_int(T), _p(init), ctypes.byref(p), _p(opt["filtered"]), _p(opt["labels"]),
        _p(out["ncomp"]), _p(out["comp"]), _p(out["var"]), _p(opt["u"]), _p(out["thresh"]), _p(out["norm"]),
        _p(opt["bg"]), _p(out["avg"]), _p(opt["fstats"]), _p(out["final_bg"]), _p(out["final_weight"]),
        ctypes.byref(favg),
    )
    out["final_avg"] = favg.value
    out.update({k: v for k, v in opt.items() if v is not None})
    return out


def extract_batch(frames, params_list, n_threads):
    """CPU-baseline batch: frames (clips, T, H, W) uint16 -> ncomp, comp, var (regions only)."""
    frames = np.ascontiguousarray(frames, dtype=np.uint16)
    C, T, H, W = frames.shape
    arr = (OrcParams * C)(*params_list)
    mc = params_list[0].max_comp
    ncomp = np.zeros((C, T), np.int32)
    comp = np.zeros((C, T, mc, 8), np.int32)
    var = np.zeros((C, T, mc), np.float64)
    lib().orc_extract_batch(_p(frames), ctypes.c_int(C), ctypes.c_int(T), arr, _p(ncomp), _p(comp), _p(var), ctypes.c_int(n_threads))
    return ncomp, comp, var
