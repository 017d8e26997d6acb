"""``ml_tools.imageprocessing`` on the B200 (ml_tools/imageprocessing.py:11-248 of the reference).

Same function names, arguments and return values as the reference module; the arithmetic runs in
``libcptrack.so`` (csrc/image_kernels.cu, csrc/detect_kernels.cu).  numpy arrays go in and come out
(the callers mutate the returned arrays in place); there is no CPU fallback.  The batched paths
(``ClipTrackExtractor``, ``Interpreter.preprocess_segments``) do not go through these per-image
helpers: they fuse the same arithmetic inside their own kernels.
"""
import numpy as np

from .. import engine as _engine

INTER_NEAREST = 0  # cv2.INTER_NEAREST
INTER_LINEAR = 1   # cv2.INTER_LINEAR


def _torch():
    import torch

    return torch


def _to_device_f32(eng, data):
    torch = _torch()
    host = np.ascontiguousarray(data, dtype=np.float32)  # np.float32(data): the reference's first step
    return torch.from_numpy(host).to(eng.device)


def resize_cv(image, dim, interpolation=INTER_LINEAR, extra_h=0, extra_v=0):
    """``cv2.resize(np.float32(image), dsize=(dim[0] + extra_h, dim[1] + extra_v))`` (imageprocessing.py:77-82)."""
    image = np.asarray(image)
    if image.ndim != 2:
        raise NotImplementedError("resize_cv: single-channel images only")
    torch = _torch()
    eng = _engine.get_engine()
    eng.ctx.use_torch_stream()
    dw, dh = int(dim[0] + extra_h), int(dim[1] + extra_v)
    sh, sw = image.shape
    d_src = _to_device_f32(eng, image)
    d_out = torch.empty((dh, dw), dtype=torch.float32, device=eng.device)
    eng.ctx.resize_pad_f32(d_src, sw, sh, dw, dh, 0, 0, dw, dh, 0.0, interpolation, d_out)
    return d_out.cpu().numpy()


def resize_and_pad(frame, new_dim, region, crop_region, keep_edge=False, pad=None, interpolation=INTER_LINEAR,
                   extra_h=0, extra_v=0, edge_offset=(0, 0, 0, 0), original_region=None):
    """Aspect-preserving resize into ``new_dim`` with padding (imageprocessing.py:11-70)."""
    frame = np.asarray(frame)
    if frame.ndim != 2:
        raise NotImplementedError("resize_and_pad: single-channel images only")
    scale_percent = (np.array(new_dim[:2]) / np.array(frame.shape[:2])).min()
    width = min(max(round(frame.shape[1] * scale_percent), 1), new_dim[0])
    height = min(max(round(frame.shape[0] * scale_percent), 1), new_dim[1])
    if pad is None:
        pad = np.min(frame)
    if original_region is None:
        original_region = region
    offset_x = (new_dim[1] - width) // 2
    offset_y = (new_dim[0] - height) // 2
    if keep_edge and crop_region is not None:
        if original_region.left <= crop_region.left:
            offset_x = min(edge_offset[0], new_dim[1] - width)
        elif original_region.right >= crop_region.right:
            offset_x = max((new_dim[1] - edge_offset[2]) - width, 0)
        if original_region.top <= crop_region.top:
            offset_y = min(edge_offset[1], new_dim[0] - height)
        elif original_region.bottom >= crop_region.bottom:
            offset_y = max(new_dim[0] - height - edge_offset[3], 0)
    if offset_x < 0 or offset_y < 0:
        raise ValueError("resize_and_pad: negative paste offset (edge_offset larger than the free space)")
    torch = _torch()
    eng = _engine.get_engine()
    eng.ctx.use_torch_stream()
    sh, sw = frame.shape
    d_src = _to_device_f32(eng, frame)
    d_out = torch.empty((int(new_dim[0]), int(new_dim[1])), dtype=torch.float32, device=eng.device)
    eng.ctx.resize_pad_f32(d_src, sw, sh, int(width), int(height), int(offset_x), int(offset_y), int(new_dim[1]), int(new_dim[0]),
                           float(pad), interpolation, d_out)
    out = d_out.cpu().numpy()
    return out if frame.dtype == np.float32 else out.astype(frame.dtype)


def square_clip(data, frames_per_row, tile_dim, frame_samples, normalize=True):
    """Lay frames out side by side in rows (imageprocessing.py:85-104).  A pure gather: the batched
    preprocessing kernel writes tiles straight into this layout; this single-segment form copies on the host."""
    new_frame = np.zeros((frames_per_row * tile_dim[0], frames_per_row * tile_dim[1]))
    i = 0
    success = False
    for x in range(frames_per_row):
        for y in range(frames_per_row):
            frame = data[frame_samples[i]]
            if normalize:
                frame, stats = globals()["normalize"](frame, new_max=255)
                if not stats[0]:
                    continue
            success = True
            new_frame[x * tile_dim[0] : (x + 1) * tile_dim[0], y * tile_dim[1] : (y + 1) * tile_dim[1]] = np.float32(frame)
            i += 1
    return new_frame, success


def _is_weak(v):
    return isinstance(v, (int, float)) and not isinstance(v, np.generic)


def normalize(data, min=None, max=None, new_max=1):
    """Normalize an array so that the values range from 0 -> new_max (imageprocessing.py:151-169).
    Returns normalized array, stats tuple (Success, max used, min used)."""
    data = np.asarray(data)
    if data.size == 0:
        return np.zeros(data.shape), (False, None, None)
    torch = _torch()
    eng = _engine.get_engine()
    eng.ctx.use_torch_stream()
    exact_in_f32 = data.dtype in (np.float32, np.uint8, np.int8, np.uint16, np.int16, np.bool_)
    d_in = _to_device_f32(eng, data)
    if max is None or min is None:
        if exact_in_f32:
            d_mm = torch.empty((2,), dtype=torch.float32, device=eng.device)
            eng.ctx.minmax_f32(d_in, data.size, d_mm)
            mn_mx = d_mm.cpu().numpy()
            found_min, found_max = data.dtype.type(mn_mx[0]), data.dtype.type(mn_mx[1])
        else:  # wider types: the extrema themselves must not be rounded
            found_min, found_max = np.amin(data), np.amax(data)
        if max is None:
            max = found_max
        if min is None:
            min = found_min
    if max == min:
        if max == 0:
            return np.zeros(data.shape), (False, max, min)
        # data / max: true division -- fp32 only when both sides are fp32 (or max is a weak python scalar)
        use_f64 = not (data.dtype == np.float32 and (_is_weak(max) or np.asarray(max).dtype == np.float32))
    else:
        # numpy promotion of new_max * (float32(data) - min) / (max - min): python scalars are weak
        types = [np.float32] + [np.asarray(v).dtype for v in (min, max, new_max) if not _is_weak(v)]
        out_dtype = np.result_type(*types)
        if out_dtype not in (np.float32, np.float64):
            out_dtype = np.dtype(np.float64)
        use_f64 = out_dtype == np.float64
    d_out = torch.empty(data.shape, dtype=torch.float64 if use_f64 else torch.float32, device=eng.device)
    eng.ctx.normalize_f32(d_in, data.size, float(min), float(max), float(new_max), use_f64, d_out)
    return d_out.cpu().numpy(), (True, max, min)


def detect_objects(image, otsus=False, threshold=30, kernel=(15, 15)):
    """uint8 -> GaussianBlur -> threshold -> morphologyEx CLOSE -> connectedComponentsWithStats
    (imageprocessing.py:240-248).  Returns (n, labels int32, stats int32 (n,5), centroids float64 (n,2))."""
    from . import detect

    return detect.detect_objects(image, otsus=otsus, threshold=threshold, kernel=kernel)


def detect_objects_ir(image, otsus=False, threshold=100, kernel=(15, 15)):
    """morphologyEx OPEN -> threshold -> connectedComponentsWithStats (imageprocessing.py:185-199)."""
    from . import detect

    return detect.detect_objects_ir(image, otsus=otsus, threshold=threshold, kernel=kernel)


def detect_objects_both(salicencyMap, backsub, threshold=30, kernel=(15, 15), otsus=False):
    """imageprocessing.py:202-238."""
    from . import detect

    return detect.detect_objects_both(salicencyMap, backsub, threshold=threshold, kernel=kernel, otsus=otsus)


def fast_nl_means_denoising(image):
    """``cv2.fastNlMeansDenoising(np.uint8(image), None)`` as ``ClipTracker._get_filtered_frame`` calls it
    (track/cliptracker.py:116-117): h = 3, 7x7 template, 21x21 search window; uint8 (H, W) or (N, H, W)."""
    image = np.ascontiguousarray(np.uint8(image))
    if image.ndim not in (2, 3):
        raise ValueError("fast_nl_means_denoising: (H, W) or (N, H, W) uint8 expected")
    torch = _torch()
    eng = _engine.get_engine()
    eng.ctx.use_torch_stream()
    n = 1 if image.ndim == 2 else image.shape[0]
    H, W = image.shape[-2:]
    d_src = torch.from_numpy(image).to(eng.device)
    d_dst = torch.empty_like(d_src)
    eng.ctx.nlm_denoise_u8(d_src, W, H, n, d_dst)
    return d_dst.cpu().numpy()


def clear_frame(frame):
    filtered = frame.filtered
    thermal = frame.thermal
    if len(filtered) == 0 or len(thermal) == 0:
        return False
    return bool(np.amax(thermal) != np.amin(thermal) and np.amax(filtered) != np.amin(filtered))
