"""``ml_tools.preprocess`` on the B200 (ml_tools/preprocess.py:19-202 of the reference):
``preprocess_frame`` (crop by region, resize with aspect, median subtraction, normalisation) and
``preprocess_movement`` (25 frames tiled 5x5 per channel).

These are the per-object forms the reference exposes; each call runs the same device kernels as the
batched ``Interpreter.preprocess_segments`` path on a batch of one.  No CPU fallback.
"""
import logging

import numpy as np

from .. import engine as _engine
from .. import native
from . import imageprocessing
from .frame import Frame, TrackChannels

MIN_SIZE = 4
EDGE = 1
res_x = 120
res_y = 160


def preprocess_fn(x):
    """tf.keras inception-style input scaling (preprocess.py:19-22)."""
    x /= 127.5
    x -= 1.0
    return x


def _rect(region):
    return int(region.x), int(region.y), int(region.width), int(region.height)


def preprocess_frame(frame, out_dim, region, background=None, crop_rectangle=None, calculate_filtered=True,
                     filtered_norm_limits=None, thermal_norm_limits=None, cropped=False, sub_median=True, median=None,
                     clip_thermals_at_zero=True):
    """One track-frame -> ``Frame`` with ``out_dim`` thermal / filtered (/ mask) (preprocess.py:56-113)."""
    if cropped:
        raise NotImplementedError("preprocess_frame(cropped=True): pass the full frame and its region")
    if out_dim[0] != out_dim[1]:
        raise NotImplementedError("preprocess_frame: square output only")
    import torch

    thermal = np.asarray(frame.thermal)
    H, W = thermal.shape
    eng = _engine.get_engine(None, W, H, 1)
    ctx = eng.ctx
    ctx.use_torch_stream()
    x, y, w, h = _rect(region)
    if w <= 0 or h <= 0 or x < 0 or y < 0 or x + w > W or y + h > H:
        raise ValueError("region {} is empty or outside the {}x{} frame".format(region, W, H))
    if thermal.dtype != np.uint16 and (not np.array_equal(thermal, np.round(thermal)) or thermal.min() < 0 or thermal.max() > 65535):
        raise NotImplementedError("preprocess_frame: thermal must hold uint16 counts")
    d_t = torch.from_numpy(np.ascontiguousarray(thermal, dtype=np.uint16).view(np.int16)).to(eng.device).view(torch.uint16)
    if calculate_filtered:
        filt = None
        if background is None:
            logging.warning("Not calculating filtered frame as no background was supplied")
        else:
            filt = np.float32(thermal) - np.float32(background)
    else:
        filt = None if frame.filtered is None else np.float32(frame.filtered)

    samples = np.zeros(1, native.SAMPLE_DTYPE)
    samples["x"], samples["y"], samples["width"], samples["height"] = x, y, w, h
    if sub_median and median is None:
        # np.median(frame.thermal) on the device
        d_s = torch.from_numpy(samples.view(np.uint8).copy()).to(eng.device)
        d_tr = torch.empty((1, native.TRACK_NORM_DTYPE.itemsize), dtype=torch.uint8, device=eng.device)
        ctx.preprocess_limits(None, None, 0, d_tr, 1)
        ctx.preprocess_medians(d_t, d_s, 1, d_tr)
        median = float(d_s.cpu().numpy().view(native.SAMPLE_DTYPE)["median"][0])
    samples["median"] = np.float32(median) if sub_median else 0.0

    fused = filtered_norm_limits is not None and thermal_norm_limits is None and filt is not None
    if not fused:
        return _preprocess_frame_general(frame, out_dim, region, filt, crop_rectangle, thermal_norm_limits, filtered_norm_limits,
                                         sub_median, float(samples["median"][0]), clip_thermals_at_zero)
    lo, hi = filtered_norm_limits
    tr = np.zeros(1, native.TRACK_NORM_DTYPE)
    tr["clip_at_zero"] = int(bool(clip_thermals_at_zero))
    tr["has_limits"] = int(lo is not None)
    tr["filtered_min"] = 0.0 if lo is None else np.float32(lo)
    tr["filtered_max"] = np.float32(hi)
    d_f = torch.from_numpy(np.ascontiguousarray(filt)).to(eng.device)
    d_s = torch.from_numpy(samples.view(np.uint8).copy()).to(eng.device)
    d_tr = torch.from_numpy(tr.view(np.uint8).copy()).to(eng.device)
    size = int(out_dim[0])
    d_out = torch.empty((1, size, size, 2), dtype=torch.float32, device=eng.device)
    d_seg = torch.zeros((1, 1), dtype=torch.int32, device=eng.device)
    crop = None if crop_rectangle is None else (crop_rectangle.x, crop_rectangle.y, crop_rectangle.width, crop_rectangle.height)
    ctx.preprocess_segments(d_t, d_f, d_s, d_tr, d_seg, 1, 1, 1, size, crop, 0, d_out)
    tile = d_out.cpu().numpy()[0]
    out = Frame(np.ascontiguousarray(tile[:, :, 0]), np.ascontiguousarray(tile[:, :, 1]), frame.frame_number,
                flow_clipped=frame.flow_clipped, ffc_affected=frame.ffc_affected, region=region)
    if not calculate_filtered and frame.mask is not None:
        out.mask = imageprocessing.resize_and_pad(region.subimage(frame.mask), out_dim, region, crop_rectangle, keep_edge=True,
                                                  pad=0, interpolation=imageprocessing.INTER_NEAREST)
    out.preprocessed = True
    return out


def _preprocess_frame_general(frame, out_dim, region, filt, crop_rectangle, thermal_norm_limits, filtered_norm_limits,
                              sub_median, median, clip_thermals_at_zero):
    """The less common option combinations, composed from the single-image device helpers."""
    cropped = Frame(np.float32(region.subimage(frame.thermal)), None if filt is None else region.subimage(filt),
                    frame.frame_number, mask=None if frame.mask is None else region.subimage(frame.mask),
                    flow_clipped=frame.flow_clipped, ffc_affected=frame.ffc_affected, region=region)
    cropped.resize_with_aspect(out_dim, crop_rectangle, True)
    if sub_median:
        cropped.thermal -= np.float32(median)
    if thermal_norm_limits is None and clip_thermals_at_zero:
        np.clip(cropped.thermal, 0, None, out=cropped.thermal)
    if filtered_norm_limits is not None:
        if cropped.filtered is not None:
            cropped.filtered, _ = imageprocessing.normalize(cropped.filtered, min=filtered_norm_limits[0], max=filtered_norm_limits[1],
                                                            new_max=255)
        t_min, t_max = (None, None) if thermal_norm_limits is None else thermal_norm_limits
        cropped.thermal, _ = imageprocessing.normalize(cropped.thermal, min=t_min, max=t_max, new_max=255)
    else:
        cropped.normalize()
    cropped.preprocessed = True
    return cropped


def preprocess_single_frame(preprocessed_frame, channels, preprocess_fn=None, save_info=""):
    data = []
    for channel in channels:
        if isinstance(channel, str):
            channel = TrackChannels[channel]
        data.append(preprocessed_frame.get_channel(channel))
    image = np.stack(data, axis=2)
    if preprocess_fn:
        image = preprocess_fn(image)
    return image


def preprocess_movement(preprocess_frames, frames_per_row, frame_size, channels, preprocess_fn=None, sample=None, seed=None):
    """Tile already preprocessed frames 5x5 per channel (preprocess.py:151-202).  A gather of host arrays the caller
    already holds, so it stays a host copy; the batched path writes tiles into this layout from the kernel."""
    from ..batch import pad_segment_samples

    if len(preprocess_frames) == 0:
        return None
    frame_samples = pad_segment_samples(len(preprocess_frames), frames_per_row * 5, seed)
    frame_types = {}
    data = []
    for channel in channels:
        if isinstance(channel, str):
            channel = TrackChannels[channel]
        if channel in frame_types:
            data.append(frame_types[channel])
            continue
        channel_segment = [frame.get_channel(channel) for frame in preprocess_frames]
        channel_data, success = imageprocessing.square_clip(channel_segment, frames_per_row, (frame_size, frame_size),
                                                            frame_samples, normalize=False)
        if not success:
            return None
        data.append(channel_data)
        frame_types[channel] = channel_data
    data = np.stack(data, axis=2)
    if preprocess_fn:
        data = preprocess_fn(data)
    return np.float32(data)
