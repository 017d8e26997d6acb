"""Batched clip extraction driver: many independent clips per launch, one persistent CTA per
clip (``csrc/extract_kernel.cu``).  This is the engine behind ``ClipTrackExtractor.parse_clip``
and the clip-sharded multi-GPU path; it only moves buffers and builds descriptors.
"""
import numpy as np

from . import native


def linear_clips(n_frames, background_thresh, weight_table, flags=native.CLIP_UPDATE_BACKGROUND, init_offsets=None):
    """Descriptors for clips packed back to back: clip i occupies frames [sum(n[:i]), sum(n[:i+1])).

    ``init_offsets`` (absolute frame indices) default to each clip's own first frame
    (track/cliptrackextractor.py:129-139 initialises the background from the first frame read).
    Scalars broadcast over clips.
    """
    n_frames = np.asarray(n_frames, dtype=np.int64).reshape(-1)
    n = len(n_frames)
    clips = np.zeros(n, dtype=native.CLIP_DTYPE)
    starts = np.concatenate([[0], np.cumsum(n_frames)[:-1]]) if n else np.zeros(0, np.int64)
    clips["frame_offset"] = starts
    clips["init_offset"] = starts if init_offsets is None else np.asarray(init_offsets, dtype=np.int64)
    clips["out_offset"] = starts
    clips["n_frames"] = n_frames
    clips["background_thresh"] = background_thresh
    clips["weight_table"] = weight_table
    clips["flags"] = flags
    return clips


class BatchExtractor:
    """Owns a native context and the output buffers of one device."""

    def __init__(self, device=0, width=160, height=120, edge_pixels=1, max_regions=16):
        import torch

        if not torch.cuda.is_available():
            raise native.NativeError("CUDA device required: the extraction path has no CPU fallback")
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.ctx = native.Context(device, width, height, edge_pixels, max_regions)
        self.width, self.height, self.max_regions = width, height, max_regions

    # ------------------------------------------------------------------ device-resident path
    def extract_device(self, d_frames, clips, keep_filtered=True, keep_labels=True, keep_state=False, d_state=None,
                       out=None, use_torch_stream=True):
        """Run the kernel on frames already in HBM.

        d_frames: torch.uint16 CUDA tensor, any shape whose trailing dims are (H, W).
        clips: numpy CLIP_DTYPE array (host) or a CUDA uint8 tensor holding the same records.
        Returns dict(regions, info, filtered, labels, state) of CUDA tensors (raw bytes for the
        two record arrays; view them with ``regions_numpy`` / ``info_numpy``).
        """
        torch = self.torch
        if use_torch_stream:
            self.ctx.use_torch_stream()
        if isinstance(clips, np.ndarray):
            n_clips = len(clips)
            total = int((clips["out_offset"] + clips["n_frames"]).max()) if n_clips else 0
            d_clips = torch.from_numpy(clips.view(np.uint8).reshape(-1).copy()).to(self.device)
        else:
            raise TypeError("clips must be a numpy CLIP_DTYPE array")
        npx = self.width * self.height
        if out is None:
            out = {}
        def buf(name, shape, dtype):
            t = out.get(name)
            if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
                t = torch.empty(shape, dtype=dtype, device=self.device)
                out[name] = t
            return t
        regions = buf("regions", (max(total, 1), self.max_regions, native.REGION_DTYPE.itemsize), torch.uint8)
        info = buf("info", (max(total, 1), native.INFO_DTYPE.itemsize), torch.uint8)
        filtered = buf("filtered", (max(total, 1), self.height, self.width), torch.float32) if keep_filtered else None
        labels = buf("labels", (max(total, 1), self.height, self.width), torch.uint8) if keep_labels else None
        if d_state is None and keep_state:
            d_state = torch.zeros((max(n_clips, 1), self.ctx.state_bytes), dtype=torch.uint8, device=self.device)
        self.ctx.extract_batch(d_frames, d_clips, n_clips, regions, info, filtered, labels, d_state)
        out.update(regions=regions, info=info, filtered=filtered, labels=labels, state=d_state, total_frames=total,
                   d_clips=d_clips)
        return out

    @staticmethod
    def regions_numpy(regions_tensor):
        a = regions_tensor.cpu().numpy()
        return a.reshape(-1).view(native.REGION_DTYPE).reshape(a.shape[0], a.shape[1])

    @staticmethod
    def info_numpy(info_tensor):
        a = info_tensor.cpu().numpy()
        return a.reshape(-1).view(native.INFO_DTYPE)

    # ------------------------------------------------------------------ host-buffer path (e2e)
    def extract_host(self, h_frames, clips, keep_filtered=False, keep_labels=False, chunk_clips=0, out=None):
        """Host uint16 frames in, host records out: H2D / kernel / D2H pipelined inside the library."""
        h_frames = np.ascontiguousarray(h_frames, dtype=np.uint16)
        total = int((clips["out_offset"] + clips["n_frames"]).max()) if len(clips) else 0
        out = out if out is not None else {}
        def buf(name, shape, dtype):
            a = out.get(name)
            if a is None or a.shape != tuple(shape) or a.dtype != np.dtype(dtype):
                a = native.pinned_empty(shape, dtype)
                out[name] = a
            return a
        regions = buf("regions", (max(total, 1), self.max_regions), native.REGION_DTYPE)
        info = buf("info", (max(total, 1),), native.INFO_DTYPE)
        filtered = buf("filtered", (max(total, 1), self.height, self.width), np.float32) if keep_filtered else None
        labels = buf("labels", (max(total, 1), self.height, self.width), np.uint8) if keep_labels else None
        clips = np.ascontiguousarray(clips)
        self.ctx.extract_batch_host(h_frames, clips, total, regions, info, filtered, labels, chunk_clips)
        out["total_frames"] = total
        return out
