"""``RawDatabase``: a CPTV file plus its metadata ``.txt`` as the dataset builder sees it
(reference: ``src/ml_tools/rawdb.py:42-147``).

``load_frames`` is the second caller of the background recurrence: every kept frame becomes a
``Frame(thermal, thermal - background, n)`` while a ``WeightedBackground`` is updated with the mean of the last 45 kept
frames.  Here the file is inflated on the host, its frames are decoded on the device (``csrc/cptv_kernels.cu``) and the
recurrence runs in the extraction kernels (``csrc/extract_kernel.cu``) -- the per-pixel work never touches the CPU.
``load_frames_batch`` does the same for many files with one decode and one extraction launch.

Reference behaviour kept on purpose:

* the first frame of the file initialises the background even when it is a background frame that is then skipped
  (``tracker_version >= 10``, rawdb.py:104-108);
* when that first frame is *kept*, no background update follows it (``back_processed``, rawdb.py:84-122) --
  ``CPT_CLIP_SKIP_FIRST_UPDATE``;
* ``self.background`` is the live array of the ``WeightedBackground`` (rawdb.py:102), i.e. the background after the
  LAST frame;
* ``Frame.filtered`` is float64 (``np.float32(pix) - float64 background``); the values are integers, so widening the
  device's fp32 image is exact;
* the model is decided by the mean of the first frame (rawdb.py:86-92), which also picks the weight step.
"""
import json
import logging
from pathlib import Path

import numpy as np

from .. import native
from ..cptv import CptvReader, decode_clips_device
from ..engine import get_engine
from ..piclassifier.cptvmotiondetector import is_affected_by_ffc
from .frame import Frame
from .rectangle import Rectangle

special_datasets = [
    "tag_frames",
    "original_frames",
    "background_frame",
    "predictions",
    "overlay",
]

FPS = 9

RES_X = 160
RES_Y = 120


def plan_frames(background_flags, tracker_version):
    """Which file frames ``load_frames`` keeps, and whether the background update after the first kept frame is
    skipped (host logic of rawdb.py:80-122).  Returns ``(kept indices, skip_first_update)``."""
    kept = [i for i, is_background in enumerate(background_flags) if not (is_background and tracker_version >= 10)]
    # the frame that initialised the background (file frame 0) was "back_processed": if it is kept, it is frame 0 of
    # the clip and no update follows it
    return kept, bool(kept) and kept[0] == 0


class RawDatabase:
    def __init__(self, database_filename):
        self.file = Path(database_filename)
        self.meta_data_file = self.file.with_suffix(".txt")
        self._meta_data = None
        self.background = None
        self.ffc_frames = None
        self.frames = None
        self.model = None
        self.crop_rectangle = Rectangle(1, 1, 160 - 2, 120 - 2)

    def frames_kept(self):
        return None

    def get_frame(self, frame_number):
        if self.frames is None or frame_number > len(self.frames):
            return None
        return self.frames[frame_number]

    def get_frames(self):
        return self.frames

    def get_clip_background(self):
        return self.background

    def load_frames(self):
        load_frames_batch([self])

    @property
    def meta_data(self):
        if self._meta_data is not None:
            return self._meta_data
        if not self.meta_data_file.is_file():
            logging.warning("Could not load meta data for %s", self.meta_data_file)
            return None
        with open(self.meta_data_file, "r") as t:
            # add in some metadata stats
            self._meta_data = json.load(t)
        return self._meta_data


def load_frames_batch(databases, device=None):
    """``RawDatabase.load_frames`` for many files: one device decode and one extraction launch for all of them."""
    import torch

    databases = list(databases)
    if not databases:
        return
    engine = get_engine(device, RES_X, RES_Y, 1)
    readers, versions = [], []
    for db in databases:
        # (the reference reads the version before opening the file: a missing metadata file fails here, as it does there)
        versions.append(db.meta_data.get("tracker_version", 11))
        reader = CptvReader(str(db.file))
        reader.get_header()
        readers.append(reader)
    d_frames, clip_first, metas = decode_clips_device(engine, readers)

    # per clip: kept frames, model (mean of the first frame, rawdb.py:86-92)
    n_clips = len(databases)
    plans = [plan_frames([f.background_frame for f in metas[i]], versions[i]) for i in range(n_clips)]
    first_idx = torch.tensor([clip_first[i] for i in range(n_clips) if clip_first[i + 1] > clip_first[i]], device=d_frames.device)
    first_mean = {}
    if len(first_idx):
        means = d_frames.view(torch.int16)[first_idx].to(torch.int32).bitwise_and(0xFFFF).to(torch.float64).mean(dim=(1, 2)).cpu().numpy()
        k = 0
        for i in range(n_clips):
            if clip_first[i + 1] > clip_first[i]:
                first_mean[i] = float(means[k])
                k += 1

    # gather the kept frames of every clip back to back (init frame first) so that each clip is a linear run
    gather, lengths, init_offsets, frame_offsets, flags, bts, slots = [], [], [], [], [], [], []
    pos = 0
    for i, db in enumerate(databases):
        kept, skip_first = plans[i]
        n_file = clip_first[i + 1] - clip_first[i]
        if n_file == 0:
            db.model = None
            lengths.append(0); init_offsets.append(pos); frame_offsets.append(pos); flags.append(0); bts.append(0); slots.append(0)
            continue
        if first_mean[i] > 10000:
            db.model, weight_add = "lepton3.5", 1
        else:
            db.model, weight_add = "lepton3", 0.1
        base = clip_first[i]
        gather.append(base)                      # the frame that initialises the background
        gather.extend(base + j for j in kept)
        init_offsets.append(pos)
        frame_offsets.append(pos + 1)
        lengths.append(len(kept))
        flags.append(native.CLIP_UPDATE_BACKGROUND | (native.CLIP_SKIP_FIRST_UPDATE if skip_first else 0))
        bts.append(30000)                        # masks / regions are not used here: a threshold nothing reaches
        slots.append(engine.ctx.weight_table(weight_add, max_frames=max(len(kept) + 2, 1024)))
        pos += 1 + len(kept)
    if pos == 0:
        for db in databases:
            db.frames, db.ffc_frames, db.background = [], [], None
        return
    idx = torch.tensor(gather, dtype=torch.long, device=d_frames.device)
    d_run = d_frames.view(torch.int16)[idx].contiguous().view(torch.uint16)

    from ..batch import linear_clips

    clips = linear_clips(lengths, np.array(bts), np.array(slots), flags=np.array(flags))
    clips["frame_offset"] = np.array(frame_offsets)
    clips["init_offset"] = np.array(init_offsets)
    out = engine.extract_device(d_run, clips, keep_filtered=True, keep_labels=False, keep_state=True, out={})
    torch.cuda.synchronize()
    filtered = out["filtered"].cpu().numpy()
    thermal = d_run.view(torch.int16).cpu().numpy().view(np.uint16)
    for i, db in enumerate(databases):
        kept, _ = plans[i]
        o0 = int(clips["out_offset"][i])
        f0 = frame_offsets[i]
        db.frames = [
            Frame(thermal[f0 + n], filtered[o0 + n].astype(np.float64), n) for n in range(len(kept))
        ]
        db.ffc_frames = [n for n, j in enumerate(kept) if is_affected_by_ffc(metas[i][j])]
        if clip_first[i + 1] == clip_first[i]:
            db.background = None
        elif kept:
            state = engine.ctx.state_read(out["state"], i)
            db.background = state["background"].astype(np.float64)
        else:
            # only the initialising frame: WeightedBackground.process_frame's first call (motiondetector.py:199-210, 239-244)
            bg = thermal[init_offsets[i]].astype(np.float64)
            bg[0], bg[-1] = bg[1], bg[-2]
            bg[:, 0], bg[:, -1] = bg[:, 1], bg[:, -2]
            db.background = bg
