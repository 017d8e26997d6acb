#!/usr/bin/env python
"""Golden fixtures for ``RawDatabase.load_frames`` from the UNMODIFIED reference (``/root/reference/src/ml_tools/rawdb.py``),
run in the build container:  python tests/golden/make_golden_rawdb.py  ->  tests/golden/rawdb_<clip>.npz

Per clip: the model, the FFC frame list, the final background, a CRC32 of every frame's thermal and filtered image and
a few filtered images in full.  possum.cptv starts with a background frame (skipped, tracker_version 11); hedgehog.cptv
has none, so its first frame is kept and not followed by a background update."""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402

FULL = (0, 1, 2, 45, 46, 100)


def main():
    ref_harness.setup()
    from ml_tools.rawdb import RawDatabase

    for name in ("possum", "hedgehog"):
        db = RawDatabase(os.path.join(HERE, "clips", name + ".cptv"))
        db._meta_data = {"tracker_version": 11}
        db.load_frames()
        frames = db.frames
        assert all(f.filtered.dtype == np.float64 for f in frames)
        out = dict(
            n_frames=len(frames), model=db.model, ffc_frames=np.array(db.ffc_frames, dtype=np.int64),
            background=db.background,
            thermal_crc=np.array([zlib.crc32(np.ascontiguousarray(f.thermal).tobytes()) for f in frames], dtype=np.uint32),
            filtered_crc=np.array([zlib.crc32(np.ascontiguousarray(f.filtered.astype(np.float32)).tobytes()) for f in frames], dtype=np.uint32),
            full_index=np.array([i for i in FULL if i < len(frames)]),
            full_filtered=np.stack([frames[i].filtered for i in FULL if i < len(frames)]),
        )
        np.savez_compressed(os.path.join(HERE, "rawdb_%s.npz" % name), **out)
        print(name, len(frames), db.model, db.ffc_frames[:4], float(db.background.sum()))


if __name__ == "__main__":
    main()
