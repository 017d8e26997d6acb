#!/usr/bin/env python
"""Generate the committed golden fixtures by running the UNMODIFIED Python reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Every fixture is the reference's own output (``ClipTrackExtractor`` with
``Config.get_defaults()``, optionally ``denoise=False``) captured through wrappers that
only *record* what the reference computes: the background used for each frame, the K2
normalised image and threshold handed to ``detect_objects``, the label image / stats /
centroids it returns, the filtered region lists and the final tracks.  Inputs are the
reference's two test clips (copied as data fixtures to ``tests/golden/clips``) and
seeded synthetic clips from ``classifier_pipeline_b200.synthetic``.
"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402

U_STRIDE = 8  # keep every 8th normalised image in full; CRC32 for all


def region_row(frame_number, r):
    return [
        frame_number,
        r.x,
        r.y,
        r.width,
        r.height,
        r.mass,
        float(r.pixel_variance),
        r.id,
        int(bool(r.was_cropped)),
        int(bool(r.is_along_border)),
        float(r.centroid[0]),
        float(r.centroid[1]),
    ]


def track_dict(track):
    return dict(
        id=track.get_id(),
        start_frame=int(track.start_frame),
        end_frame=int(track.end_frame),
        start_s=track.start_s,
        end_s=track.end_s,
        score=float(track.stats.score),
        stats={k: float(v) for k, v in track.stats._asdict().items()},
        positions=[
            dict(
                x=int(b.x),
                y=int(b.y),
                width=int(b.width),
                height=int(b.height),
                mass=int(b.mass),
                frame_number=int(b.frame_number),
                pixel_variance=float(b.pixel_variance),
                blank=bool(b.blank),
                centroid=[float(b.centroid[0]), float(b.centroid[1])],
            )
            for b in track.bounds_history
        ],
    )


def run_reference(source, denoise):
    """Run the reference extractor on ``source`` (path or registered memory clip)."""
    ref_harness.setup()
    import cv2

    cv2.setNumThreads(1)
    from config.config import Config
    from track.clip import Clip
    from track.cliptrackextractor import ClipTrackExtractor
    import track.cliptrackextractor as cte
    import track.cliptracker as ct

    config = Config.get_defaults()
    config.tracking["thermal"].denoise = denoise
    ext = ClipTrackExtractor(config.tracking, False, cache_to_disk=False)
    clip = Clip(config.tracking["thermal"], source)

    rec = dict(bg=[], avg=[], det=[], norm=[])
    orig_detect, orig_norm = cte.detect_objects, ct.normalize
    state = dict(want_norm=False)

    def detect_wrapper(image, otsus=False, threshold=30, kernel=(15, 15)):
        out = orig_detect(image, otsus=otsus, threshold=threshold, kernel=kernel)
        rec["det"].append(
            dict(
                u=np.uint8(image).copy(),
                threshold=np.float64(threshold),
                labels=out[1].copy(),
                stats=out[2].copy(),
                centroids=out[3].copy(),
            )
        )
        return out

    def norm_wrapper(data, min=None, max=None, new_max=1):
        out = orig_norm(data, min=min, max=max, new_max=new_max)
        if state["want_norm"]:
            state["want_norm"] = False
            rec["norm"].append((float(out[1][1]), float(out[1][2])))
        return out

    orig_gff = ext._get_filtered_frame

    def gff_wrapper(clip_, thermal, sub_change=True, denoise=True):
        # state of the background that frame t is filtered against
        rec["bg"].append(ext.background_alg.background.copy())
        rec["avg"].append(float(ext.background_alg.get_average()))
        state["want_norm"] = True
        return orig_gff(clip_, thermal, sub_change=sub_change, denoise=denoise)

    cte.detect_objects = detect_wrapper
    ct.normalize = norm_wrapper
    ext._get_filtered_frame = gff_wrapper
    try:
        ext.parse_clip(clip)
    finally:
        cte.detect_objects, ct.normalize = orig_detect, orig_norm
    return config, ext, clip, rec


def pack(name, source, denoise, out_dir, input_pix):
    config, ext, clip, rec = run_reference(source, denoise)
    frames = clip.frame_buffer.frames
    T = len(frames)
    assert T == len(rec["det"]) == len(rec["bg"])
    thermal = np.stack([f.thermal for f in frames])
    if input_pix is not None:
        assert np.array_equal(thermal, input_pix[-T:])
    bg = np.stack(rec["bg"])
    assert np.all(bg == np.rint(bg)) and bg.min() >= 0 and bg.max() < 65536
    bg = bg.astype(np.int32)
    bg_delta = np.diff(bg, axis=0).astype(np.int32)
    filtered = np.stack([f.filtered for f in frames])
    assert np.array_equal(filtered, thermal.astype(np.float64) - bg)
    labels = np.stack([d["labels"] for d in rec["det"]])
    assert labels.max() < 256
    for f, d in zip(frames, rec["det"]):
        assert np.array_equal(f.mask, d["labels"])
    u = np.stack([d["u"] for d in rec["det"]])
    ncomp = np.array([len(d["stats"]) for d in rec["det"]], dtype=np.int32)
    stats = np.concatenate([d["stats"] for d in rec["det"]]).astype(np.int32)
    cents = np.concatenate([d["centroids"] for d in rec["det"]]).astype(np.float64)
    regions = [region_row(t, r) for t, rs in enumerate(clip.region_history) for r in rs]
    regions = np.array(regions, dtype=np.float64).reshape(-1, 12)
    meta = dict(
        name=name,
        denoise=bool(denoise),
        camera_model=clip.camera_model,
        background_thresh=int(clip.background_thresh),
        weight_add=float(ext.background_alg.weight_add),
        frames=T,
        res=[int(clip.res_x), int(clip.res_y)],
        input_crc=int(zlib.crc32(thermal.tobytes())),
        ffc_frames=[int(x) for x in clip.ffc_frames],
        tracks=[track_dict(t) for t in clip.tracks],
        filtered_tracks=[[reason, track_dict(t)] for reason, t in clip.filtered_tracks],
        tracker_version=ext.tracker_version,
        numpy=np.__version__,
        cv2=__import__("cv2").__version__,
    )
    stats_obj = clip.stats
    np.savez_compressed(
        os.path.join(out_dir, name + ".npz"),
        meta=np.array(json.dumps(meta)),
        bg_first=bg[0].astype(np.uint16),
        bg_delta=bg_delta,
        bg_final=ext.background_alg.background.astype(np.int32),
        weight_final=ext.background_alg.background_weight.astype(np.float64),
        avg_final=np.float64(ext.background_alg.average),
        avg=np.array(rec["avg"], dtype=np.float64),
        norm_max=np.array([m[0] for m in rec["norm"]], dtype=np.float64),
        norm_min=np.array([m[1] for m in rec["norm"]], dtype=np.float64),
        thresh=np.array([d["threshold"] for d in rec["det"]], dtype=np.float64),
        u_crc=np.array([zlib.crc32(x.tobytes()) for x in u], dtype=np.uint32),
        u_sub=u[::U_STRIDE].copy(),
        labels=labels.astype(np.uint8),
        ncomp=ncomp,
        stats=stats,
        centroids=cents,
        regions=regions,
        fs_min=np.array(stats_obj.frame_stats_min, dtype=np.float64),
        fs_max=np.array(stats_obj.frame_stats_max, dtype=np.float64),
        fs_median=np.array(stats_obj.frame_stats_median, dtype=np.float64),
        fs_mean=np.array(stats_obj.frame_stats_mean, dtype=np.float64),
        filtered_sum=np.float64(stats_obj.filtered_sum),
    )
    print(
        "{:28s} frames={:4d} tracks={} filtered={} regions={} size={:.0f} kB".format(
            name,
            T,
            len(clip.tracks),
            len(clip.filtered_tracks),
            len(regions),
            os.path.getsize(os.path.join(out_dir, name + ".npz")) / 1e3,
        )
    )
    return clip


def main():
    from classifier_pipeline_b200.synthetic import make_clip

    out_dir = HERE
    clips_dir = os.path.join(HERE, "clips")
    for clip_name in ("possum", "hedgehog"):
        path = os.path.join(clips_dir, clip_name + ".cptv")
        for denoise in (True, False):
            pack("{}_{}".format(clip_name, "nlm" if denoise else "raw"), path, denoise, out_dir, None)
    for index, frames, denoise in ((0, 120, False), (1, 120, False), (2, 100, False), (3, 100, False), (4, 48, True)):
        pix, model = make_clip(index, frames=frames)
        key = "synthetic-{}-{}".format(index, frames)
        ref_harness.register_memory_clip(key, pix, model)
        pack("synth{}_{}".format(index, "nlm" if denoise else "raw"), key, denoise, out_dir, pix)


if __name__ == "__main__":
    main()
