"""File-level driver (track/trackextractor.py:122-251): one clip file -> tracks -> the metadata JSON the reference's
``extract.py`` writes next to it (``Clip.get_metadata`` + a thumbnail per track + tracker version / config)."""
import json
import logging
from pathlib import Path

from ..classify.thumbnail import best_trackless_thumb, get_thumbnail_info
from ..ml_tools import tools
from .clip import Clip
from .cliptrackextractor import ClipTrackExtractor


def extract_file(filename, config, cache_to_disk, retrack=False, to_stdout=False, max_frames=None, save_meta=True):
    filename = Path(filename)
    if not filename.is_file():
        raise Exception("File {} not found.".format(filename))
    logging.info("Tracking %s", filename)
    if filename.suffix != ".cptv":
        raise NotImplementedError("only thermal .cptv clips are tracked here (the reference's IR path does not run as written)")
    track_extractor = ClipTrackExtractor(config.tracking, config.use_opt_flow, cache_to_disk, verbose=config.verbose, max_frames=max_frames)
    clip = Clip(track_extractor.config, filename)
    clip.frames_per_second = 9
    existing_metadata = None
    if filename.with_suffix(".txt").exists():
        existing_metadata = tools.load_clip_metadata(filename.with_suffix(".txt"))
    if retrack:
        logging.info("Retracking")
        clip.load_metadata(existing_metadata)
    if not track_extractor.parse_clip(clip):
        logging.error("Could not parse %s", filename)
        return None
    if retrack:
        for track in clip.tracks:
            track.trim()
            track.set_end_s(clip.frames_per_second)
    meta_filename = filename.with_suffix(".txt")
    logging.info("saving meta data %s", meta_filename)
    metadata = get_metadata(existing_metadata, filename, meta_filename, clip, track_extractor, to_stdout, save_meta)
    if cache_to_disk:
        clip.frame_buffer.remove_cache()
    return clip, track_extractor, metadata


def get_metadata(existing_metadata, filename, meta_filename, clip, track_extractor, to_stdout=False, save=True):
    metadata = clip.get_metadata()
    for i, track in enumerate(clip.tracks):
        best_thumb, best_score = get_thumbnail_info(clip, track)
        if best_thumb is None:
            metadata["tracks"][i]["thumbnail"] = None
            continue
        metadata["tracks"][i]["thumbnail"] = {
            "region": best_thumb.region, "contours": best_thumb.contours, "median_diff": best_thumb.median_diff,
            "score": round(best_score),
        }
    if len(clip.tracks) == 0:
        metadata["thumbnail_region"] = best_trackless_thumb(clip)  # no tracks: choose a clip thumb
    metadata["source"] = str(filename)
    metadata["tracking_time"] = round(track_extractor.tracking_time, 1)
    metadata["algorithm"] = {"tracker_version": track_extractor.tracker_version, "tracker_config": track_extractor.config.as_dict()}
    if existing_metadata is not None:
        # merge new metadata with old: the tracks are all that is replaced
        existing_metadata.pop("tracks", None)
        existing_metadata.pop("Tracks", None)
        existing_metadata.update(metadata)
        metadata = existing_metadata
    if to_stdout:
        print(json.dumps(metadata, cls=tools.CustomJSONEncoder))
    elif save:
        with open(meta_filename, "w") as f:
            json.dump(metadata, f, indent=4, cls=tools.CustomJSONEncoder)
    return metadata
