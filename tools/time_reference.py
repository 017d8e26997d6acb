"""Time the UNMODIFIED reference's own extractor (ClipTrackExtractor.parse_clip: Python + numpy + OpenCV) on the bench's
synthetic clip family, in the build container (the only place /root/reference exists): one process per core, as
track/trackextractor.py:80-85 does with its multiprocessing.Pool, cv2 threads pinned to 1.  Writes the committed record
profiles/reference_python_cpu.json that bench.py quotes beside the live C-port baseline.

    python tools/time_reference.py [frames_per_clip] [clips_per_core]
"""
import json
import multiprocessing
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def work(args):
    index, frames, denoise = args
    import cv2

    cv2.setNumThreads(1)
    from tests.golden import ref_harness

    ref_harness.setup()
    from classifier_pipeline_b200.synthetic import make_clip
    from config.config import Config
    from track.clip import Clip
    from track.cliptrackextractor import ClipTrackExtractor

    pix, model = make_clip(index, frames=frames)
    key = "mem{}".format(index)
    ref_harness.register_memory_clip(key, pix, model)
    config = Config.get_defaults()
    config.tracking["thermal"].denoise = denoise
    ext = ClipTrackExtractor(config.tracking, False, cache_to_disk=False)
    clip = Clip(config.tracking["thermal"], key)
    t0 = time.perf_counter()
    ext.parse_clip(clip)
    return time.perf_counter() - t0, len(clip.tracks)


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    per_core = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    cores = os.cpu_count() or 1
    out = {"where": "build container (no GPU), {} cores".format(cores), "script": "tools/time_reference.py",
           "workload": "{} synthetic clips x {} frames per run, one process per core, cv2.setNumThreads(1)".format(cores * per_core, frames)}
    for denoise in (False, True):
        n = cores * per_core
        f = frames if not denoise else max(frames // 6, 30)
        t0 = time.perf_counter()
        with multiprocessing.Pool(cores) as pool:
            res = pool.map(work, [(i, f, denoise) for i in range(n)])
        wall = time.perf_counter() - t0
        key = "denoise_on" if denoise else "denoise_off"
        out[key] = {"frames_per_s_all_cores": n * f / wall, "frames_per_s_per_core": f / (sum(r[0] for r in res) / n),
                    "wall_s": wall, "tracks": int(sum(r[1] for r in res)), "frames_per_clip": f}
    json.dump(out, open(os.path.join(ROOT, "profiles", "reference_python_cpu.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
