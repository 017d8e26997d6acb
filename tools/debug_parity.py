"""Stage-by-stage mismatch report of the CUDA kernel vs the C oracle (GPU box, debugging aid)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def first_bad(a, b):
    bad = np.nonzero((a != b).reshape(len(a), -1).any(axis=1))[0]
    return (int(bad[0]), len(bad)) if len(bad) else None


def main(names):
    import torch

    from classifier_pipeline_b200 import native
    from classifier_pipeline_b200.batch import BatchExtractor, linear_clips
    from oracle import oracle as orc
    from tests import helpers

    ex = BatchExtractor(device=0, max_regions=32)
    for name in names:
        d, meta = helpers.load_golden(name)
        init, tracked = helpers.clip_input(name)
        T = len(tracked)
        o = orc.extract_clip(tracked, init, orc.make_params(background_thresh=meta["background_thresh"], weight_add=meta["weight_add"], max_comp=32))
        frames = np.concatenate([init[None], tracked])
        d_frames = torch.from_numpy(frames.view(np.int16)).cuda().view(torch.uint16)
        slot = ex.ctx.weight_table(meta["weight_add"], max_frames=4096)
        clips = linear_clips([T], meta["background_thresh"], slot)
        clips["frame_offset"] = 1
        clips["init_offset"] = 0
        clips["out_offset"] = 0
        out = ex.extract_device(d_frames, clips, keep_filtered=True, keep_labels=True, keep_state=True, out={})
        torch.cuda.synchronize()
        info = ex.info_numpy(out["info"])[:T]
        regions = ex.regions_numpy(out["regions"])[:T]
        filt = out["filtered"][:T].cpu().numpy()
        labels = out["labels"][:T].cpu().numpy()
        print("==", name, "T", T)
        fb = first_bad(filt, o["filtered"])
        print(" filtered first bad:", fb)
        if fb:
            t = fb[0]
            diff = np.nonzero(filt[t] != o["filtered"][t])
            print("   n px", len(diff[0]), "examples", [(int(y), int(x), float(filt[t, y, x]), float(o["filtered"][t, y, x])) for y, x in list(zip(*diff))[:6]])
        for key, ref in (("background_average", o["avg"]), ("threshold", o["thresh"]), ("n_components", o["ncomp"])):
            bad = np.nonzero(info[key] != ref)[0]
            print(" ", key, "bad frames", bad[:8], [(info[key][i], ref[i]) for i in bad[:4]])
        bad = np.nonzero(info["norm_max"].astype(np.float32) != o["norm"][:, 0])[0]
        print("  norm_max bad", bad[:8])
        fb = first_bad(labels, o["labels"])
        print(" labels first bad:", fb)
        if fb:
            t = fb[0]
            diff = np.nonzero(labels[t] != o["labels"][t])
            print("   n px", len(diff[0]), [(int(y), int(x), int(labels[t, y, x]), int(o["labels"][t, y, x])) for y, x in list(zip(*diff))[:10]])
        nbad = 0
        for t in range(T):
            n = min(int(o["ncomp"][t]), 32)
            r = regions[t, :n]
            got = np.stack([r["x"], r["y"], r["width"], r["height"], r["area"], r["sum_x"], r["sum_y"], r["key"]], axis=1)
            if not np.array_equal(got, o["comp"][t, :n]):
                if nbad < 3:
                    print("  region mismatch t", t, got.tolist(), o["comp"][t, :n].tolist())
                nbad += 1
            elif n and not np.allclose(r["pixel_variance"], o["var"][t, :n], rtol=1e-6, atol=1e-6):
                if nbad < 3:
                    print("  variance mismatch t", t, r["pixel_variance"], o["var"][t, :n])
                nbad += 1
        print(" region frames bad:", nbad)
        st = ex.ctx.state_read(out["state"], 0)
        print(" final bg equal:", np.array_equal(st["background"], o["final_bg"]), "avg", st["average"], o["final_avg"])


if __name__ == "__main__":
    main(sys.argv[1:] or ["possum_raw", "synth1_raw"])
