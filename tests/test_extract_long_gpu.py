"""Parity on the benchmark configuration and on the long-clip paths of the recurrence.

* full 900-frame synthetic clips of both camera models (BASELINE configs[1] shape) under both launch plans, bit-exact
  against the C oracle (motiondetector.py:197-244, cliptrackextractor.py:168-176);
* clips long enough for per-pixel weight counters to pass 1024 (the keep-test table then comes from global memory instead
  of the shared-memory prefix), for both weight_add values;
* a keep-test table shorter than the clip: counters that reach the end of the table stop keeping (documented limit of
  cpt_set_weight_table) -- identical to the oracle before any counter gets there, and both launch plans agree after.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def extractor():
    from classifier_pipeline_b200.batch import BatchExtractor

    return BatchExtractor(device=0, max_regions=32)


def _device_run(extractor, clips_pix, bts, was, max_frames, keep_state=True):
    import torch
    from classifier_pipeline_b200.batch import linear_clips

    lengths = [len(p) for p in clips_pix]
    frames = np.concatenate(clips_pix)
    slots = np.array([extractor.ctx.weight_table(w, max_frames=max_frames) for w in was])
    clips = linear_clips(lengths, np.array(bts), slots)
    d_frames = torch.from_numpy(frames.view(np.int16)).cuda().view(torch.uint16)
    out = extractor.extract_device(d_frames, clips, keep_filtered=True, keep_labels=True, keep_state=keep_state, out={})
    torch.cuda.synchronize()
    host = dict(
        filtered=out["filtered"].cpu().numpy(), labels=out["labels"].cpu().numpy(), info=extractor.info_numpy(out["info"]).copy(),
        regions=extractor.regions_numpy(out["regions"]).copy(), clips=clips,
    )
    if keep_state:
        host["state"] = [extractor.ctx.state_read(out["state"], i) for i in range(len(lengths))]
    return host


def _check_clip(host, i, o, n, upto=None):
    o0 = int(host["clips"]["out_offset"][i])
    m = n if upto is None else upto
    sl = slice(o0, o0 + m)
    assert np.array_equal(host["filtered"][sl], o["filtered"][:m]), i
    assert np.array_equal(host["labels"][sl], o["labels"][:m]), i
    info = host["info"][sl]
    assert np.array_equal(info["n_components"], o["ncomp"][:m]), i
    assert np.array_equal(info["threshold"], o["thresh"][:m]), i
    assert np.array_equal(info["background_average"], o["avg"][:m]), i
    assert np.array_equal(info["norm_max"].astype(np.float32), o["norm"][:m, 0]), i
    assert np.array_equal(info["norm_min"].astype(np.float32), o["norm"][:m, 1]), i
    seen = 0
    for t in range(m):
        k = int(o["ncomp"][t])
        seen += k
        r = host["regions"][o0 + t, :k]
        got = np.stack([r["x"], r["y"], r["width"], r["height"], r["area"], r["sum_x"], r["sum_y"], r["key"]], axis=1)
        assert np.array_equal(got, o["comp"][t, :k]), (i, t)
        np.testing.assert_allclose(r["pixel_variance"], o["var"][t, :k], rtol=1e-6, atol=1e-6)
    return seen


@pytest.mark.parametrize("single", [False, True])
def test_bench_config_clips_match_the_oracle(extractor, single):
    """Four 900-frame clips of the bench family (two per camera model), whole outputs bit-exact."""
    from classifier_pipeline_b200.synthetic import clip_model, make_clip
    from oracle import oracle as orc

    idx = [100, 101, 102, 103]
    pix = [make_clip(i, frames=900)[0] for i in idx]
    bts = [clip_model(i)[2] for i in idx]
    was = [clip_model(i)[3] for i in idx]
    extractor.ctx.force_single_kernel(single)
    try:
        host = _device_run(extractor, pix, bts, was, max_frames=1024)
    finally:
        extractor.ctx.force_single_kernel(False)
    seen = 0
    for i in range(len(idx)):
        o = orc.extract_clip(pix[i], pix[i][0], orc.make_params(background_thresh=bts[i], weight_add=was[i], max_comp=32))
        seen += _check_clip(host, i, o, 900)
        st = host["state"][i]
        assert np.array_equal(st["background"], o["final_bg"]), i
        assert st["average"] == o["final_avg"], i
    assert seen > 1000


def _dark_start_clip(index, frames, lift):
    """A clip whose first frame (the one the background is initialised from) is `lift` counts below the rest: every
    pixel keeps its background for as long as `lift` exceeds the accumulated weight, so the counters run up."""
    from classifier_pipeline_b200.synthetic import make_clip

    pix = make_clip(index, frames=frames)[0].astype(np.int64)
    pix[1:] += lift
    return np.clip(pix, 0, 65535).astype(np.uint16)


@pytest.mark.parametrize("weight_add,bt,index,lift", [(1.0, 50, 201, 1800), (0.1, 20, 200, 400)])
@pytest.mark.parametrize("single", [False, True])
def test_counters_past_the_shared_memory_table(extractor, weight_add, bt, index, lift, single):
    """2600-frame clips: counters pass kSmemWeights = 1024 (global-memory table reads); with weight_add = 1 they reach
    `lift` and the background finally resets, with 0.1 they keep for the whole clip."""
    from oracle import oracle as orc

    T = 2600
    pix = _dark_start_clip(index, T, lift)
    extractor.ctx.force_single_kernel(single)
    try:
        host = _device_run(extractor, [pix], [bt], [weight_add], max_frames=4096)
    finally:
        extractor.ctx.force_single_kernel(False)
    o = orc.extract_clip(pix, pix[0], orc.make_params(background_thresh=bt, weight_add=weight_add, max_comp=32))
    _check_clip(host, 0, o, T)
    st = host["state"][0]
    assert np.array_equal(st["background"], o["final_bg"])
    assert st["average"] == o["final_avg"]
    slot = extractor.ctx.weight_table(weight_add, max_frames=4096)
    weights = np.array([extractor.ctx.weight_value(slot, int(k)) for k in range(int(st["weight_count"].max()) + 1)])
    assert np.array_equal(weights[st["weight_count"]], o["final_weight"])
    if weight_add == 0.1:
        assert st["weight_count"].max() > 2000  # the counters really went past the shared-memory prefix


def test_table_shorter_than_the_clip():
    """max_frames = 300 on a 700-frame dark-start clip: the device equals the oracle while no counter has reached the
    end of the table; from there counters stop keeping (their background resets) and both launch plans still agree."""
    from classifier_pipeline_b200.batch import BatchExtractor
    from oracle import oracle as orc

    extractor = BatchExtractor(device=0, max_regions=32)  # (a fresh context: a context keeps the longest table it was asked for)
    T, cap = 700, 300
    pix = _dark_start_clip(202, T, 400)
    runs = []
    for single in (False, True):
        extractor.ctx.force_single_kernel(single)
        try:
            runs.append(_device_run(extractor, [pix], [20], [0.1], max_frames=cap))
        finally:
            extractor.ctx.force_single_kernel(False)
    o = orc.extract_clip(pix, pix[0], orc.make_params(background_thresh=20, weight_add=0.1, max_comp=32))
    for host in runs:
        _check_clip(host, 0, o, T, upto=cap)
        assert host["state"][0]["weight_count"].max() <= cap
    a, b = runs
    assert np.array_equal(a["filtered"][:T], b["filtered"][:T])
    assert np.array_equal(a["labels"][:T], b["labels"][:T])
    assert np.array_equal(a["info"]["n_components"][:T], b["info"]["n_components"][:T])
    assert np.array_equal(a["info"]["threshold"][:T], b["info"]["threshold"][:T])
    assert np.array_equal(a["state"][0]["background"], b["state"][0]["background"])
    # the cap really bit: after it the device differs from the uncapped oracle
    assert not np.array_equal(a["filtered"][cap + 2 : T], o["filtered"][cap + 2 : T])
