"""Deterministic synthetic Lepton-shaped clips.

A clip is ``(frames, 120, 160)`` uint16: a static scene ``base + 40 sin(x/25) + 30 cos(y/18)
+ N(0,6)`` fixed per clip, per-frame sensor noise ``N(0,4)`` and one or two warm Gaussian
blobs that drift with constant velocity and wrap at the borders.  Even clip indices are
``lepton3`` (base 3000, background_thresh 20, weight_add 0.1), odd ones ``lepton3.5``
(base 29000, background_thresh 50, weight_add 1).  Clip *i* uses ``seed = 1234 + i``.

``make_clip`` is the numpy generator used by the parity tests and the CPU baseline;
``make_clips_torch`` generates the same family of clips directly in device memory for
the large bench batches (same statistics, different random stream).
"""
import numpy as np

WIDTH = 160
HEIGHT = 120
BASE_SEED = 1234

MODELS = (
    # name, base level, background_thresh, weight_add
    ("lepton3", 3000.0, 20, 0.1),
    ("lepton3.5", 29000.0, 50, 1.0),
)


def clip_model(index):
    return MODELS[index % 2]


def make_clip(index, frames=900, width=WIDTH, height=HEIGHT, seed=None):
    """Return ``(pix uint16 (frames,height,width), model_name)`` for clip ``index``."""
    name, base, _, _ = clip_model(index)
    rng = np.random.default_rng(BASE_SEED + index if seed is None else seed)
    xs = np.arange(width, dtype=np.float64)[None, :]
    ys = np.arange(height, dtype=np.float64)[:, None]
    scene = base + 40.0 * np.sin(xs / 25.0) + 30.0 * np.cos(ys / 18.0) + rng.normal(0.0, 6.0, (height, width))
    n_blobs = int(rng.integers(1, 3))
    blobs = []
    for _ in range(n_blobs):
        blobs.append(
            dict(
                sigma=rng.uniform(3.0, 7.0),
                amp=rng.uniform(120.0, 400.0),
                vx=rng.uniform(-1.5, 1.5),
                vy=rng.uniform(-1.0, 1.0),
                birth=int(rng.uniform(0, frames / 2.0)),
                x0=rng.uniform(0, width),
                y0=rng.uniform(0, height),
            )
        )
    out = np.empty((frames, height, width), dtype=np.uint16)
    for t in range(frames):
        img = scene + rng.normal(0.0, 4.0, (height, width))
        for b in blobs:
            if t < b["birth"]:
                continue
            age = t - b["birth"]
            cx = (b["x0"] + b["vx"] * age) % width
            cy = (b["y0"] + b["vy"] * age) % height
            dx = np.abs(xs - cx)
            dx = np.minimum(dx, width - dx)
            dy = np.abs(ys - cy)
            dy = np.minimum(dy, height - dy)
            img = img + b["amp"] * np.exp(-(dx * dx + dy * dy) / (2.0 * b["sigma"] ** 2))
        out[t] = np.clip(np.rint(img), 0, 65535).astype(np.uint16)
    return out, name


def make_clips_torch(n_clips, frames, device, first_index=0, width=WIDTH, height=HEIGHT, chunk=16):
    """Same clip family generated with torch on ``device`` → uint16 tensor viewed as int16 storage.

    Returns ``(tensor (n_clips, frames, height, width) torch.uint16, model index per clip)``.
    """
    import torch

    out = torch.empty((n_clips, frames, height, width), dtype=torch.uint16, device=device)
    xs = torch.arange(width, dtype=torch.float32, device=device)[None, None, None, :]
    ys = torch.arange(height, dtype=torch.float32, device=device)[None, None, :, None]
    ts = torch.arange(frames, dtype=torch.float32, device=device)[None, :, None, None]
    for c0 in range(0, n_clips, chunk):
        c1 = min(n_clips, c0 + chunk)
        n = c1 - c0
        gen = torch.Generator(device=device)
        gen.manual_seed(BASE_SEED + first_index + c0)
        idx = torch.arange(first_index + c0, first_index + c1, device=device)
        base = torch.where(idx % 2 == 0, 3000.0, 29000.0).float()[:, None, None, None]
        scene = (
            base
            + 40.0 * torch.sin(xs / 25.0)
            + 30.0 * torch.cos(ys / 18.0)
            + 6.0 * torch.randn((n, 1, height, width), device=device, generator=gen)
        )
        img = scene + 4.0 * torch.randn((n, frames, height, width), device=device, generator=gen)

        def u(lo, hi):
            return lo + (hi - lo) * torch.rand((n, 1, 1, 1), device=device, generator=gen)

        two = torch.rand((n, 1, 1, 1), device=device, generator=gen) < 0.5
        for b in range(2):
            sigma, amp = u(3.0, 7.0), u(120.0, 400.0)
            vx, vy = u(-1.5, 1.5), u(-1.0, 1.0)
            birth = torch.floor(u(0.0, frames / 2.0))
            x0, y0 = u(0.0, width), u(0.0, height)
            age = ts - birth
            cx = torch.remainder(x0 + vx * age, width)
            cy = torch.remainder(y0 + vy * age, height)
            dx = (xs - cx).abs()
            dx = torch.minimum(dx, width - dx)
            dy = (ys - cy).abs()
            dy = torch.minimum(dy, height - dy)
            alive = age >= 0
            if b == 1:
                alive = alive & two
            img = img + torch.where(alive, amp, 0.0) * torch.exp(-(dx * dx + dy * dy) / (2.0 * sigma * sigma))
        v = torch.clamp(torch.round(img), 0, 65535).to(torch.int32)
        out[c0:c1].view(torch.int16).copy_(torch.where(v >= 32768, v - 65536, v).to(torch.int16))
    models = [(first_index + i) % 2 for i in range(n_clips)]
    return out, models


def pack_clips_torch(d_frames):
    """CPTV v2 frame payloads of clips resident on the device (the packing ``cptv/reader.py`` undoes): per frame an int32
    start value, then ``W*H - 1`` two's-complement deltas along the boustrophedon scan of the frame-to-frame change, 8 bits
    each where they fit and 16 otherwise, MSB first.

    ``d_frames``: ``(n_clips, frames, H, W)`` torch.uint16 CUDA tensor.  Returns ``(stream, table, clip_first)``: the
    payloads back to back as a pinned uint8 array (plus four slack bytes), one ``native.CPTV_FRAME_DTYPE`` row per frame
    and the first row of every clip."""
    import torch

    from . import native

    C, T, H, W = d_frames.shape
    n = H * W
    sizes = np.empty((C, T), np.int64)
    wide = np.zeros((C, T), bool)
    parts8, parts16, starts = [], [], []
    for c in range(C):
        cur = d_frames[c].view(torch.int16).to(torch.int32) & 0xFFFF
        change = cur - torch.cat([torch.zeros_like(cur[:1]), cur[:-1]])
        change[:, 1::2] = change[:, 1::2].flip(-1)  # odd rows right to left
        change = change.reshape(T, n)
        deltas = change[:, 1:] - change[:, :-1]
        need16 = (deltas.abs().amax(dim=1) > 127) | (deltas.amin(dim=1) < -128)
        wide[c] = need16.cpu().numpy()
        starts.append(change[:, 0].cpu().numpy().astype("<i4"))
        parts8.append(deltas.to(torch.int8).cpu().numpy())
        if wide[c].any():
            d16 = deltas[need16].to(torch.int16)
            parts16.append(((d16 >> 8) & 0xFF).to(torch.uint8).cpu().numpy()[..., None].repeat(2, -1))
            parts16[-1][..., 1] = (d16 & 0xFF).to(torch.uint8).cpu().numpy()
        else:
            parts16.append(None)
        sizes[c] = 4 + np.where(wide[c], 2 * (n - 1), n - 1)
    offsets = np.concatenate([[0], np.cumsum(sizes.reshape(-1))])
    stream = native.pinned_empty((int(offsets[-1]) + 4,), np.uint8)
    stream[-4:] = 0
    table = np.zeros(C * T, native.CPTV_FRAME_DTYPE)
    table["payload_offset"] = offsets[:-1]
    table["bit_width"] = np.where(wide.reshape(-1), 16, 8)
    for c in range(C):
        k16 = 0
        for t in range(T):
            o = int(offsets[c * T + t])
            stream[o : o + 4] = np.frombuffer(starts[c][t].tobytes(), np.uint8)
            if wide[c, t]:
                stream[o + 4 : o + 4 + 2 * (n - 1)] = parts16[c][k16].reshape(-1)
                k16 += 1
            else:
                stream[o + 4 : o + 4 + n - 1] = parts8[c][t].view(np.uint8)
    return stream, table, np.arange(C + 1, dtype=np.int64) * T
