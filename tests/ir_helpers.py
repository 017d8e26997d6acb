"""Synthetic 640x480 IR material shared by the fixture generator (tests/golden/make_golden_ir.py) and the GPU tests."""
import types

import numpy as np

W, H = 640, 480


def ir_filtered(seed, n_blobs=6):
    """A background-subtracted IR frame: dark with a few fragmented bright patches and salt noise (uint8)."""
    rng = np.random.default_rng(seed)
    img = np.zeros((H, W), np.float64)
    ys, xs = np.mgrid[0:H, 0:W]
    for _ in range(n_blobs):
        cx, cy = rng.uniform(40, W - 40), rng.uniform(40, H - 40)
        for _ in range(int(rng.integers(1, 5))):  # fragments of one animal
            fx, fy = cx + rng.uniform(-45, 45), cy + rng.uniform(-35, 35)
            sx, sy = rng.uniform(3, 14), rng.uniform(3, 12)
            img += rng.uniform(80, 250) * np.exp(-((xs - fx) ** 2 / (2 * sx * sx) + (ys - fy) ** 2 / (2 * sy * sy)))
    img[img < 25] = 0
    salt = rng.random((H, W)) > 0.9995
    img[salt] = rng.uniform(30, 255, salt.sum())
    return np.clip(img, 0, 255).astype(np.uint8)


def thermal_like(seed, shape=(120, 160)):
    """A normalised 0..255 image with warm blobs on noise (input of detect_objects with its default kernel / Otsu)."""
    rng = np.random.default_rng(seed)
    h, w = shape
    ys, xs = np.mgrid[0:h, 0:w]
    img = rng.normal(40, 9, shape)
    for _ in range(int(rng.integers(1, 4))):
        cx, cy, s = rng.uniform(0, w), rng.uniform(0, h), rng.uniform(4, 12)
        img += rng.uniform(60, 200) * np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2 * s * s))
    return np.clip(img, 0, 255)


def ir_video(seed, frames=140):
    """BGR frames: a static textured scene with sensor noise; a bright box walks through from frame 112 to 128."""
    rng = np.random.default_rng(seed)
    scene = rng.integers(30, 120, (H // 8, W // 8, 3)).repeat(8, 0).repeat(8, 1).astype(np.int16)
    out = []
    for t in range(frames):
        f = scene + rng.integers(-3, 4, (H, W, 3))
        if 112 <= t < 128:
            x0 = 60 + (t - 112) * 30
            f[200:280, x0 : x0 + 70] += 90
        out.append(np.clip(f, 0, 255).astype(np.uint8))
    return out


class _Window:
    start = types.SimpleNamespace(dt="00:00")
    end = types.SimpleNamespace(dt="00:00")

    def use_sunrise_sunset(self):
        return False

    def inside_window(self):
        return True


def ir_config():
    cfg = types.SimpleNamespace(recorder=types.SimpleNamespace(use_low_power_mode=False, rec_window=_Window(), preview_secs=1, min_secs=5, max_secs=600),
                                location=types.SimpleNamespace())
    headers = types.SimpleNamespace(model="IR", res_x=W, res_y=H, fps=10)
    return cfg, headers
