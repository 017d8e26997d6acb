"""``detect_objects`` on the device for one image of any size (ml_tools/imageprocessing.py:240-248):
uint8 cast -> GaussianBlur -> threshold -> morphologyEx CLOSE -> connectedComponentsWithStats."""
import numpy as np

from .. import engine as _engine

MAX_COMPONENTS = 8192


def detect_objects(image, otsus=False, threshold=30, kernel=(15, 15)):
    """Returns (n_labels, labels int32 (H, W), stats int32 (n, 5), centroids float64 (n, 2)), row 0 = background."""
    import torch

    if otsus:
        raise NotImplementedError("detect_objects(otsus=True): Otsu thresholding is not built (the tracker never asks for it)")
    kernel = tuple(kernel)
    if kernel != (5, 5):
        # cv2.GaussianBlur derives sigma and fixed-point taps from the kernel size; only the tracker's (5, 5) is built
        raise NotImplementedError("detect_objects: only the (5, 5) kernel the tracker uses is built")
    image = np.uint8(image)  # numpy's cast, exactly as the reference
    if image.ndim != 2:
        raise ValueError("detect_objects: single-channel image expected")
    H, W = image.shape
    eng = _engine.get_engine()
    eng.ctx.use_torch_stream()
    d_img = torch.from_numpy(np.ascontiguousarray(image)).to(eng.device)
    d_labels = torch.empty((H, W), dtype=torch.int32, device=eng.device)
    d_stats = torch.empty((MAX_COMPONENTS + 1, 5), dtype=torch.int32, device=eng.device)
    d_cent = torch.empty((MAX_COMPONENTS + 1, 2), dtype=torch.float64, device=eng.device)
    # a tuple "kernel" reaches cv2.morphologyEx as a 2x1 structuring element (both entries non-zero)
    n = eng.ctx.detect_objects_u8(d_img, W, H, threshold, 5, 1, MAX_COMPONENTS, d_labels, d_stats, d_cent)
    return n, d_labels.cpu().numpy(), d_stats[:n].cpu().numpy(), d_cent[:n].cpu().numpy()
