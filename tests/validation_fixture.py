"""Seeded synthetic stand-in for the reference's validation loader (`val_dataset.pytorch(batch_size=1, ...)`,
stage_8_torch.py:984-996) and canned model outputs; shared by tools/make_validation_golden.py (reference side, build container)
and tests/test_validation.py (this repo's driver)."""
import numpy as np
import torch

N_IMAGES = 7
SIZES = [(640, 640), (480, 640), (427, 640), (640, 426), (500, 375), (640, 640), (333, 500)]     # (H, W) of the original images


def synthetic_loader():
    rng = np.random.default_rng(11)
    for i in range(N_IMAGES):
        h, w = SIZES[i]
        m = int(rng.integers(1, 6))
        xy = rng.uniform(0, 0.6, size=(m, 2)) * (w, h)
        wh = rng.uniform(0.05, 0.4, size=(m, 2)) * (w, h)
        boxes = torch.from_numpy(np.concatenate([xy, wh], 1).astype(np.float32))[None]          # (1, m, 4) COCO xywh
        cats = torch.from_numpy(rng.integers(0, 80, size=(1, m)).astype(np.int64))
        img = torch.from_numpy(rng.random((1, 3, h, w), dtype=np.float32))
        yield {'images': img, 'boxes': boxes, 'categories': cats}


def canned_model_outputs():
    """What the model returns per image: (boxes (k,4) xyxy px, classes (k,2) [conf, class]) or (None, None)."""
    rng = np.random.default_rng(12)
    outs = []
    for i in range(N_IMAGES):
        k = [3, 0, 5, 1, 0, 7, 2][i]
        if k == 0:
            outs.append((None, None))
            continue
        xy = rng.uniform(0, 400, size=(k, 2))
        wh = rng.uniform(10, 230, size=(k, 2))
        b = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        c = np.stack([rng.uniform(0.25, 1.0, size=k), rng.integers(0, 80, size=k)], 1).astype(np.float32)
        outs.append((torch.from_numpy(b), torch.from_numpy(c)))
    return outs
