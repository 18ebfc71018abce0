"""Seeded synthetic workload generators -- TEST INFRASTRUCTURE (shared by the reference
harness, tests/, bench.py's input generation and __graft_entry__.smoke()).

yolov8n.pt and COCO are not available offline (BASELINE.json north_star), so:
  * images are integer-only functions of a seed (bit-stable across hosts), returned as
    uint8 CHW and fed to the model as float32 u8/255 like torchvision ToTensor
    (reference stage_8_torch.py:985-990);
  * float weights are random-init YOLOv8n tensors that the *reference's own* stage_1..7
    pipeline then BN-fuses, calibrates and quantises (oracle/ref_harness.py).
This module holds no arithmetic of the hot path itself.
"""
import numpy as np

N_BN = 4             # images in the BN data-dependent init pass
N_CALIB = 6          # calibration images fed to the reference stage_4
H = W = 640

# class-branch tuning (SURVEY hard part 8): chosen so that per-image candidate counts of the
# golden image set span {0, 1..999, >=1000}; see oracle/ref_harness.py --tune
CLS_BIAS = -4.0
CLS_GAIN = 2.0
CONV_GAIN = 1.0     # extra gain on He-normal conv weights so deep features stay image-dependent


def synth_image_u8(seed):
    """uint8 (3,640,640). Four families by seed % 4: multi-scale blocks, white noise,
    rectangles on flat background, low-contrast blocks (max < 255 -> per-image scale != 127)."""
    rng = np.random.Generator(np.random.PCG64(int(seed) + 7919))
    fam = int(seed) % 4
    if fam == 1:
        img = rng.integers(0, 256, size=(3, H, W), dtype=np.int32)
    elif fam == 2:
        img = np.empty((3, H, W), np.int32)
        img[:] = rng.integers(0, 256, size=(3, 1, 1), dtype=np.int32)
        for _ in range(int(rng.integers(3, 12))):
            y0, x0 = (int(v) for v in rng.integers(0, H - 40, size=2))
            hh, ww = (int(v) for v in rng.integers(24, 320, size=2))
            col = rng.integers(0, 256, size=(3, 1, 1), dtype=np.int32)
            img[:, y0:y0 + hh, x0:x0 + ww] = col
        img += rng.integers(-6, 7, size=(3, H, W), dtype=np.int32)
        img = np.clip(img, 0, 255)
    else:
        acc = np.zeros((3, H, W), np.int32)
        tot = 0
        for cell, wgt in ((160, 4), (32, 3), (8, 2), (2, 1)):
            n = H // cell
            g = rng.integers(0, 256, size=(3, n, n), dtype=np.int32)
            acc += wgt * np.repeat(np.repeat(g, cell, axis=1), cell, axis=2)
            tot += wgt
        img = acc // tot
        if fam == 3:
            gain = int(rng.integers(40, 200))
            img = (img * gain) >> 8
    return img.astype(np.uint8)


def to_input_tensor(img_u8):
    """(3,H,W) uint8 -> torch float32 (1,3,H,W) in [0,1]  (ToTensor semantics: u8 / 255)."""
    import torch
    return (torch.from_numpy(img_u8.astype(np.float32)) / 255.0).unsqueeze(0).contiguous()


def to_input_array(imgs_u8):
    """list/array of (3,H,W) uint8 -> numpy float32 (N,3,H,W) = u8/255 (same rounding as torch)."""
    a = np.asarray(imgs_u8, dtype=np.uint8).astype(np.float32)
    return (a / np.float32(255.0)).astype(np.float32)


def synth_float_weights(sd, model=None):
    """Re-randomise a stage_1-architecture state_dict (reference stage_1.py:40-765): He-normal conv
    weights, non-trivial BN statistics, dfl.weight = arange(16) like ultralytics, class-branch final
    conv tuned by CLS_BIAS / CLS_GAIN.  Order and shapes are preserved (stage_1 maps by position).
    If `model` (the stage_1 nn.Module) is given, BN running statistics are then set from one
    train-mode pass over N_BN synthetic images (data-dependent init), so that deep features stay
    image-dependent instead of collapsing onto the biases."""
    import torch
    gen = torch.Generator().manual_seed(0)
    out = type(sd)()
    keys = list(sd.keys())
    for name in keys:
        t = sd[name]
        if name == 'dfl.weight':
            v = torch.arange(16, dtype=torch.float32).reshape(1, 16, 1, 1)
        elif name.endswith('num_batches_tracked'):
            v = t.clone()
        elif name.endswith('running_mean'):
            v = torch.randn(t.shape, generator=gen) * 0.1
        elif name.endswith('running_var'):
            v = torch.rand(t.shape, generator=gen) * 0.5 + 0.75
        elif t.dim() == 4:
            fan_in = t.shape[1] * t.shape[2] * t.shape[3]
            v = torch.randn(t.shape, generator=gen) * (2.0 / fan_in) ** 0.5 * CONV_GAIN
            if 'down' in name and name.split('.')[1] == '6':      # final class conv (Sequential idx 6 in stage_1)
                v = v * CLS_GAIN
        elif t.dim() == 1 and name.endswith('weight'):            # BN gamma
            v = torch.rand(t.shape, generator=gen) * 0.5 + 0.75
        elif t.dim() == 1 and name.endswith('bias'):
            v = torch.randn(t.shape, generator=gen) * 0.1
            if 'down' in name and name.split('.')[1] == '6':
                v = v + CLS_BIAS
        else:
            v = t.clone()
        out[name] = v.to(t.dtype)
    if model is None:
        return out
    model.load_state_dict(out)
    bns = [m for m in model.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    for m in bns:
        m.momentum = 1.0
    model.train()
    x = torch.cat([to_input_tensor(synth_image_u8(2000 + i)) for i in range(N_BN)])
    with torch.no_grad():
        try:
            model(x)
        except Exception:      # the reference head is batch-1 only (.view(1,64,-1)); all BNs ran before it
            pass
    model.eval()
    for m in bns:
        m.momentum = 0.03
    return type(sd)((k, v.clone()) for k, v in model.state_dict().items())
