"""Calibration max-reduction (stage_4 / stage_5 of the reference) on the GPU.

Reference behaviour (paths relative to /root/reference/quantisation/):
  * utils/save_a.py:11-26  save_max_a(maxim_a, matr, layer): appends abs(matr).max() of one float tap tensor to
    maxim_a[layer] (one entry per calibration image);
  * stage_4.py:1007-1011   writes results/max_a_all.txt as  `name: [tensor(v), ...]`  (str() of fp32 tensors = 4 decimals);
  * stage_5.py:11-33 + utils/stage_5_common_func.py:11-26 (mode 'max')  parses that text back and writes
    results/max_a.txt  (`start: 1.0` first, then `name: max`), which is what the stage_8 hot path reads (utils/max_a.py).
The abs-max itself runs in libayq.so (ayq_absmax_f32, one launch per tap tensor, any batch size); there is no CPU path.
"""
import torch

from . import engine as _eng


def absmax_per_image(t):
    """t: CUDA float32 (n, ...) -> CUDA float32 (n): max|t[i]| per image."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _eng.AyqError('absmax_per_image: needs a CUDA tensor (this library has no CPU path)')
    lib = _eng.load_library()
    x = t.contiguous().to(torch.float32)
    n = x.shape[0]
    per = x.numel() // max(n, 1)
    out = torch.empty((n,), dtype=torch.float32, device=x.device)
    _eng.check(lib.ayq_absmax_f32(x.data_ptr(), out.data_ptr(), n, per, _eng._stream_ptr(x.device)))
    return out


def save_max_a(maxim_a, matr, layer):
    """Drop-in for utils/save_a.py:11-26.  The reference is called with batch-1 tensors and appends ONE value (the max
    over the whole tensor); a batched tensor appends one value per image, i.e. what n batch-1 calls would append."""
    vals = absmax_per_image(matr)
    lst = maxim_a.setdefault(layer, [])
    if matr.shape[0] == 1:
        lst.append(vals[0])
    else:
        lst.extend(vals.unbind(0))


def format_max_a_all(maxim_a):
    """stage_4.py:1007-1011: one line per tap, `name: [tensor(1.4271), ...]` (python str() of 0-dim fp32 tensors)."""
    lines = []
    for key, value in maxim_a.items():
        value = [v.detach().cpu() if isinstance(v, torch.Tensor) else torch.tensor(float(v)) for v in value]
        lines.append(f'{key}: {value}\n')
    return ''.join(lines)


def parse_max_a_all(text):
    """stage_5.py:11-27: {name: [float, ...]} from the max_a_all.txt text (values carry 4 decimals)."""
    out = {}
    for el in text.splitlines():
        if not el.strip():
            continue
        key, value = el.split(': ', 1)
        value = value.replace('[', '').replace(']', '')
        vals = []
        for tnsr in value.split(', '):
            tnsr = tnsr.replace('tensor(', '').replace(')', '')
            if "device='cuda:0'" not in tnsr:
                vals.append(float(tnsr))
        out[key] = vals
    return out


def format_max_a(max_a_all):
    """utils/stage_5_common_func.py:11-26 with MAX_ACTIVATIONS_MODE = 'max': first line `start: 1.0`, then the maximum
    of |v| per tap, SKIPPING the first column exactly like `list(df)[1:]` does."""
    lines = ['start: 1.0\n']
    for name in list(max_a_all)[1:]:
        lines.append(f'{name}: {max(abs(v) for v in max_a_all[name])}\n')
    return ''.join(lines)
