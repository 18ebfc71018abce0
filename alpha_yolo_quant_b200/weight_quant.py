"""Weight quantiser on the GPU: drop-in for conv_quant() of /root/reference/quantisation/stage_6_full_quant.py:89-126
 -- the step that turns the BN-fused float weights of stage_2 into the integer weights, integer
biases and per-channel scales that stage_7 formats and the stage_8 hot path loads.  The reference spends minutes per model
here on Verilog text dumps and a numpy convolution of the calibration image; the arithmetic itself is a per-channel
abs-max, a scale and a rounding, done here by quant_weights_kernel behind ayq_quant_weights_f32 (include/ayq.h).

Bit-exact against the reference executed under numpy >= 2 (tests/golden/golden_wquant_k8.npz, tests/test_gpu_weight_quant.py).
No CPU path: non-CUDA tensors raise.
"""
import numpy as np
import torch

from . import engine as _eng
from .lut import scale  # noqa: F401  (reference name)


def conv_quant(layer_name, conv, bias_conv, scale_input=0, start=False, k=8):
    """conv: CUDA float32 (C, Cin, kh, kw); bias_conv: CUDA float32 (C, 1, 1, 1) or (C,); scale_input: python float (ignored
    when start=True, like the reference).  Returns (weights int64 (C, Cin, kh, kw), bias int64 (1, C, 1, 1), scale_res
    float64 (1, C, 1, 1)) as CUDA tensors -- the three arrays the reference pickles into weights_pickle/ and bias_scales/."""
    if not (isinstance(conv, torch.Tensor) and conv.is_cuda and isinstance(bias_conv, torch.Tensor) and bias_conv.is_cuda):
        raise _eng.AyqError(f'conv_quant({layer_name}): needs CUDA tensors (this library has no CPU path)')
    lib = _eng.load_library()
    w = conv.contiguous().to(torch.float32)
    b = bias_conv.contiguous().to(torch.float32).reshape(-1)
    c = w.shape[0]
    if b.numel() != c:
        raise _eng.AyqError(f'conv_quant({layer_name}): {b.numel()} biases for {c} output channels')
    per = w.numel() // max(c, 1)
    qw = torch.empty(w.shape, dtype=torch.int8, device=w.device)
    qb = torch.empty((c,), dtype=torch.int64, device=w.device)
    sr = torch.empty((c,), dtype=torch.float64, device=w.device)
    si = float(2 ** (k - 1) - 1) if start else float(scale_input)
    _eng.check(lib.ayq_quant_weights_f32(w.data_ptr(), b.data_ptr(), c, per, int(k), si, qw.data_ptr(), qb.data_ptr(),
                                         sr.data_ptr(), _eng._stream_ptr(w.device)))
    return qw.to(torch.int64), qb.reshape(1, c, 1, 1), sr.reshape(1, c, 1, 1)
