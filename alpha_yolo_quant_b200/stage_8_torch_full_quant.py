"""Drop-in for the names that /root/reference/quantisation/stage_8_torch_full_quant.py defines and its
driver code uses, backed by libayq.so (hand-written sm_100a CUDA behind include/ayq.h).

Reference usage (stage_8_torch_full_quant.py:1278-1294)          this module
    model = Yolov8().to(device)                                     same
    model.load_state_dict(torch.load(QUANT_WEIGHTS))                same (127 keys, fp32 tensors holding ints)
    boxes, classes = model(img)        # img (1,3,640,640)          same; (None, None) when nothing passes
                                                                    + model.forward_batch(x) -> list of those pairs
Module globals all_scales / max_a_dict / lookup / lookup_final / lookup_exp / device / K are filled by
configure() (the reference fills them at import from cwd-relative files, :432-436; configure(main_dir=...)
reads the same files, configure(workload=npz) reads this repo's fixture format).

Free functions keep the reference signatures and operate on CUDA fp32 tensors that carry integers.
Nothing here falls back to the CPU: a non-CUDA tensor or a missing libayq.so raises.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import engine as _eng
from . import loaders as _loaders
from . import lut as _lut
from . import plan as _plan
from .lut import scale, create_sigmoid_lookup_table, create_exponent_lookup_table  # noqa: F401  (reference names)
from .loaders import load_scales, max_a  # noqa: F401

K = 8                                    # stage_0.py:7
MAIN_DIR_NAME = f'{K}_nano'              # stage_0.py:14
SIGMOID_RANGE = 6                        # :434 (stage_8_torch.py uses 7, :264)
device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')       # :37
all_scales = {}
max_a_dict = {}
lookup = {}
lookup_final = {}
lookup_exp = {}
_state = {'configured': False}


def configure(main_dir=None, k=8, workload=None, sigmoid_range=6):
    """Fill the module globals the reference builds at import (:432-436)."""
    global K, MAIN_DIR_NAME, SIGMOID_RANGE, all_scales, max_a_dict, lookup, lookup_final, lookup_exp
    sd = None
    if workload is not None:
        K, sd, all_scales, max_a_dict = _loaders.load_workload_npz(workload)
    else:
        K = int(k)
        MAIN_DIR_NAME = main_dir if main_dir is not None else f'{K}_nano'
        all_scales = load_scales(MAIN_DIR_NAME)
        max_a_dict = max_a(f'{MAIN_DIR_NAME}/results/max_a.txt')
    SIGMOID_RANGE = sigmoid_range
    lookup = create_sigmoid_lookup_table(sigmoid_range, K)
    lookup_final = create_sigmoid_lookup_table(12, 16)
    lookup_exp = create_exponent_lookup_table(_plan.DFL_RANGE, K)
    _state['configured'] = True
    return sd


# ----------------------------------------------------------------------------- quantised layer library
def _dev_tensor(arr, dev):
    return torch.as_tensor(np.ascontiguousarray(arr), dtype=torch.float32).to(dev)


def _need_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _eng.AyqError(f'{name}: needs a CUDA tensor (this library has no CPU path)')


def requantize(arr_q_input, old_scale, new_scale, bit_size, device, bit_size_for_koeff=8):
    """utils/rescale_coeff_torch.py:14-46.  Returns (arr_q float32, rescale_koeff, shift_val) like the reference;
    a non-positive scale zeroes the input in place (:34-37); a coefficient > 255 after the retry raises
    RescaleOverflow where the reference print()s and exit()s (:31-33)."""
    _need_cuda(arr_q_input, 'requantize')
    lib = _eng.load_library()
    old_pass = old_scale if isinstance(old_scale, float) else old_scale.shape[1]
    new_pass = new_scale if isinstance(new_scale, float) else new_scale.shape[1]
    x = arr_q_input.contiguous()
    if not (old_pass > 0 and new_pass > 0):
        arr_q_input[...] = 0
        return torch.zeros_like(x, dtype=torch.float32), 0, 0
    k, s = _plan.rescale_coeffs(old_scale, new_scale, bit_size_for_koeff)
    per_channel = 1 if k.numel() > 1 else 0
    n = x.shape[0] if x.dim() == 4 else 1
    c = x.shape[1] if x.dim() == 4 else 1
    hw = x.numel() // max(n * c, 1)
    if per_channel and k.numel() != c:
        raise _eng.AyqError(f'requantize: {k.numel()} per-channel scales for a tensor with {c} channels')
    kd = k.to(torch.float32).to(x.device)
    inv = torch.as_tensor(np.ldexp(1.0, -s.numpy().astype(np.int64)), dtype=torch.float32).to(x.device)
    xf = x.to(torch.float32)
    y = torch.empty_like(xf)
    _eng.check(lib.ayq_requantize_f32(xf.data_ptr(), y.data_ptr(), kd.data_ptr(), inv.data_ptr(), per_channel,
                                      n, c, hw, bit_size, _eng._stream_ptr(x.device)))
    shape = old_scale.shape if isinstance(old_scale, torch.Tensor) and old_scale.dim() == 4 else ()
    return y, k.reshape(shape).to(x.device), s.reshape(shape).to(x.device)


def _lut_apply(x, table, name):
    _need_cuda(x, name)
    lib = _eng.load_library()
    key_min, arr = _lut.table_to_array(table)
    xf = x.contiguous().to(torch.float32)
    y = torch.empty_like(xf)
    lut_d = _dev_tensor(arr, x.device)
    _eng.check(lib.ayq_lut_f32(xf.data_ptr(), y.data_ptr(), lut_d.data_ptr(), key_min, key_min + len(arr) - 1,
                               xf.numel(), _eng._stream_ptr(x.device)))
    return y


def sigmoid_quant(x, lookup, device):
    """utils/silu_torch.py:4-18: table lookup, keys missing from the table give 0."""
    return _lut_apply(x, lookup, 'sigmoid_quant')


def exponent_quant(x, lookup, device):
    """utils/exp_torch.py:4-18"""
    return _lut_apply(x, lookup, 'exponent_quant')


def quant_matrix(matrix, k, start=False):
    """utils/quant_matrix_torch.py:57-70: per-image max-abs quantiser.  Returns (res_matrix, all_scales (n,1))."""
    _need_cuda(matrix, 'quant_matrix')
    lib = _eng.load_library()
    x = matrix.contiguous().to(torch.float32)
    n = x.shape[0]
    per = x.numel() // max(n, 1)
    y = torch.empty_like(x)
    amax = torch.empty((n,), dtype=torch.float32, device=x.device)
    scales = torch.empty((n,), dtype=torch.float32, device=x.device)
    if start:
        raise _eng.AyqError('quant_matrix(start=True) is not on the stage_8 path')
    _eng.check(lib.ayq_quant_input_f32(x.data_ptr(), y.data_ptr(), amax.data_ptr(), scales.data_ptr(), n, per, k,
                                       _eng._stream_ptr(x.device)))
    return y, scales.reshape(n, 1)


def silu(x, scale_x, a_input):
    """stage_8_torch_full_quant.py:439-452: fixed-point SiLU + requantise to scale(a_input, K).
    x: conv accumulators (n,C,H,W) fp32; scale_x: fp32 (1,C,1,1) from all_scales.  Returns (tensor, new scale)."""
    return _silu_impl(x, scale_x, a_input, SIGMOID_RANGE, K, lookup)


def _silu_impl(x, scale_x, a_input, SIGMOID_RANGE, K, lookup):
    """shared by stage_8_torch.silu (sigmoid range 7, stage_8_torch.py:261-276) and the full-quant silu (range 6)"""
    _need_cuda(x, 'silu')
    lib = _eng.load_library()
    xf = x.contiguous().to(torch.float32)
    n, c = xf.shape[0], xf.shape[1]
    hw = xf.numel() // (n * c)
    sx = scale_x.reshape(-1).to(torch.float32).cpu()
    k1, i1 = _plan._k_inv(sx, scale(SIGMOID_RANGE, K))
    new_scale = scale(a_input, K)
    k2, i2 = _plan._k_inv(scale(1, K) * sx, new_scale)
    tab = _dev_tensor(np.stack([k1, i1, k2, i2]), x.device)
    _, arr = _lut.table_to_array(lookup)
    lut_d = _dev_tensor(arr, x.device)
    y = torch.empty_like(xf)
    _eng.check(lib.ayq_silu_f32(xf.data_ptr(), y.data_ptr(), tab.data_ptr(), lut_d.data_ptr(), n, c, hw, K,
                                _eng._stream_ptr(x.device)))
    return y, new_scale


def requant_last_layers(input_tensor, input_scale, k=16):
    """:472-476"""
    input_tensor, rescale, shift = requantize(input_tensor, input_scale, scale(_plan.DFL_RANGE, k), k, device)
    return input_tensor, scale(_plan.DFL_RANGE, k)


def nms_quant(dets, scores, thresh):
    """:248-294 (thresh is unused there too).  Returns the kept indices as a float tensor, stable tie-break."""
    _need_cuda(dets, 'nms_quant')
    lib = _eng.load_library()
    b = dets.contiguous().to(torch.float32)
    s = scores.contiguous().to(torch.float32)
    nb = b.shape[0]
    if nb > 16384:
        raise _eng.AyqError('nms_quant: at most 16384 boxes')
    if nb and (bool((s != s.round()).any()) or float(s.min()) < 0 or float(s.max()) > 131071):
        raise _eng.AyqError('nms_quant: scores must be integers in [0, 131071] (the reference feeds 16-bit LUT scores)')
    keep = torch.empty((1000,), dtype=torch.float32, device=b.device)
    cnt = torch.zeros((1,), dtype=torch.int32, device=b.device)
    _eng.check(lib.ayq_nms_boxes(b.data_ptr(), s.data_ptr(), nb, keep.data_ptr(), cnt.data_ptr(), _eng._stream_ptr(b.device)))
    return keep[:int(cnt.item())]


_nms_engine = {}


def _engine_for_nms(dev):
    """coord_quant only needs the NMS kernels; reuse any configured model's engine on that device."""
    e = _nms_engine.get(dev.index)
    if e is None:
        raise _eng.AyqError('coord_quant: create a Yolov8 (load_state_dict + .to(cuda)) on this device first')
    return e


def coord_quant(prediction):
    """:297-361: prediction (1,84,8400) -> [tensor (n,6)] or None.  Like the reference, only the first image is
    returned (the reference `return`s inside its loop, :361)."""
    _need_cuda(prediction, 'coord_quant')
    e = _engine_for_nms(prediction.device)
    dets, counts = e.nms(prediction[:1])
    k = int(counts[0].item())
    if k == 0:
        return None
    # rows before scale_boxes/clip_boxes: the kernel already clamps to [0,640], which clip_boxes would do next
    return [dets[0, :k].clone()]


# ----------------------------------------------------------------------------- model
def _conv_shapes():
    """(state_dict prefix, cout, cin, k) for the 63 convs, from stage_8_torch_full_quant.py:489-694 (W=0.25, D=0.33)."""
    def c2f(prefix_conv0, bottles, prefix_conv1, cin, cout, n, neck_in=None):
        out = [(prefix_conv0, cout, neck_in or cin, 1)]
        for b in bottles:
            out += [(b + '.0', cout // 2, cout // 2, 3), (b + '.2', cout // 2, cout // 2, 3)]
        out.append((prefix_conv1, cout, (2 + n) * cout // 2, 1))
        return out
    L = [('conv0.0', 16, 3, 3), ('conv1.0', 32, 16, 3)]
    L += c2f('cf2_conv_0.0', ['cf2_bottle_0'], 'cf2_conv_1.0', 32, 32, 1)
    L += [('conv3.0', 64, 32, 3)]
    L += c2f('cf2_conv_2.0', ['cf2_bottle_2', 'cf2_bottle_3'], 'cf2_conv_3.0', 64, 64, 2)
    L += [('conv5.0', 128, 64, 3)]
    L += c2f('cf2_conv_4.0', ['cf2_bottle_4', 'cf2_bottle_5'], 'cf2_conv_5.0', 128, 128, 2)
    L += [('conv7.0', 256, 128, 3)]
    L += c2f('cf2_conv_6.0', ['cf2_bottle_6'], 'cf2_conv_7.0', 256, 256, 1)
    L += [('sppf_conv_1.0', 128, 256, 1), ('sppf_conv_2.0', 256, 512, 1)]
    L += c2f('cf2_conv_8.0', ['cf2_bottle_7'], 'cf2_conv_9.0', 128, 128, 1, neck_in=384)
    L += c2f('cf2_conv_10.0', ['cf2_bottle_8'], 'cf2_conv_11.0', 64, 64, 1, neck_in=192)
    L += [('conv8.0', 64, 64, 3)]
    L += c2f('cf2_conv_12.0', ['cf2_bottle_9'], 'cf2_conv_13.0', 128, 128, 1, neck_in=192)
    L += [('conv9.0', 128, 128, 3)]
    L += c2f('cf2_conv_14.0', ['cf2_bottle_10'], 'cf2_conv_15.0', 256, 256, 1, neck_in=384)
    for name, cin in (('detect_5', 64), ('detect_6', 128), ('detect_x', 256)):
        L += [(f'{name}_up.0', 64, cin, 3), (f'{name}_up.2', 64, 64, 3), (f'{name}_up.4', 64, 64, 1)]
        L += [(f'{name}_down.0', 80, cin, 3), (f'{name}_down.2', 80, 80, 3), (f'{name}_down.4', 80, 80, 1)]
    return L


class Yolov8(nn.Module):
    """Quantised YOLOv8n with the reference's state_dict layout (Appendix D: 63 x {weight,bias} + dfl.weight).
    The tensors are only the container load_state_dict() fills; forward() runs the compiled CUDA plan."""

    def __init__(self, max_batch=512, taps=False):
        super().__init__()
        for prefix, cout, cin, k in _conv_shapes():
            seq, idx = prefix.split('.')
            if not hasattr(self, seq):
                setattr(self, seq, nn.Module())
            holder = nn.Module()
            holder.weight = nn.Parameter(torch.zeros(cout, cin, k, k), requires_grad=False)
            holder.bias = nn.Parameter(torch.zeros(cout), requires_grad=False)
            getattr(self, seq).add_module(idx, holder)
        self.dfl = nn.Module()
        self.dfl.weight = nn.Parameter(torch.zeros(1, 16, 1, 1), requires_grad=False)
        self._engine = None
        self._plan = None
        self._max_batch = max_batch
        self._taps = taps
        self._scales, self._max_a, self._K, self._sig = None, None, None, None

    # reference order of keys must match Appendix D (tests check against the fixture's sd_keys)
    def load_state_dict(self, state_dict, strict=True, assign=False):
        r = super().load_state_dict(state_dict, strict=strict)
        self._engine = None                    # weights changed: recompile lazily
        return r

    _HEAD = 'int'                             # stage_8_torch.Yolov8 overrides: 'float'

    def _cfg(self):
        """(configured, all_scales, max_a_dict, K, sigmoid_range) of the module this class lives in"""
        return _state['configured'], all_scales, max_a_dict, K, SIGMOID_RANGE

    def _register(self, index):
        _nms_engine[index] = self._engine

    def _ensure_engine(self):
        if self._engine is not None:
            return self._engine
        configured, scales_, max_a_, k_, sig_ = self._cfg()
        if not configured:
            raise _eng.AyqError('call configure(main_dir=... | workload=...) before running the model (all_scales / max_a_dict)')
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            raise _eng.AyqError('Yolov8: move the model to a CUDA device (.to("cuda")); there is no CPU path')
        sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
        self._plan = _plan.compile_plan(sd, scales_, max_a_, k_, sig_, taps=self._taps, head=self._HEAD)
        self._engine = _eng.Engine(self._plan, dev.index or 0, self._max_batch)
        self._register(dev.index or 0)
        return self._engine

    @property
    def engine(self):
        return self._ensure_engine()

    def forward_batch(self, x):
        """x (N,3,640,640) fp32 in [0,1] -> list of (boxes, classes) | (None, None); element i == model(x[i:i+1]).
        uint8 images (the loader's format before ToTensor, :985-990 of stage_8_torch.py) are accepted too: u8 / 255 then
        happens on the GPU, with the same results as on (x / 255).float()."""
        e = self._ensure_engine()
        if not x.is_cuda:
            x = x.to(e.device)                 # the reference does x.to(device) inside forward (:710)
        dets, counts = e.forward(x)
        return _eng.unpack_detections(dets, counts)

    def forward(self, x):
        """:704-1275.  Batch-1 like the reference (.view(1,64,-1) :1158); use forward_batch for N > 1."""
        if x.shape[0] != 1:
            raise _eng.AyqError('Yolov8.forward is batch-1 like the reference; use forward_batch(x) for N > 1')
        return self.forward_batch(x)[0]
