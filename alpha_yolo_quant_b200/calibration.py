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


# ----------------------------------------------------------------------------- calibration forward (stage_4.py:475-946)
# state_dict prefix, reference tap name, stride of every convolution in forward order; C2f / SPPF / neck wiring below.
_STRIDE2 = ('conv0.0', 'conv1.0', 'conv3.0', 'conv5.0', 'conv7.0', 'conv8.0', 'conv9.0')


class CalibrationModel:
    """The BN-fused float YOLOv8n of stage_4 (weights_batchnf.pickle, the stage_2 output) on the GPU with the reference's
    64 abs-max taps: the input ('start') and every convolution output before its SiLU (stage_4.py:477-909).  Convolution
    (with the tap fused into its epilogue), SiLU, the SPPF pools and the upsampling are libayq.so kernels
    (ayq_calib_*); concatenation / channel split / residual add are torch tensor plumbing.  Batched: a tensor of n images
    appends n values per tap, exactly what n iterations of the reference loop (:978-983) append.

        m = CalibrationModel(torch.load('8_nano/results/weights_batchnf.pickle'), device='cuda')
        maxim_a = {}
        m.forward(images, maxim_a)                    # images (n,3,640,640) float32 in [0,1]
        open('8_nano/results/max_a_all.txt', 'w').write(format_max_a_all(maxim_a))
        open('8_nano/results/max_a.txt', 'w').write(format_max_a(parse_max_a_all(format_max_a_all(maxim_a))))
    """

    def __init__(self, fused_state_dict, device='cuda'):
        self.lib = _eng.load_library()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _eng.AyqError('CalibrationModel: needs a CUDA device (this library has no CPU path)')
        self.sd = {k: torch.as_tensor(v).to(torch.float32).to(self.device).contiguous() for k, v in fused_state_dict.items()}

    # -- kernels
    def _conv(self, x, prefix, tap, maxim_a, silu=True):
        w, b = self.sd[prefix + '.weight'], self.sd[prefix + '.bias']
        n, cin, h, wd = x.shape
        cout, ks = w.shape[0], w.shape[2]
        stride = 2 if prefix in _STRIDE2 else 1
        ho, wo = (h + 2 * (ks // 2) - ks) // stride + 1, (wd + 2 * (ks // 2) - ks) // stride + 1
        x = x.contiguous()
        y = torch.empty((n, cout, ho, wo), dtype=torch.float32, device=x.device)
        amax = torch.zeros((n,), dtype=torch.float32, device=x.device)
        st = _eng._stream_ptr(x.device)
        _eng.check(self.lib.ayq_calib_conv_f32(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), amax.data_ptr(), n, cin, h, wd,
                                               cout, ks, stride, st))
        maxim_a.setdefault(tap, []).extend(amax.unbind(0))                   # save_max_a(maxim_a, x, tap)
        if silu:
            _eng.check(self.lib.ayq_calib_silu_f32(y.data_ptr(), y.numel(), st))
        return y

    def _pool(self, x):
        y = torch.empty_like(x)
        _eng.check(self.lib.ayq_calib_maxpool5_f32(x.data_ptr(), y.data_ptr(), x.shape[0] * x.shape[1], x.shape[2], x.shape[3], _eng._stream_ptr(x.device)))
        return y

    def _up(self, x):
        x = x.contiguous()
        y = torch.empty((x.shape[0], x.shape[1], 2 * x.shape[2], 2 * x.shape[3]), dtype=torch.float32, device=x.device)
        _eng.check(self.lib.ayq_calib_upsample2_f32(x.data_ptr(), y.data_ptr(), x.shape[0] * x.shape[1], x.shape[2], x.shape[3], _eng._stream_ptr(x.device)))
        return y

    def _c2f(self, x, conv0, bottles, conv1, taps, add, maxim_a):
        x = self._conv(x, conv0, taps[0], maxim_a)
        half = x.shape[1] // 2
        parts = [x[:, :half], x[:, half:]]
        cur = x[:, half:]
        for i, bt in enumerate(bottles):
            y = self._conv(cur, bt + '.0', taps[1 + 2 * i], maxim_a)
            y = self._conv(y, bt + '.2', taps[2 + 2 * i], maxim_a)
            cur = y + cur if add else y                                      # x += x_bottle (:520)
            parts.append(cur)
        return self._conv(torch.cat(parts, 1), conv1, taps[-1], maxim_a)

    def forward(self, x, maxim_a):
        """x: (n,3,H,W) float32 CUDA (or host: moved like the reference's img.to(device)).  Appends to maxim_a in place."""
        x = x.to(self.device, torch.float32).contiguous()
        save_max_a(maxim_a, x, 'start')                                      # :477
        with torch.cuda.device(self.device):
            c = lambda t, p, tap, **kw: self._conv(t, p, tap, maxim_a, **kw)
            x = c(x, 'conv0.0', 'conv_p1')
            x = c(x, 'conv1.0', 'conv_p2')
            x = self._c2f(x, 'cf2_conv_0.0', ['cf2_bottle_0'], 'cf2_conv_1.0', ['conv_0_c2f', 'conv_b_0_c2f', 'conv_b_1_c2f', 'conv_b_2_c2f'], True, maxim_a)
            x = c(x, 'conv3.0', 'conv_p3')
            r1 = x = self._c2f(x, 'cf2_conv_2.0', ['cf2_bottle_2', 'cf2_bottle_3'], 'cf2_conv_3.0',
                               ['conv_2_c2f', 'conv_b1_c2f', 'conv_b2_c2f', 'conv_b3_c2f', 'conv_b4_c2f', 'conv_b5_c2f'], True, maxim_a)
            x = c(x, 'conv5.0', 'conv_5')
            r2 = x = self._c2f(x, 'cf2_conv_4.0', ['cf2_bottle_4', 'cf2_bottle_5'], 'cf2_conv_5.0',
                               ['cf2_conv_4', 'cf2_bconv_4', 'cf2_bconv1_4', 'cf2_bconv_5', 'cf2_bconv1_5', 'cf2_6_conv_last'], True, maxim_a)
            x = c(x, 'conv7.0', 'conv7')
            x = self._c2f(x, 'cf2_conv_6.0', ['cf2_bottle_6'], 'cf2_conv_7.0', ['cf2_conv_6', 'cf2_bottle_6', 'cf2_bottle_61', 'cf2_conv_7'], True, maxim_a)
            x = c(x, 'sppf_conv_1.0', 'sppf_conv_1')
            p1 = self._pool(x); p2 = self._pool(p1); p3 = self._pool(p2)
            sppf = x = c(torch.cat((x, p1, p2, p3), 1), 'sppf_conv_2.0', 'sppf_conv_2')
            r4 = x = self._c2f(torch.cat((self._up(x), r2), 1), 'cf2_conv_8.0', ['cf2_bottle_7'], 'cf2_conv_9.0',
                               ['cf2_conv_8', 'cf2_conv_80', 'cf2_conv_81', 'cf2_conv_9'], False, maxim_a)
            r5 = x = self._c2f(torch.cat((self._up(x), r1), 1), 'cf2_conv_10.0', ['cf2_bottle_8'], 'cf2_conv_11.0',
                               ['cf2_conv_10', 'cf2_bottle_8', 'cf2_bottle_81', 'cf2_conv_11'], False, maxim_a)
            x = c(x, 'conv8.0', 'conv8')
            r6 = x = self._c2f(torch.cat((x, r4), 1), 'cf2_conv_12.0', ['cf2_bottle_9'], 'cf2_conv_13.0',
                               ['cf2_conv_12', 'cf2_bottle_9', 'cf2_bottle_90', 'cf2_conv_13'], False, maxim_a)
            x = c(x, 'conv9.0', 'conv9')
            r7 = self._c2f(torch.cat((x, sppf), 1), 'cf2_conv_14.0', ['cf2_bottle_10'], 'cf2_conv_15.0',
                           ['cf2_conv_14', 'cf2_bottle_10', 'cf2_bottle_101', 'cf2_conv_15'], False, maxim_a)
            for feat, det, nm in ((r5, 'detect_5', 'x_result_5'), (r6, 'detect_6', 'x_result_6'), (r7, 'detect_x', 'x')):
                for br in ('up', 'down'):
                    t = c(feat, f'{det}_{br}.0', f'{nm}_{br}_0')
                    t = c(t, f'{det}_{br}.2', f'{nm}_{br}_1')
                    c(t, f'{det}_{br}.4', f'{nm}_{br}_2', silu=False)
        return maxim_a
