"""CPU oracle -- TEST INFRASTRUCTURE.  A numpy restatement of the reference's integer
YOLOv8n forward + q_NMS (quantisation/stage_8_torch_full_quant.py).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it;
the product (alpha_yolo_quant_b200) never does.

Parity pin: tests/test_oracle_golden.py checks every integer tensor this file produces
against tests/golden/golden_k*.npz, which oracle/ref_harness.py recorded by executing the
UNMODIFIED reference in the build container (the reference ships no golden vectors of its
own).

Every function cites the reference lines it restates (paths relative to
/root/reference/quantisation/).  The arithmetic is SURVEY.md Appendix A: integer
accumulators, fp32-rounded products (RN32), round-half-up shifts.
"""
import numpy as np

F32 = np.float32
DFL_RANGE = 14.8264799118042          # stage_8_torch_full_quant.py:436,473


# ----------------------------------------------------------------------------- helpers
def scale(a, k):
    """utils/scale.py:4-5"""
    return (2 ** (k - 1) - 1) / a


def max_a_from_text(txt):
    """utils/max_a.py:1-7"""
    d = {}
    for line in str(txt).splitlines():
        if not line.strip():
            continue
        d[line.split(' ')[0][:-1]] = float(line.split(' ')[1].rstrip('\n'))
    return d


def _lut_dequant(i, max_val, bits):
    # utils/silu.py:22-29 (numpy-2 dtype rules: float32 array / python float -> float32)
    arr = np.array((i,)).astype(np.float32)
    s = (2 ** (bits - 1) - 1) / max_val
    arr /= s
    return arr[0]


def _lut_quant(v, bits):
    # utils/silu.py:14-19 with max_val = 1
    m = 2 ** (bits - 1) - 1
    q = np.round(v * (m / 1))
    return np.clip(q, -m, m)[0]


def sigmoid_lut(max_conv_value, bits):
    """utils/silu.py:32-50 -> float32 array indexed by (i + M), i in [-M, M].
    Scalar loop on purpose: numpy's scalar float32 pow and its vectorised array pow differ in the
    last ulp for a handful of the 65 535 16-bit entries, and the reference uses scalars."""
    m = 2 ** (bits - 1) - 1
    out = np.empty((2 * m + 1,), np.float32)
    for i in range(-m, m + 1):
        d = _lut_dequant(i, max_conv_value, bits)
        sig = 1 / (1 + (np.e ** (-d)))              # utils/silu.py:4-5 on a float32 scalar
        out[i + m] = _lut_quant(np.array((sig,)), bits)
    return out


def exponent_lut(max_conv_value, bits):
    """utils/exponent.py:32-50 -> float32 array indexed by (i + 2^bits - 1), i in [-(2^bits-1), 0]."""
    top = 2 ** bits - 1
    out = np.empty((top + 1,), np.float32)
    for i in range(-top, 1):
        d = _lut_dequant(i, max_conv_value, bits)
        out[i + top] = _lut_quant(np.array((np.exp(d),)), bits)
    return out


def coeffs(old, new, koeff_bits=8):
    """utils/rescale_coeff_torch.py:20-33.  old: python float or fp32 array (C,); new: python float.
    All arithmetic in fp32 like the torch tensors of the reference.  Returns (k int64[C], s int64[C])."""
    old32 = np.asarray(old, dtype=F32).reshape(-1)
    new32 = F32(new)
    with np.errstate(all='ignore'):
        shift = F32(koeff_bits) + np.floor(np.log2((old32 / new32).astype(F32)).astype(F32))
        k = np.round((F32(2) ** shift).astype(F32) * (new32 / old32).astype(F32))
        if k.max() > 2 ** koeff_bits - 1:
            shift = shift - F32(1)
            k = np.round((F32(2) ** shift).astype(F32) * (new32 / old32).astype(F32))
            if k.max() > 2 ** koeff_bits - 1:
                raise OverflowError('rescale coefficient > %d' % (2 ** koeff_bits - 1))   # reference: print + exit()
    return k.astype(np.int64), shift.astype(np.int64)


def rn32_mul_int(k, x):
    """RN32(k * x) as int64: the fp32 product of two fp32-representable integers (Appendix A)."""
    return (np.asarray(k, F32) * np.asarray(x, F32)).astype(F32).astype(np.int64)


def rsh(t, s):
    """floor((t + 2^(s-1)) / 2^s) == (t // 2^(s-1)) // 2 + (t // 2^(s-1)) % 2  (utils/rescale_coeff_torch.py:42-44)"""
    t = np.asarray(t, np.int64)
    s = np.asarray(s, np.int64)
    pos = np.maximum(s, 1)
    up = (t + (np.int64(1) << (pos - 1))) >> pos
    # s <= 0 (coarse old scale, e.g. K=4 class logits): t // 2^(s-1) is the exact, even product
    # t * 2^(1-s), so the result is t * 2^(-s)
    return np.where(s >= 1, up, t << np.maximum(-s, 0))


def requant_apply(x, k, s, bits):
    """utils/rescale_coeff_torch.py:42-46 with per-channel (k, s) broadcast over NCHW axis 1."""
    m = 2 ** (bits - 1) - 1
    if k.size > 1:
        shp = [1] * x.ndim
        shp[1] = -1
        k = k.reshape(shp)
        s = s.reshape(shp)
    return np.clip(rsh(rn32_mul_int(k, x), s), -m, m)


def requantize(x, old, new, bits):
    k, s = coeffs(old, new)
    return requant_apply(x, k, s, bits), k, s


# ----------------------------------------------------------------------------- layers
def conv2d_int(x, w, b, stride, pad):
    """nn.Conv2d on integer tensors (stage_8_torch_full_quant.py:713 ...): exact integer result.
    x (N,C,H,W) int, w (O,C,kh,kw) int8, b (O,) int64 or None.  float64 GEMM is exact below 2^53."""
    n, c, h, wd = x.shape
    o, _, kh, kw = w.shape
    ho = (h + 2 * pad - kh) // stride + 1
    wo = (wd + 2 * pad - kw) // stride + 1
    xp = np.zeros((n, c, h + 2 * pad, wd + 2 * pad), np.float64)
    xp[:, :, pad:pad + h, pad:pad + wd] = x
    cols = np.empty((n, kh * kw * c, ho, wo), np.float64)
    t = 0
    for ky in range(kh):
        for kx in range(kw):
            cols[:, t * c:(t + 1) * c] = xp[:, :, ky:ky + stride * ho:stride, kx:kx + stride * wo:stride]
            t += 1
    wm = w.transpose(0, 2, 3, 1).reshape(o, kh * kw * c).astype(np.float64)
    y = np.einsum('ok,nkp->nop', wm, cols.reshape(n, kh * kw * c, ho * wo), optimize=True)
    y = np.rint(y).astype(np.int64).reshape(n, o, ho, wo)
    if b is not None:
        y += np.asarray(b, np.int64).reshape(1, -1, 1, 1)
    return y


def maxpool5(x):
    """nn.MaxPool2d(5, 1, padding=2) (stage_8_torch_full_quant.py:572-576): -inf padding."""
    n, c, h, w = x.shape
    big = np.iinfo(np.int64).min
    xp = np.full((n, c, h + 4, w + 4), big, np.int64)
    xp[:, :, 2:2 + h, 2:2 + w] = x
    out = xp[:, :, 0:h, 0:w].copy()
    for dy in range(5):
        for dx in range(5):
            np.maximum(out, xp[:, :, dy:dy + h, dx:dx + w], out=out)
    return out


def upsample2(x):
    """nn.Upsample(None, 2, 'nearest') (stage_8_torch_full_quant.py:588)"""
    return x.repeat(2, axis=2).repeat(2, axis=3)


def quant_input(x, k):
    """utils/quant_matrix_torch.py:57-70: per image a = max|x| (a 0-dim fp32 tensor), scale = scale(a, k) = M / a,
    q = rint(fl32(x*scale)).  `python_int / tensor` is Tensor.__rtruediv__ = reciprocal(a) * M in torch, i.e.
    fl32(fl32(1/a) * M), NOT fl32(M/a): the two differ in the last ulp for some a (golden image 8, a = 244/255).
    Returns int64 (N,3,H,W) and the per-image fp32 scales."""
    x = np.asarray(x, F32)
    out = np.empty(x.shape, np.int64)
    scales = np.empty((x.shape[0],), F32)
    m = 2 ** (k - 1) - 1
    for i in range(x.shape[0]):
        a = np.abs(x[i]).max()
        with np.errstate(all='ignore'):
            s = F32(F32(1) / F32(a)) * F32(m)
            out[i] = np.rint((np.clip(x[i], -a, a) * s).astype(F32)).astype(np.int64)
        scales[i] = s
    return out, scales


class Workload:
    """Artefacts of reference stages 1-7 as exported by oracle/ref_harness.py (Appendix D formats)."""

    def __init__(self, npz_path):
        z = np.load(npz_path, allow_pickle=False)
        self.K = int(z['K'])
        self.sd = {name: z['sd/' + name] for name in z['sd_keys']}
        self.scales = {name: z['scale/' + name].astype(F32) for name in z['scale_keys']}
        self.max_a = max_a_from_text(z['max_a_txt'])


# (all_scales key, state_dict conv prefix) in forward order -- stage_8_torch_full_quant.py:713-1117
_SD = {
    'Conv_P1': 'conv0.0', 'Conv_P2': 'conv1.0', 'C2F_2_conv_0': 'cf2_conv_0.0', 'C2F_2_bottle_0': 'cf2_bottle_0.0',
    'C2F_2_bottle_1': 'cf2_bottle_0.2', 'C2F_2_conv_1': 'cf2_conv_1.0', 'Conv_P3': 'conv3.0',
    'C2F_4_conv_0': 'cf2_conv_2.0', 'C2F_4_bottle_0': 'cf2_bottle_2.0', 'C2F_4_bottle_1': 'cf2_bottle_2.2',
    'C2F_4_bottle_2': 'cf2_bottle_3.0', 'C2F_4_bottle_3': 'cf2_bottle_3.2', 'C2F_4_conv_1': 'cf2_conv_3.0',
    'Conv_P4': 'conv5.0', 'C2F_6_conv_0': 'cf2_conv_4.0', 'C2F_6_bottle_0': 'cf2_bottle_4.0',
    'C2F_6_bottle_1': 'cf2_bottle_4.2', 'C2F_6_bottle_2': 'cf2_bottle_5.0', 'C2F_6_bottle_3': 'cf2_bottle_5.2',
    'C2F_6_conv_1': 'cf2_conv_5.0', 'Conv_P5': 'conv7.0', 'C2F_8_conv_0': 'cf2_conv_6.0',
    'C2F_8_bottle_0': 'cf2_bottle_6.0', 'C2F_8_bottle_1': 'cf2_bottle_6.2', 'C2F_8_conv_1': 'cf2_conv_7.0',
    'SPPF_conv_0': 'sppf_conv_1.0', 'SPPF_conv_1': 'sppf_conv_2.0',
    'C2F_12_conv_0': 'cf2_conv_8.0', 'C2F_12_bottle_0': 'cf2_bottle_7.0', 'C2F_12_bottle_1': 'cf2_bottle_7.2',
    'C2F_12_conv_1': 'cf2_conv_9.0', 'C2F_15_conv_0': 'cf2_conv_10.0', 'C2F_15_bottle_0': 'cf2_bottle_8.0',
    'C2F_15_bottle_1': 'cf2_bottle_8.2', 'C2F_15_conv_1': 'cf2_conv_11.0', 'Conv_16': 'conv8.0',
    'C2F_18_conv_0': 'cf2_conv_12.0', 'C2F_18_bottle_0': 'cf2_bottle_9.0', 'C2F_18_bottle_1': 'cf2_bottle_9.2',
    'C2F_18_conv_1': 'cf2_conv_13.0', 'Conv_19': 'conv9.0', 'C2F_21_conv_0': 'cf2_conv_14.0',
    'C2F_21_bottle_0': 'cf2_bottle_10.0', 'C2F_21_bottle_1': 'cf2_bottle_10.2', 'C2F_21_conv_1': 'cf2_conv_15.0',
    'x_result_5_up_0': 'detect_5_up.0', 'x_result_5_up_1': 'detect_5_up.2', 'x_result_5_up_2': 'detect_5_up.4',
    'x_result_5_down_0': 'detect_5_down.0', 'x_result_5_down_1': 'detect_5_down.2', 'x_result_5_down_2': 'detect_5_down.4',
    'x_result_6_up_0': 'detect_6_up.0', 'x_result_6_up_1': 'detect_6_up.2', 'x_result_6_up_2': 'detect_6_up.4',
    'x_result_6_down_0': 'detect_6_down.0', 'x_result_6_down_1': 'detect_6_down.2', 'x_result_6_down_2': 'detect_6_down.4',
    'x_up_0': 'detect_x_up.0', 'x_up_1': 'detect_x_up.2', 'x_up_2': 'detect_x_up.4',
    'x_down_0': 'detect_x_down.0', 'x_down_1': 'detect_x_down.2', 'x_down_2': 'detect_x_down.4',
}


class OracleYolov8:
    """Sequential restatement of Yolov8.forward (stage_8_torch_full_quant.py:704-1275), batched:
    every image is processed exactly as the reference processes a batch of one."""

    def __init__(self, wl, sigmoid_range=6):
        self.wl = wl
        self.K = wl.K
        self.MK = 2 ** (wl.K - 1) - 1
        self.lut = sigmoid_lut(sigmoid_range, wl.K)           # :434
        self.lut16 = sigmoid_lut(12, 16)                      # :435
        self.lut_exp = exponent_lut(DFL_RANGE, wl.K)          # :436
        self.sigmoid_range = sigmoid_range
        self.trace = None

    # -- per-layer pieces
    def _conv(self, x, key):
        pre = _SD[key]
        w = self.wl.sd[pre + '.weight']
        b = self.wl.sd[pre + '.bias']
        kk = w.shape[2]
        stride = 2 if key in ('Conv_P1', 'Conv_P2', 'Conv_P3', 'Conv_P4', 'Conv_P5', 'Conv_16', 'Conv_19') else 1
        y = conv2d_int(x, w, b, stride, 1 if kk == 3 else 0)
        if self.trace is not None:
            self.trace['conv'].append(y)
        return y

    def _silu(self, acc, key, next_a):
        """silu() :439-452.  Returns (int tensor, python-float scale of the result)."""
        K, MK = self.K, self.MK
        sx = self.wl.scales[key]                              # fp32 (C,)
        k1, s1 = coeffs(sx, scale(self.sigmoid_range, K))
        r1 = requant_apply(acc, k1, s1, K)
        sig = self.lut[r1 + MK]                               # sigmoid_quant, utils/silu_torch.py:4-18
        pr = (sig * acc.astype(F32)).astype(F32)              # res_silu *= res_conv_copy   (fp32)
        pr = np.rint(pr)
        scale_silu = (F32(scale(1, K)) * sx).astype(F32)      # :448
        new = scale(self.wl.max_a[next_a], K)                 # :450
        k2, s2 = coeffs(scale_silu, new)
        out = requant_apply(pr, k2, s2, K)
        if self.trace is not None:
            self.trace['silu'].append(out)
            self.trace['coeff'] += [(k1, s1), (k2, s2)]
        return out, new

    def _cs(self, x, key, next_a):
        return self._silu(self._conv(x, key), key, next_a)

    def _requant(self, x, old, new, bits=None):
        q, k, s = requantize(x, old, new, self.K if bits is None else bits)
        if self.trace is not None:
            self.trace['requant'].append(q)
            self.trace['coeff'].append((k, s))
        return q

    def _c2f(self, x, name, a_keys, n_bottle, add):
        """C2f block, e.g. :723-748 (backbone, add=True) and :908-929 (neck, add=False)."""
        x, s0 = self._cs(x, f'{name}_conv_0', a_keys[0])
        half = x.shape[1] // 2
        parts = [x[:, :half], x[:, half:]]
        cur = x[:, half:]
        for i in range(n_bottle):
            y, _ = self._cs(cur, f'{name}_bottle_{2 * i}', a_keys[1 + 2 * i])
            y, sy = self._cs(y, f'{name}_bottle_{2 * i + 1}', a_keys[2 + 2 * i])
            y = self._requant(y, sy, s0)
            cur = y + cur if add else y                       # x += x_bottle_k  (unclipped, :742)
            parts.append(cur)
        return self._cs(np.concatenate(parts, 1), f'{name}_conv_1', a_keys[-1])

    # -- whole forward
    def forward_maps(self, img, trace=False, raw_head=False):
        """img float32 (N,3,640,640) in [0,1].  Returns the three box maps and three class maps of the Detect head
        (requantised, stage_8_torch_full_quant.py), or with raw_head=True the six raw conv accumulators with their
        per-channel scales, which is what stage_8_torch.py:915-922 dequantises (oracle/float_head.py)."""
        K = self.K
        self.trace = {'conv': [], 'silu': [], 'requant': [], 'coeff': []} if trace else None
        x, _ = quant_input(img, K)                                                    # :708
        x, _ = self._cs(x, 'Conv_P1', 'conv_p2')
        x, _ = self._cs(x, 'Conv_P2', 'conv_0_c2f')
        x, _ = self._c2f(x, 'C2F_2', ['conv_b_0_c2f', 'conv_b_1_c2f', 'conv_b_2_c2f', 'conv_p3'], 1, True)
        x, _ = self._cs(x, 'Conv_P3', 'conv_2_c2f')
        x, s1 = self._c2f(x, 'C2F_4', ['conv_b1_c2f', 'conv_b2_c2f', 'conv_b3_c2f', 'conv_b4_c2f', 'conv_b5_c2f', 'conv_5'], 2, True)
        r1 = x
        x, _ = self._cs(x, 'Conv_P4', 'cf2_conv_4')
        x, s2 = self._c2f(x, 'C2F_6', ['cf2_bconv_4', 'cf2_bconv1_4', 'cf2_bconv_5', 'cf2_bconv1_5', 'cf2_6_conv_last', 'conv7'], 2, True)
        r2 = x
        x, _ = self._cs(x, 'Conv_P5', 'cf2_conv_6')
        x, _ = self._c2f(x, 'C2F_8', ['cf2_bottle_6', 'cf2_bottle_61', 'cf2_conv_7', 'sppf_conv_1'], 1, True)
        x, _ = self._cs(x, 'SPPF_conv_0', 'sppf_conv_2')                               # :876-879
        p1 = maxpool5(x)
        p2 = maxpool5(p1)
        p3 = maxpool5(p2)
        x, s_sppf = self._cs(np.concatenate((x, p1, p2, p3), 1), 'SPPF_conv_1', 'cf2_conv_8')   # :888-894
        sppf_out = x
        u = self._requant(upsample2(x), s_sppf, s2)                                    # :900-903
        x, s4 = self._c2f(np.concatenate((u, r2), 1), 'C2F_12', ['cf2_conv_80', 'cf2_conv_81', 'cf2_conv_9', 'cf2_conv_10'], 1, False)
        r4 = x
        u = self._requant(upsample2(x), s4, s1)                                        # :935-939
        x, _ = self._c2f(np.concatenate((u, r1), 1), 'C2F_15', ['cf2_bottle_8', 'cf2_bottle_81', 'cf2_conv_11', 'conv8'], 1, False)
        r5 = x
        x, s3 = self._cs(x, 'Conv_16', 'cf2_conv_12')                                  # :969-971
        r4q = self._requant(r4, s4, s3)                                                # :975
        x, _ = self._c2f(np.concatenate((x, r4q), 1), 'C2F_18', ['cf2_bottle_9', 'cf2_bottle_90', 'cf2_conv_13', 'conv9'], 1, False)
        r6 = x
        x, s3 = self._cs(x, 'Conv_19', 'cf2_conv_14')                                  # :1006-1008
        sq = self._requant(sppf_out, s_sppf, s3)                                       # :1012
        x, _ = self._c2f(np.concatenate((x, sq), 1), 'C2F_21', ['cf2_bottle_10', 'cf2_bottle_101', 'cf2_conv_15', 'x_down_0'], 1, False)
        r7 = x

        box_maps, cls_acc = [], []
        for feat, nm in ((r5, 'x_result_5'), (r6, 'x_result_6'), (r7, 'x')):           # :1039-1119
            u, _ = self._cs(feat, f'{nm}_up_0', f'{nm}_up_1')
            u, _ = self._cs(u, f'{nm}_up_1', f'{nm}_up_2')
            u = self._conv(u, f'{nm}_up_2')
            if raw_head:
                box_maps.append((u, self.wl.scales[f'{nm}_up_2']))
                d, _ = self._cs(feat, f'{nm}_down_0', f'{nm}_down_1')
                d, _ = self._cs(d, f'{nm}_down_1', f'{nm}_down_2')
                cls_acc.append((self._conv(d, f'{nm}_down_2'), self.wl.scales[f'{nm}_down_2']))
                continue
            u = self._requant(u, self.wl.scales[f'{nm}_up_2'], scale(DFL_RANGE, K))    # requant_last_layers :472-476
            box_maps.append(u)
            d, _ = self._cs(feat, f'{nm}_down_0', f'{nm}_down_1')
            d, _ = self._cs(d, f'{nm}_down_1', f'{nm}_down_2')
            cls_acc.append((self._conv(d, f'{nm}_down_2'), self.wl.scales[f'{nm}_down_2']))
        if raw_head:
            return box_maps, cls_acc
        cls_maps = [self._requant(a, s, scale(12, 16), 16) for a, s in cls_acc]        # exponent_requant :1146-1149
        return box_maps, cls_maps

    def decode(self, box_maps, cls_maps):
        """DFL decode + class scores, :1155-1261.  Returns dbox (N,4,8400) int64, cls (N,80,8400) int64."""
        K, MK = self.K, self.MK
        n = box_maps[0].shape[0]
        box = np.concatenate([m.reshape(n, 64, -1) for m in box_maps], 2)              # :1158
        a = box.shape[2]
        box = box.reshape(n, 4, 16, a).transpose(0, 2, 1, 3)                           # :1162 (n,16,4,a)
        y = box - box.max(1, keepdims=True)                                            # :1195
        e = self.lut_exp[y + (2 ** K - 1)]                                             # exponent() :455-469
        e = np.rint(e).astype(F32)
        ssum = e.sum(1, keepdims=True, dtype=F32)                                      # :1198
        p = ((e / ssum).astype(F32) * F32(127)).astype(F32).astype(np.int64)           # :1205 (trunc)
        dflw = self.wl.sd['dfl.weight'].reshape(1, 16, 1, 1).astype(np.int64)
        d = (p * dflw).sum(1)                                                          # self.dfl(p) :1232 -> (n,4,a)
        # anchors :101-114, :1220-1227
        pts, strides = [], []
        for hw, st in ((80, 8.), (40, 16.), (20, 32.)):
            sx = np.arange(hw, dtype=F32) + F32(0.5)
            yy, xx = np.meshgrid(sx, sx, indexing='ij')
            pts.append(np.stack((xx.reshape(-1), yy.reshape(-1)), 0))
            strides.append(np.full((hw * hw,), st, F32))
        anchor = np.concatenate(pts, 1).astype(F32)                                    # (2, a)
        strides = np.concatenate(strides)
        z = anchor.max()
        a_scale = F32(scale(F32(z), 16))                                               # 32767 / 79.5 in fp32
        anchor_q = np.rint((anchor * a_scale).astype(F32))
        scale_dfl = float(self.wl.scales['dfl'].reshape(-1)[0])                        # :1233
        kd, sd = coeffs(scale_dfl, float(a_scale))
        dq = requant_apply(d, kd, sd, 16).astype(F32)                                  # :1236
        if self.trace is not None:
            self.trace['requant'].append(dq.astype(np.int64))
            self.trace['coeff'].append((kd, sd))
        lt, rb = dq[:, :2], dq[:, 2:]                                                  # dist2bbox :117-126
        x1y1 = anchor_q[None] - lt
        x2y2 = anchor_q[None] + rb
        cxy = ((x1y1 + x2y2) / F32(2)).astype(F32)
        wh = x2y2 - x1y1
        dbox = (np.concatenate((cxy, wh), 1) * strides[None, None]).astype(F32)        # :1243
        cls = np.concatenate([m.reshape(n, 80, -1) for m in cls_maps], 2)              # :1246
        score = self.lut16[cls + 32767]                                                # :1250
        return dbox, score

    @staticmethod
    def nms_one(dbox, score):
        """coord_quant :297-361 + nms_quant :248-294 + tail :1267-1275 for ONE image.
        dbox (4,a) fp32 xywh, score (80,a) fp32.  Returns (boxes (n,4), classes (n,2)) fp32 or (None, None).
        Tie-break pinned to (score desc, candidate index asc)  (SURVEY hard part 3)."""
        cx, cy, w, h = dbox
        dw = (w / F32(2)).astype(F32)
        dh = (h / F32(2)).astype(F32)
        xyxy = np.stack((cx - dw, cy - dh, cx + dw, cy + dh), 1).astype(F32)           # xywh2xyxy :129-148
        xyxy = np.trunc(xyxy).astype(F32)                                              # .to(torch.int) :316
        conf = score.max(0)
        j = score.argmax(0)                                                            # first max :326
        cand = np.nonzero(conf > 8192)[0]                                              # :299-327
        if cand.size == 0:
            return None, None
        bx = xyxy[cand]
        cf = conf[cand].astype(F32)
        jj = j[cand].astype(F32)
        off = (jj * F32(7680)).astype(F32)[:, None]                                    # :340
        b = (bx + off).astype(F32)                                                     # :344
        x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
        areas = (((x2 - x1).astype(F32) + F32(412)).astype(F32) * ((y2 - y1).astype(F32) + F32(412)).astype(F32)).astype(F32)   # :258
        order = np.argsort(-cf, kind='stable')[:1000]                                  # :260
        keep = []
        while order.size > 0:                                                          # :265-283
            i = order[0]
            keep.append(i)
            rest = order[1:]
            xx1 = np.maximum(x1[i], x1[rest])
            yy1 = np.maximum(y1[i], y1[rest])
            xx2 = np.minimum(x2[i], x2[rest])
            yy2 = np.minimum(y2[i], y2[rest])
            ww = np.maximum(F32(0), ((xx2 - xx1).astype(F32) + F32(412)).astype(F32))
            hh = np.maximum(F32(0), ((yy2 - yy1).astype(F32) + F32(412)).astype(F32))
            inter = (ww * hh).astype(F32)
            inter = (inter * F32(2.22)).astype(F32)
            ok = inter <= ((areas[i] + areas[rest]).astype(F32) - inter).astype(F32)
            order = rest[ok]
        keep = np.array(keep[:300], dtype=np.int64)                                    # :354
        out = np.concatenate((bx[keep], cf[keep, None], jj[keep, None]), 1).astype(F32)
        out[:, :4] = (out[:, :4] / F32(412.1635)).astype(F32)                          # :358
        out[:, 4] = (out[:, 4] / F32(32767.0)).astype(F32)                             # :359
        out[:, :4] = np.clip(out[:, :4], F32(0), F32(640))                             # scale_boxes/clip_boxes :363-423
        return out[:, :4].copy(), out[:, 4:6].copy()                                   # convert_res :426-429

    def forward(self, img, trace=False):
        """list of (boxes, classes) per image, element i == reference model(img[i:i+1])."""
        box_maps, cls_maps = self.forward_maps(img, trace)
        dbox, score = self.decode(box_maps, cls_maps)
        self.last = dict(box_maps=box_maps, cls_maps=cls_maps, dbox=dbox, score=score)
        return [self.nms_one(dbox[i], score[i]) for i in range(dbox.shape[0])]
