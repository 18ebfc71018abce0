"""Plan compiler: reference artefacts -> binary plan blob for libayq.so (ayq_create).

Inputs are exactly what the reference hot path reads at import / load_state_dict time
(paths relative to /root/reference/quantisation/):
  * the stage_7 state_dict  QUANT_WEIGHTS_{K}.pickle   (stage_7.py:748-780; 127 fp32 tensors holding ints)
  * all_scales              bias_scales/*_scale.pickle (utils/save_weights.py:36-42)
  * max_a_dict              results/max_a.txt          (utils/max_a.py:1-7)
The plan fixes, per layer, everything stage_8_torch_full_quant.py recomputes on every call:
rescale coefficients of every requantize() (utils/rescale_coeff_torch.py:14-33, evaluated in fp32
with torch exactly like the reference), the SiLU / exponent / final-sigmoid tables, quantised
anchors (:1212-1227), and the graph wiring of Yolov8.forward (:704-1119) expressed as convolutions
over lists of 16-channel plane segments (concat = list of segments, residual add = duplicated
weights over both addends, SURVEY.md hard part 2).

Layout of the blob: csrc/plan_format.h.
"""
import re
import struct

import numpy as np
import torch

from . import lut as _lut

DFL_RANGE = 14.8264799118042          # stage_8_torch_full_quant.py:436,473
MAGIC = 0x31515941
VERSION = 7
OP_FIELDS = 64
OP_CONV, OP_CONV_P1, OP_POOL, OP_HEAD, OP_NMS, OP_HEAD_FLOAT, OP_NMS_FLOAT = 1, 2, 3, 4, 5, 6, 7
EPI_SILU, EPI_REQUANT8, EPI_REQUANT16 = 0, 1, 2
OUT_IDENT, OUT_REQUANT = 0, 1
MAX_OUT = 3

# all_scales key -> state_dict prefix, in forward order (stage_8_torch_full_quant.py:489-694, :713-1117)
LAYERS = [
    ('Conv_P1', 'conv0.0'), ('Conv_P2', 'conv1.0'), ('C2F_2_conv_0', 'cf2_conv_0.0'), ('C2F_2_bottle_0', 'cf2_bottle_0.0'),
    ('C2F_2_bottle_1', 'cf2_bottle_0.2'), ('C2F_2_conv_1', 'cf2_conv_1.0'), ('Conv_P3', 'conv3.0'),
    ('C2F_4_conv_0', 'cf2_conv_2.0'), ('C2F_4_bottle_0', 'cf2_bottle_2.0'), ('C2F_4_bottle_1', 'cf2_bottle_2.2'),
    ('C2F_4_bottle_2', 'cf2_bottle_3.0'), ('C2F_4_bottle_3', 'cf2_bottle_3.2'), ('C2F_4_conv_1', 'cf2_conv_3.0'),
    ('Conv_P4', 'conv5.0'), ('C2F_6_conv_0', 'cf2_conv_4.0'), ('C2F_6_bottle_0', 'cf2_bottle_4.0'),
    ('C2F_6_bottle_1', 'cf2_bottle_4.2'), ('C2F_6_bottle_2', 'cf2_bottle_5.0'), ('C2F_6_bottle_3', 'cf2_bottle_5.2'),
    ('C2F_6_conv_1', 'cf2_conv_5.0'), ('Conv_P5', 'conv7.0'), ('C2F_8_conv_0', 'cf2_conv_6.0'),
    ('C2F_8_bottle_0', 'cf2_bottle_6.0'), ('C2F_8_bottle_1', 'cf2_bottle_6.2'), ('C2F_8_conv_1', 'cf2_conv_7.0'),
    ('SPPF_conv_0', 'sppf_conv_1.0'), ('SPPF_conv_1', 'sppf_conv_2.0'),
    ('C2F_12_conv_0', 'cf2_conv_8.0'), ('C2F_12_bottle_0', 'cf2_bottle_7.0'), ('C2F_12_bottle_1', 'cf2_bottle_7.2'),
    ('C2F_12_conv_1', 'cf2_conv_9.0'), ('C2F_15_conv_0', 'cf2_conv_10.0'), ('C2F_15_bottle_0', 'cf2_bottle_8.0'),
    ('C2F_15_bottle_1', 'cf2_bottle_8.2'), ('C2F_15_conv_1', 'cf2_conv_11.0'), ('Conv_16', 'conv8.0'),
    ('C2F_18_conv_0', 'cf2_conv_12.0'), ('C2F_18_bottle_0', 'cf2_bottle_9.0'), ('C2F_18_bottle_1', 'cf2_bottle_9.2'),
    ('C2F_18_conv_1', 'cf2_conv_13.0'), ('Conv_19', 'conv9.0'), ('C2F_21_conv_0', 'cf2_conv_14.0'),
    ('C2F_21_bottle_0', 'cf2_bottle_10.0'), ('C2F_21_bottle_1', 'cf2_bottle_10.2'), ('C2F_21_conv_1', 'cf2_conv_15.0'),
    ('x_result_5_up_0', 'detect_5_up.0'), ('x_result_5_up_1', 'detect_5_up.2'), ('x_result_5_up_2', 'detect_5_up.4'),
    ('x_result_5_down_0', 'detect_5_down.0'), ('x_result_5_down_1', 'detect_5_down.2'), ('x_result_5_down_2', 'detect_5_down.4'),
    ('x_result_6_up_0', 'detect_6_up.0'), ('x_result_6_up_1', 'detect_6_up.2'), ('x_result_6_up_2', 'detect_6_up.4'),
    ('x_result_6_down_0', 'detect_6_down.0'), ('x_result_6_down_1', 'detect_6_down.2'), ('x_result_6_down_2', 'detect_6_down.4'),
    ('x_up_0', 'detect_x_up.0'), ('x_up_1', 'detect_x_up.2'), ('x_up_2', 'detect_x_up.4'),
    ('x_down_0', 'detect_x_down.0'), ('x_down_1', 'detect_x_down.2'), ('x_down_2', 'detect_x_down.4'),
]
LAYER_INDEX = {name: i for i, (name, _) in enumerate(LAYERS)}
SD_PREFIX = dict(LAYERS)
STRIDE2 = ('Conv_P1', 'Conv_P2', 'Conv_P3', 'Conv_P4', 'Conv_P5', 'Conv_16', 'Conv_19')


class RescaleOverflow(ValueError):
    """The reference prints 'Problem with rescale coeff' and calls exit() (utils/rescale_coeff_torch.py:31-33)."""


def rescale_coeffs(old_scale, new_scale, bit_size_for_koeff=8):
    """(rescale_koeff, shift_val) of requantize(), utils/rescale_coeff_torch.py:20-33, as fp32 torch tensors.
    old_scale: python float or fp32 tensor (any shape); new_scale: python float."""
    if isinstance(old_scale, float) and isinstance(new_scale, float):
        old_scale = torch.tensor(old_scale, dtype=torch.float32)
        new_scale = torch.tensor(new_scale, dtype=torch.float32)
    else:
        old_scale = torch.as_tensor(old_scale, dtype=torch.float32)
    shift_val = bit_size_for_koeff + torch.floor(torch.log2(old_scale / new_scale))
    rescale_koeff = torch.round((2 ** shift_val) * (new_scale / old_scale))
    if rescale_koeff.max() > (2 ** bit_size_for_koeff) - 1:
        shift_val = shift_val - 1
        rescale_koeff = torch.round((2 ** shift_val) * (new_scale / old_scale))
        if rescale_koeff.max() > (2 ** bit_size_for_koeff) - 1:
            raise RescaleOverflow(f'Problem with rescale coeff: {rescale_koeff} > {(2 ** bit_size_for_koeff) - 1} '
                                  f'({old_scale} and {new_scale})')
    return rescale_koeff.reshape(-1), shift_val.reshape(-1)


def _k_inv(old, new):
    """-> (k float32[C], 2^-s float32[C], k int, s int) ready for the device epilogue."""
    k, s = rescale_coeffs(old, new)
    k = k.numpy().astype(np.float32)
    s = s.numpy().astype(np.float64)
    inv = np.ldexp(1.0, -s.astype(np.int64)).astype(np.float32)
    return k, inv


def parse_max_a(text):
    """utils/max_a.py:1-7 on the content of results/max_a.txt"""
    d = {}
    for el in str(text).splitlines(True):
        if not el.strip():
            continue
        d[el.split(' ')[0][:-1]] = float(el.split(' ')[1].rstrip('\n'))
    return d


# ----------------------------------------------------------------------------- logical tensors
class Part:
    """`nplanes` 16-channel planes whose value is the SUM of the physical plane ranges in `addends`."""
    __slots__ = ('nplanes', 'addends')

    def __init__(self, nplanes, addends):
        self.nplanes = nplanes
        self.addends = list(addends)          # [(buf, plane0)]


class LT:
    """Logical activation tensor = channel concatenation of Parts (all at one spatial size / scale)."""

    def __init__(self, parts, h, w, ps_buf=None, ps_only=False):
        self.parts, self.h, self.w = list(parts), h, w
        self.ps_buf = ps_buf        # buffer holding a phase-split copy [(y&1)*2 + (x&1)][plane][n][h/2][w/2][16] (for stride-2 consumers)
        self.ps_only = ps_only      # the phase-split copy is the only materialisation (parts is empty)

    @property
    def nplanes(self):
        return sum(p.nplanes for p in self.parts)

    def split_half(self):
        """torch.split(x, c/2, dim=1) on a single-part tensor (stage_8_torch_full_quant.py:54-58)."""
        assert len(self.parts) == 1 and len(self.parts[0].addends) == 1
        (buf, p0), n = self.parts[0].addends[0], self.parts[0].nplanes
        assert n % 2 == 0
        return (LT([Part(n // 2, [(buf, p0)])], self.h, self.w),
                LT([Part(n // 2, [(buf, p0 + n // 2)])], self.h, self.w))

    def add(self, other):
        """x += x_bottle (unclipped residual, :742): keep both addends, the consumer duplicates its weights."""
        assert len(self.parts) == len(other.parts) == 1 and self.parts[0].nplanes == other.parts[0].nplanes
        return LT([Part(self.parts[0].nplanes, self.parts[0].addends + other.parts[0].addends)], self.h, self.w)

    @staticmethod
    def cat(ts):
        return LT([p for t in ts for p in t.parts], ts[0].h, ts[0].w)


class PlanBuilder:
    def __init__(self, sd, scales, max_a, K, sigmoid_range=6, taps=False, img=640, head='int', phase_split=None):
        self.sd = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in sd.items()}
        self.scales = {k: torch.as_tensor(np.asarray(v, dtype=np.float32)).reshape(-1) for k, v in scales.items()}
        self.max_a = max_a
        self.K = int(K)
        self.MK = 2 ** (self.K - 1) - 1
        self.sigmoid_range = sigmoid_range
        self.taps = taps
        self.img = img
        assert head in ('int', 'float')
        import os as _os
        self.phase_split = (_os.environ.get('AYQ_NO_PS') is None) if phase_split is None else bool(phase_split)
        self.head = head                # 'int': stage_8_torch_full_quant.py (DFL / scores / q_NMS in integers); 'float': stage_8_torch.py:915-961
        self.bufs = []              # (name, nplanes, H, W, elem_bytes)
        self.ops = []               # list of int lists
        self.data = bytearray()
        self.info = {'layers': {}, 'bufs': {}, 'silu_taps': [], 'requant_taps': [], 'acc_taps': [], 'ps_bufs': []}
        self.coeff_log = []         # (k, s) of every requantize() in reference call order is rebuilt by tests from info
        _, lut_arr = _lut.cached_array('sigmoid', sigmoid_range, self.K)
        self.lut_off = self.add_data(lut_arr.astype(np.float32))
        self.n_acc_taps = 0

    # -- blob helpers
    def add_data(self, arr):
        b = np.ascontiguousarray(arr).tobytes()
        while len(self.data) % 256:
            self.data.append(0)
        off = len(self.data)
        self.data += b
        return off

    def new_buf(self, name, nplanes, h, w, elem_bytes=1):
        self.bufs.append((name, nplanes, h, w, elem_bytes))
        self.info['bufs'][name] = len(self.bufs) - 1
        return len(self.bufs) - 1

    @staticmethod
    def fbits(x):
        return struct.unpack('<i', struct.pack('<f', float(x)))[0]

    # -- layers
    def conv(self, name, x, next_a=None, epi=EPI_SILU, outs=None, out_scale_new=None, acc_buf=False):
        """One Conv2d + epilogue.  x: LT.  outs: list of dicts {requant: (old,new)|None, up: bool} for EPI_SILU.
        Returns (list of LT (one per out), python-float scale of the silu result)."""
        w = self.sd[SD_PREFIX[name] + '.weight']
        b = self.sd[SD_PREFIX[name] + '.bias']
        cout, cin, ks, _ = w.shape
        stride = 2 if name in STRIDE2 else 1
        use_ps = stride == 2 and x.ps_buf is not None
        assert not x.ps_only or use_ps, name
        nplanes_in = cin // 16
        assert use_ps or cin == 16 * x.nplanes, (name, cin, x.nplanes)
        assert cout % 16 == 0
        hout = (x.h + 2 * (ks // 2) - ks) // stride + 1
        wout = (x.w + 2 * (ks // 2) - ks) // stride + 1
        wq = np.rint(w).astype(np.int64)
        assert np.abs(wq).max() <= 127 and np.array_equal(wq, w)
        # K chunks: for part, for addend, for tap, for plane
        kc, wrows = [], []
        c0 = 0
        if use_ps:
            # stride-2 3x3 conv on the phase-split copy: tap (ky, kx) of output (oy, ox) reads input (2oy + ky - 1, 2ox + kx - 1)
            # = phase ((ky-1)&1, (kx-1)&1) at (oy + dy, ox + dx) with dy = -1 for ky == 0 else 0: a stride-1 box per tap, so
            # the TMA moves whole 16*bw byte rows instead of one strided 16-byte pixel at a time
            assert ks == 3 and x.h % 2 == 0 and x.w % 2 == 0
            for ky in range(3):
                for kx in range(3):
                    ph = (((ky - 1) & 1) << 1) | ((kx - 1) & 1)
                    for pl in range(nplanes_in):
                        kc.append((x.ps_buf, ph * nplanes_in + pl, 0 if ky == 0 else 1, 0 if kx == 0 else 1))
                        wrows.append(wq[:, 16 * pl:16 * pl + 16, ky, kx])
        for part in ([] if use_ps else x.parts):
            for (buf, p0) in part.addends:
                for ky in range(ks):
                    for kx in range(ks):
                        for pl in range(part.nplanes):
                            kc.append((buf, p0 + pl, ky, kx))
                            ch = c0 + 16 * pl
                            wrows.append(wq[:, ch:ch + 16, ky, kx])          # (cout, 16)
            c0 += 16 * part.nplanes
        nkc = len(kc)
        if nkc % 2:
            wrows.append(np.zeros((cout, 16), np.int64))
        wpack = np.stack(wrows, 0).astype(np.int8)                           # (nkc_pad, cout, 16)
        sx = self.scales[name]
        assert sx.numel() == cout
        f = [0] * OP_FIELDS
        f[0] = OP_CONV
        f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8] = ks, stride, x.h, x.w, hout, wout, cout, nkc
        if use_ps:
            f[2], f[3], f[4] = 1, x.h // 2, x.w // 2
        f[9] = self.add_data(np.array(kc, np.int32))
        f[10] = self.add_data(wpack)
        bq = np.rint(b).astype(np.int64)
        assert np.abs(bq).max() < 2 ** 31
        f[11] = self.add_data(bq.astype(np.int32))
        f[13] = epi
        f[15] = self.lut_off
        f[40] = LAYER_INDEX[name]
        f[41] = -1
        if self.taps:
            f[41] = self.n_acc_taps
            self.n_acc_taps += 1
            self.info['acc_taps'].append((name, cout, hout, wout))
        f[42] = self.add_data(np.frombuffer(name.encode() + b'\0', np.uint8))
        f[43] = self.new_buf(f'{name}.acc', cout // 16, hout, wout, 4) if acc_buf else -1     # NCHW int32 raw accumulators
        results = []
        new_scale = None
        if epi == EPI_SILU:
            k1, i1 = _k_inv(sx, _lut.scale(self.sigmoid_range, self.K))                 # silu() :440-443
            scale_silu = (_lut.scale(1, self.K) * sx)                                    # :448  fp32 tensor
            new_scale = _lut.scale(self.max_a[next_a], self.K)                           # :450
            k2, i2 = _k_inv(scale_silu, new_scale)
            f[12] = self.add_data(np.stack([k1, i1, k2, i2]).astype(np.float32))
            f[14] = self.MK
            outs = outs if outs is not None else [dict(requant=None, up=False)]
            outs = list(outs)
            if not self.phase_split:
                outs = [o for o in outs if not o.get('ps')] or [dict(requant=None, up=False)]
            n_req = len(outs)
            if self.taps and all(o['requant'] is not None or o['up'] or o.get('ps') for o in outs):
                outs.append(dict(requant=None, up=False, tap=True))                       # raw silu result for parity tests
            assert len(outs) <= MAX_OUT
            f[16] = len(outs)
            ps_of = None
            for i, o in enumerate(outs):
                up = 2 if o['up'] else 1
                if o.get('ps'):
                    assert o['requant'] is None and not o['up'] and hout % 2 == 0 and wout % 2 == 0
                    ob = self.new_buf(f'{name}.out{i}.ps', 4 * (cout // 16), hout // 2, wout // 2)
                    self.info['ps_bufs'].append(ob)
                    base = 17 + 6 * i
                    f[base], f[base + 1], f[base + 2], f[base + 3], f[base + 4], f[base + 5] = ob, 0, OUT_IDENT, self.fbits(1.0), self.fbits(1.0), 2
                    ps_of = ob
                    results.append(LT([], hout, wout, ps_buf=ob, ps_only=True))
                    self.info['layers'].setdefault(name, {})['ps_buf'] = ob
                    continue
                ob = self.new_buf(f'{name}.out{i}', cout // 16, hout * up, wout * up)
                base = 17 + 6 * i
                f[base], f[base + 1] = ob, 0
                if o['requant'] is not None:
                    kq, iq = _k_inv(float(o['requant'][0]), float(o['requant'][1]))
                    f[base + 2], f[base + 3], f[base + 4] = OUT_REQUANT, self.fbits(kq[0]), self.fbits(iq[0])
                else:
                    f[base + 2], f[base + 3], f[base + 4] = OUT_IDENT, self.fbits(1.0), self.fbits(1.0)
                f[base + 5] = 1 if o['up'] else 0
                lt = LT([Part(cout // 16, [(ob, 0)])], hout * up, wout * up)
                results.append(lt)
                if o.get('tap') or (o['requant'] is None and not o['up']):
                    self.info['layers'].setdefault(name, {})['silu_buf'] = ob
                else:
                    self.info['layers'].setdefault(name, {}).setdefault('requant_bufs', []).append((ob, o['up']))
        else:
            bits = self.K if epi == EPI_REQUANT8 else 16
            kq, iq = _k_inv(sx, float(out_scale_new))
            f[12] = self.add_data(np.stack([kq, iq, kq, iq]).astype(np.float32))
            f[14] = 2 ** (bits - 1) - 1
            ob = self.new_buf(f'{name}.out0', cout // 16, hout, wout, 1 if epi == EPI_REQUANT8 else 2)
            f[16] = 1
            f[17], f[18], f[19], f[20], f[21], f[22] = ob, 0, OUT_IDENT, self.fbits(1.0), self.fbits(1.0), 0
            results.append(LT([Part(cout // 16, [(ob, 0)])], hout, wout))
            self.info['layers'].setdefault(name, {})['requant_bufs'] = [(ob, False)]
        if epi == EPI_SILU:
            results = results[:n_req]
            if len(results) > 1 and any(r.ps_only for r in results):
                # a phase-split copy next to the plain tensor: attach it to the first plain result, callers see one tensor less
                plain = [r for r in results if not r.ps_only]
                plain[0].ps_buf = [r for r in results if r.ps_only][0].ps_buf
                results = plain
        self.ops.append(f)
        self.info['layers'][name].update(dict(op=len(self.ops) - 1, cout=cout, hout=hout, wout=wout, nkc=nkc,
                                               ks=ks, stride=stride, macs=cout * cin * ks * ks * hout * wout,
                                               in_bytes=len(set((b_, p_) for b_, p_, _, _ in kc)) * x.h * x.w * 16 // (4 if use_ps else 1),
                                               # algorithmic output bytes: a phase-split COPY of a tensor that is also stored plain
                                               # is a layout choice, not algorithmic traffic
                                               out_bytes=sum(self.bufs[f[17 + 6 * i]][1] * self.bufs[f[17 + 6 * i]][2] * self.bufs[f[17 + 6 * i]][3]
                                                             * 16 * self.bufs[f[17 + 6 * i]][4] for i in range(f[16])
                                                             if not (f[17 + 6 * i + 5] == 2 and f[16] > 1)),
                                               # SURVEY.md 8(d) figures, independent of layout / fusion choices: the reference conv's
                                               # input read once and its output written once at 1 B per element
                                               alg_in_bytes=cin * x.h * x.w, alg_out_bytes=cout * hout * wout,
                                               kmacs=cout * 16 * nkc * hout * wout))
        return results, new_scale

    def conv_p1(self, next_a):
        """Conv_P1 with the fused input quantiser (quant_matrix, utils/quant_matrix_torch.py:57-70; :708-716)."""
        name = 'Conv_P1'
        w = np.rint(self.sd[SD_PREFIX[name] + '.weight']).astype(np.int64)   # (16,3,3,3)
        b = np.rint(self.sd[SD_PREFIX[name] + '.bias']).astype(np.int64)
        assert w.shape == (16, 3, 3, 3)
        wp = np.zeros((16, 32), np.int8)
        for ky in range(3):
            for kx in range(3):
                for c in range(3):
                    wp[:, (ky * 3 + kx) * 3 + c] = w[:, c, ky, kx]
        sx = self.scales[name]
        k1, i1 = _k_inv(sx, _lut.scale(self.sigmoid_range, self.K))
        new_scale = _lut.scale(self.max_a[next_a], self.K)
        k2, i2 = _k_inv(_lut.scale(1, self.K) * sx, new_scale)
        h = self.img // 2
        ps = self.phase_split and h % 2 == 0
        ob = self.new_buf('Conv_P1.out0', 4, h // 2, h // 2) if ps else self.new_buf('Conv_P1.out0', 1, h, h)
        if ps:
            self.info['ps_bufs'].append(ob)
        f = [0] * OP_FIELDS
        f[0] = OP_CONV_P1
        f[1], f[2], f[3] = h, h, ob
        f[4] = self.add_data(wp)
        f[5] = self.add_data(b.astype(np.int32))
        f[6] = self.add_data(np.stack([k1, i1, k2, i2]).astype(np.float32))
        f[7] = self.MK
        f[8] = self.lut_off
        f[9] = -1
        f[10] = -1
        f[11] = 1 if ps else 0                                       # P1_OUT_PS: phase-split output (its only consumer is the stride-2 Conv_P2)
        if self.taps:
            f[9] = self.n_acc_taps
            self.n_acc_taps += 1
            self.info['acc_taps'].append((name, 16, h, h))
        self.ops.append(f)
        self.info['layers'][name] = dict(op=len(self.ops) - 1, silu_buf=ob, cout=16, hout=h, wout=h, nkc=2, ks=3, stride=2,
                                         macs=16 * 27 * h * h, kmacs=16 * 32 * h * h,
                                         in_bytes=4 * 3 * self.img * self.img, out_bytes=16 * h * h)
        if ps:
            return LT([], h, h, ps_buf=ob, ps_only=True), new_scale
        return LT([Part(1, [(ob, 0)])], h, h), new_scale

    def c2f(self, x, name, a_keys, n_bottle, add, final_outs=None):
        """C2f block, e.g. :723-748 (backbone, add=True) and :908-929 (neck, add=False)."""
        (x,), s0 = self.conv(f'{name}_conv_0', x, a_keys[0])
        x0, x1 = x.split_half()
        parts = [x0, x1]
        cur = x1
        for i in range(n_bottle):
            (y,), _ = self.conv(f'{name}_bottle_{2 * i}', cur, a_keys[1 + 2 * i])
            sy = _lut.scale(self.max_a[a_keys[2 + 2 * i]], self.K)
            (y,), _ = self.conv(f'{name}_bottle_{2 * i + 1}', y, a_keys[2 + 2 * i],
                                outs=[dict(requant=(sy, s0), up=False)])                 # requantize(x, sy, s0) :741
            cur = y.add(cur) if add else y                                               # x += x_bottle_0 :742
            parts.append(cur)
        return self.conv(f'{name}_conv_1', LT.cat(parts), a_keys[-1], outs=final_outs)

    def build(self):
        K = self.K
        S = lambda key: _lut.scale(self.max_a[key], K)
        x, _ = self.conv_p1('conv_p2')
        (x,), _ = self.conv('Conv_P2', x, 'conv_0_c2f')
        PS, ID = dict(requant=None, up=False, ps=True), dict(requant=None, up=False)
        (x,), _ = self.c2f(x, 'C2F_2', ['conv_b_0_c2f', 'conv_b_1_c2f', 'conv_b_2_c2f', 'conv_p3'], 1, True, final_outs=[PS])
        (x,), _ = self.conv('Conv_P3', x, 'conv_2_c2f')
        (r1,), s1 = self.c2f(x, 'C2F_4', ['conv_b1_c2f', 'conv_b2_c2f', 'conv_b3_c2f', 'conv_b4_c2f', 'conv_b5_c2f', 'conv_5'], 2, True, final_outs=[ID, PS])
        (x,), _ = self.conv('Conv_P4', r1, 'cf2_conv_4')
        (r2,), s2 = self.c2f(x, 'C2F_6', ['cf2_bconv_4', 'cf2_bconv1_4', 'cf2_bconv_5', 'cf2_bconv1_5', 'cf2_6_conv_last', 'conv7'], 2, True, final_outs=[ID, PS])
        (x,), _ = self.conv('Conv_P5', r2, 'cf2_conv_6')
        (x,), _ = self.c2f(x, 'C2F_8', ['cf2_bottle_6', 'cf2_bottle_61', 'cf2_conv_7', 'sppf_conv_1'], 1, True)
        # SPPF :875-897
        (x,), _ = self.conv('SPPF_conv_0', x, 'sppf_conv_2')
        pb = self.new_buf('SPPF.pools', 3 * x.nplanes, x.h, x.w)
        (ib, ip0) = x.parts[0].addends[0]
        f = [0] * OP_FIELDS
        f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7] = OP_POOL, ib, ip0, x.nplanes, pb, 0, x.h, x.w
        self.ops.append(f)
        n = x.nplanes
        pools = [LT([Part(n, [(pb, i * n)])], x.h, x.w) for i in range(3)]
        s_sppf = S('cf2_conv_8')
        s3_19 = S('cf2_conv_14')                    # scale of the Conv_19 output (:1006-1012)
        # SPPF_conv_1 result feeds (a) upsample + requantize to s2 (:900-903) and (b) requantize to Conv_19's scale (:1012)
        (u, sq), _ = self.conv('SPPF_conv_1', LT.cat([x] + pools), 'cf2_conv_8',
                               outs=[dict(requant=(s_sppf, s2), up=True), dict(requant=(s_sppf, s3_19), up=False)])
        s4 = S('cf2_conv_10')
        s3_16 = S('cf2_conv_12')                    # scale of the Conv_16 output (:969-975)
        (u4, r4q), _ = self.c2f(LT.cat([u, r2]), 'C2F_12', ['cf2_conv_80', 'cf2_conv_81', 'cf2_conv_9', 'cf2_conv_10'], 1, False,
                                final_outs=[dict(requant=(s4, s1), up=True), dict(requant=(s4, s3_16), up=False)])
        (r5,), _ = self.c2f(LT.cat([u4, r1]), 'C2F_15', ['cf2_bottle_8', 'cf2_bottle_81', 'cf2_conv_11', 'conv8'], 1, False, final_outs=[ID, PS])
        (x,), _ = self.conv('Conv_16', r5, 'cf2_conv_12')
        (r6,), _ = self.c2f(LT.cat([x, r4q]), 'C2F_18', ['cf2_bottle_9', 'cf2_bottle_90', 'cf2_conv_13', 'conv9'], 1, False, final_outs=[ID, PS])
        (x,), _ = self.conv('Conv_19', r6, 'cf2_conv_14')
        (r7,), _ = self.c2f(LT.cat([x, sq]), 'C2F_21', ['cf2_bottle_10', 'cf2_bottle_101', 'cf2_conv_15', 'x_down_0'], 1, False)

        box_bufs, cls_bufs = [], []
        fl = self.head == 'float'
        for feat, nm in ((r5, 'x_result_5'), (r6, 'x_result_6'), (r7, 'x')):               # :1039-1119
            (u,), _ = self.conv(f'{nm}_up_0', feat, f'{nm}_up_1')
            (u,), _ = self.conv(f'{nm}_up_1', u, f'{nm}_up_2')
            (u,), _ = self.conv(f'{nm}_up_2', u, epi=EPI_REQUANT8, out_scale_new=_lut.scale(DFL_RANGE, K), acc_buf=fl)   # :472-476
            box_bufs.append(self.ops[-1][43] if fl else u.parts[0].addends[0][0])
            (d,), _ = self.conv(f'{nm}_down_0', feat, f'{nm}_down_1')
            (d,), _ = self.conv(f'{nm}_down_1', d, f'{nm}_down_2')
            (d,), _ = self.conv(f'{nm}_down_2', d, epi=EPI_REQUANT16, out_scale_new=_lut.scale(12, 16), acc_buf=fl)      # :1146-1149
            cls_bufs.append(self.ops[-1][43] if fl else d.parts[0].addends[0][0])
        if fl:
            # stage_8_torch.py:915-961: the head reads the six raw accumulators (the requantised copies above are unused);
            # decode and coord() run in fp32 (head_float_kernel / nms_float_kernel)
            dflw = self.sd['dfl.weight'].reshape(-1).astype(np.float32)
            assert dflw.shape == (16,)
            names = ('x_result_5', 'x_result_6', 'x')
            f = [0] * OP_FIELDS
            f[0] = OP_HEAD_FLOAT
            f[1:4] = box_bufs
            f[4:7] = cls_bufs
            f[7] = self.add_data(np.stack([self.scales[f'{nm}_up_2'].numpy() for nm in names]).astype(np.float32))
            f[8] = self.add_data(np.stack([self.scales[f'{nm}_down_2'].numpy() for nm in names]).astype(np.float32))
            f[9] = self.add_data(dflw)
            self.ops.append(f)
            f = [0] * OP_FIELDS
            f[0] = OP_NMS_FLOAT
            self.ops.append(f)
            self.info['n_anchors'] = sum((self.img // st) ** 2 for st in (8, 16, 32))
            self.info['box_bufs'], self.info['cls_bufs'] = box_bufs, cls_bufs
            return self

        # head constants :1158-1243
        _, lut_exp = _lut.cached_array('exp', DFL_RANGE, K)
        _, lut16 = _lut.cached_array('sigmoid', 12, 16)
        dflw = np.rint(self.sd['dfl.weight'].reshape(-1)).astype(np.int64)
        assert dflw.shape == (16,) and np.array_equal(dflw, self.sd['dfl.weight'].reshape(-1))
        anchors, a_scale = quantised_anchors(self.img)
        kd, idd = _k_inv(float(self.scales['dfl'].reshape(-1)[0].item()), float(a_scale))
        f = [0] * OP_FIELDS
        f[0] = OP_HEAD
        f[1:4] = box_bufs
        f[4:7] = cls_bufs
        f[7] = self.add_data(lut_exp.astype(np.float32))
        f[8] = self.add_data(lut16.astype(np.int16))
        f[9] = self.add_data(dflw.astype(np.int32))
        f[10] = self.add_data(anchors.astype(np.int32))
        f[11], f[12] = self.fbits(kd[0]), self.fbits(idd[0])
        # The final sigmoid table is monotone, so max_c LUT[l_c] = LUT[max_c l_c] and the first arg-max class is the first
        # class whose logit reaches lo16[max logit] = the smallest logit with the same table value (one gather per anchor
        # instead of 80).  f[14] tells the kernel whether the table really is monotone (else it looks every class up).
        l16 = lut16.astype(np.int64)
        mono = bool(np.all(np.diff(l16) >= 0))
        first = np.concatenate([[True], l16[1:] != l16[:-1]])
        lo_idx = np.maximum.accumulate(np.where(first, np.arange(l16.size), 0))
        f[13] = self.add_data((lo_idx - 32767).astype(np.int16))
        f[14] = 1 if mono else 0
        self.ops.append(f)
        f = [0] * OP_FIELDS
        f[0] = OP_NMS
        self.ops.append(f)
        self.info['n_anchors'] = anchors.shape[0]
        self.info['box_bufs'], self.info['cls_bufs'] = box_bufs, cls_bufs
        return self

    def blob(self):
        hdr_size = 8 + 4 * 6 + 8 * 4
        bufs = b''.join(struct.pack('<4i', b[1], b[2], b[3], b[4]) for b in self.bufs)
        ops = b''.join(struct.pack(f'<{OP_FIELDS}i', *op) for op in self.ops)
        bufs_off = hdr_size
        ops_off = bufs_off + len(bufs)
        data_off = (ops_off + len(ops) + 255) // 256 * 256
        hdr = struct.pack('<IIiiiiii4Q', MAGIC, VERSION, self.K, len(self.bufs), len(self.ops), self.img, self.img,
                          self.info['n_anchors'], bufs_off, ops_off, data_off, len(self.data))
        assert len(hdr) == hdr_size
        out = bytearray(hdr + bufs + ops)
        out += b'\0' * (data_off - len(out))
        out += self.data
        return bytes(out)


def quantised_anchors(img=640):
    """make_anchors (:101-114) + quant_anchors (:1220-1227): int anchor points (a,2) and the fp32 anchor scale."""
    pts = []
    for st in (8, 16, 32):
        hw = img // st
        sx = torch.arange(end=hw, dtype=torch.float32) + 0.5
        sy, sxx = torch.meshgrid(sx, sx, indexing='ij')
        pts.append(torch.stack((sxx, sy), -1).view(-1, 2))
    anchor = torch.cat(pts).transpose(0, 1)                      # (2, a)
    z = torch.max(anchor)
    a_scale = _lut.scale(z, 16)                                  # 0-dim fp32 tensor, like the reference
    q = torch.round(torch.clamp(anchor, -z, z) * a_scale)
    return q.transpose(0, 1).contiguous().numpy().astype(np.int64), float(a_scale.item())


class Plan:
    """Compiled plan: the blob for ayq_create plus python-side metadata (buffer ids of every tap)."""

    def __init__(self, builder):
        self.K = builder.K
        self.blob = builder.blob()
        self.info = builder.info
        self.bufs = builder.bufs
        self.n_ops = len(builder.ops)
        self.n_acc_taps = builder.n_acc_taps
        self.taps = builder.taps
        by_op = {meta['op']: nm for nm, meta in builder.info['layers'].items()}
        generic = {OP_POOL: 'sppf_pool', OP_HEAD: 'head(dfl+scores)', OP_NMS: 'q_NMS', OP_HEAD_FLOAT: 'head(float)', OP_NMS_FLOAT: 'coord(float NMS)'}
        self.op_names = [by_op.get(i, generic.get(op[0], f'op{i}')) for i, op in enumerate(builder.ops)]


def compile_plan(state_dict, all_scales, max_a_dict, K=8, sigmoid_range=6, taps=False, head='int'):
    """state_dict: the 127-key stage_7 dict; all_scales: {layer: fp32 (1,C,1,1) or (C,)}; max_a_dict: {name: float}.
    head='int' (sigmoid_range 6) is stage_8_torch_full_quant.py; head='float' with sigmoid_range=7 is stage_8_torch.py."""
    missing = [n for n, p in LAYERS if p + '.weight' not in state_dict or n not in all_scales]
    if missing:
        raise KeyError(f'plan: missing weights/scales for {missing[:4]}...')
    if 'dfl' not in all_scales or 'dfl.weight' not in state_dict:
        raise KeyError('plan: missing dfl weight / scale')
    return Plan(PlanBuilder(state_dict, all_scales, max_a_dict, K, sigmoid_range, taps, head=head).build())


def header_constants(path):
    """Parse enum / #define integer constants of csrc/plan_format.h (tests keep both sides in sync)."""
    txt = open(path).read()
    txt = re.sub(r'//.*', '', txt)
    out = {}
    for m in re.finditer(r'#define\s+(\w+)\s+(0x[0-9a-fA-F]+|\d+)u?\b', txt):
        out[m.group(1)] = int(m.group(2), 0)
    for m in re.finditer(r'enum\s*\{([^}]*)\}', txt, re.S):
        nxt = 0
        for item in m.group(1).split(','):
            item = item.strip()
            if not item:
                continue
            if '=' in item:
                nm, val = item.split('=')
                nxt = int(val.strip(), 0)
                out[nm.strip()] = nxt
            else:
                out[item] = nxt
            nxt += 1
    return out
