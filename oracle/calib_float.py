"""CPU oracle of the calibration forward -- TEST INFRASTRUCTURE (see oracle/yolo_int.py for who may import it).

Restates Yolov8.forward of /root/reference/quantisation/stage_4.py:475-946: the BN-fused FLOAT network
(Conv2d + bias, nn.SiLU) with a save_max_a() tap (utils/save_a.py:11-26: abs(t).max() over the whole tensor) on the input, on
every convolution output BEFORE its SiLU (64 taps; the 'silu_p1' and second 'conv_b_0_c2f' calls are commented out there).
Tap names and order are the reference's.  The wiring (C2f split / residual / concat, SPPF pools, upsample + concat) is the one oracle/yolo_int.py pins bit-exactly.

torch CPU float ops, like the reference.  Pin: tests/test_oracle_golden.py checks the first eight taps (everything up to
Conv_P3, weights fixture tests/golden/bnf_head_k8.npz, 125 KB) against the reference's own results/max_a_all.txt for its six
calibration images; tools/pin_calib_oracle.py checks ALL 64 taps when the harness work directory (the full 12 MB fused
weights) is present -- run in the build container, result recorded in DESIGN.md.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .yolo_int import _SD

STRIDE2 = ('Conv_P1', 'Conv_P2', 'Conv_P3', 'Conv_P4', 'Conv_P5', 'Conv_16', 'Conv_19')
# reference tap name of every conv output, in forward order (stage_4.py:477-909)
TAP = {
    'Conv_P1': 'conv_p1', 'Conv_P2': 'conv_p2', 'C2F_2_conv_0': 'conv_0_c2f', 'C2F_2_bottle_0': 'conv_b_0_c2f', 'C2F_2_bottle_1': 'conv_b_1_c2f',
    'C2F_2_conv_1': 'conv_b_2_c2f', 'Conv_P3': 'conv_p3', 'C2F_4_conv_0': 'conv_2_c2f', 'C2F_4_bottle_0': 'conv_b1_c2f',
    'C2F_4_bottle_1': 'conv_b2_c2f', 'C2F_4_bottle_2': 'conv_b3_c2f', 'C2F_4_bottle_3': 'conv_b4_c2f', 'C2F_4_conv_1': 'conv_b5_c2f',
    'Conv_P4': 'conv_5', 'C2F_6_conv_0': 'cf2_conv_4', 'C2F_6_bottle_0': 'cf2_bconv_4', 'C2F_6_bottle_1': 'cf2_bconv1_4',
    'C2F_6_bottle_2': 'cf2_bconv_5', 'C2F_6_bottle_3': 'cf2_bconv1_5', 'C2F_6_conv_1': 'cf2_6_conv_last', 'Conv_P5': 'conv7',
    'C2F_8_conv_0': 'cf2_conv_6', 'C2F_8_bottle_0': 'cf2_bottle_6', 'C2F_8_bottle_1': 'cf2_bottle_61', 'C2F_8_conv_1': 'cf2_conv_7',
    'SPPF_conv_0': 'sppf_conv_1', 'SPPF_conv_1': 'sppf_conv_2', 'C2F_12_conv_0': 'cf2_conv_8', 'C2F_12_bottle_0': 'cf2_conv_80',
    'C2F_12_bottle_1': 'cf2_conv_81', 'C2F_12_conv_1': 'cf2_conv_9', 'C2F_15_conv_0': 'cf2_conv_10', 'C2F_15_bottle_0': 'cf2_bottle_8',
    'C2F_15_bottle_1': 'cf2_bottle_81', 'C2F_15_conv_1': 'cf2_conv_11', 'Conv_16': 'conv8', 'C2F_18_conv_0': 'cf2_conv_12',
    'C2F_18_bottle_0': 'cf2_bottle_9', 'C2F_18_bottle_1': 'cf2_bottle_90', 'C2F_18_conv_1': 'cf2_conv_13', 'Conv_19': 'conv9',
    'C2F_21_conv_0': 'cf2_conv_14', 'C2F_21_bottle_0': 'cf2_bottle_10', 'C2F_21_bottle_1': 'cf2_bottle_101', 'C2F_21_conv_1': 'cf2_conv_15',
}
for _nm in ('x_result_5', 'x_result_6', 'x'):
    for _br in ('up', 'down'):
        for _i in range(3):
            TAP[f'{_nm}_{_br}_{_i}'] = f'{_nm}_{_br}_{_i}'


class Stop(Exception):
    pass


class CalibOracle:
    """sd: BN-fused float state_dict (stage_2 output) as {name: numpy float32}.  forward(x) returns the ordered tap list
    [(name, value)] for ONE image batch x (1,3,640,640) float32, like one iteration of the stage_4 loop (:978-983)."""

    def __init__(self, sd, stop_after=None):
        self.sd = {k: torch.as_tensor(np.asarray(v)) for k, v in sd.items()}
        self.stop_after = stop_after

    def _tap(self, name, t):
        self.taps.append((name, float(t.abs().max())))                      # utils/save_a.py:22-26

    def _conv(self, x, key, silu=True):
        pre = _SD[key]
        w, b = self.sd[pre + '.weight'], self.sd[pre + '.bias']
        y = F.conv2d(x, w, b, stride=2 if key in STRIDE2 else 1, padding=w.shape[2] // 2)
        self._tap(TAP[key], y)
        if key == self.stop_after:
            raise Stop()
        if silu:
            y = F.silu(y)
        return y

    def _c2f(self, x, name, n_bottle, add):
        x = self._conv(x, f'{name}_conv_0')
        half = x.shape[1] // 2
        parts = [x[:, :half], x[:, half:]]
        cur = x[:, half:]
        for i in range(n_bottle):
            y = self._conv(cur, f'{name}_bottle_{2 * i}')
            y = self._conv(y, f'{name}_bottle_{2 * i + 1}')
            cur = y + cur if add else y
            parts.append(cur)
        return self._conv(torch.cat(parts, 1), f'{name}_conv_1')

    def forward(self, img):
        self.taps = []
        x = torch.as_tensor(np.asarray(img, np.float32))
        try:
            self._tap('start', x)
            x = self._conv(x, 'Conv_P1')
            x = self._conv(x, 'Conv_P2')
            x = self._c2f(x, 'C2F_2', 1, True)
            x = self._conv(x, 'Conv_P3')
            r1 = x = self._c2f(x, 'C2F_4', 2, True)
            x = self._conv(x, 'Conv_P4')
            r2 = x = self._c2f(x, 'C2F_6', 2, True)
            x = self._conv(x, 'Conv_P5')
            x = self._c2f(x, 'C2F_8', 1, True)
            x = self._conv(x, 'SPPF_conv_0')
            p1 = F.max_pool2d(x, 5, 1, 2); p2 = F.max_pool2d(p1, 5, 1, 2); p3 = F.max_pool2d(p2, 5, 1, 2)
            sppf = x = self._conv(torch.cat((x, p1, p2, p3), 1), 'SPPF_conv_1')
            u = F.interpolate(x, scale_factor=2, mode='nearest')
            r4 = x = self._c2f(torch.cat((u, r2), 1), 'C2F_12', 1, False)
            u = F.interpolate(x, scale_factor=2, mode='nearest')
            r5 = x = self._c2f(torch.cat((u, r1), 1), 'C2F_15', 1, False)
            x = self._conv(x, 'Conv_16')
            r6 = x = self._c2f(torch.cat((x, r4), 1), 'C2F_18', 1, False)
            x = self._conv(x, 'Conv_19')
            r7 = self._c2f(torch.cat((x, sppf), 1), 'C2F_21', 1, False)
            for feat, nm in ((r5, 'x_result_5'), (r6, 'x_result_6'), (r7, 'x')):
                for br in ('up', 'down'):
                    t = self._conv(feat, f'{nm}_{br}_0')
                    t = self._conv(t, f'{nm}_{br}_1')
                    self._conv(t, f'{nm}_{br}_2', silu=False)
        except Stop:
            pass
        return self.taps


def parse_max_a_all(text):
    """results/max_a_all.txt (stage_4.py:1007-1011): 'name: [tensor(1.2345), ...]' -> ordered [(name, [floats])]"""
    import re
    out = []
    for line in str(text).splitlines():
        if not line.strip():
            continue
        name, rest = line.split(':', 1)
        out.append((name, [float(v) for v in re.findall(r'tensor\(([-0-9.e+]+)\)', rest)]))
    return out
