"""Host logic: the plan compiler against the pinned oracle and the C header (no GPU)."""
import os
import struct

import numpy as np
import pytest

from alpha_yolo_quant_b200 import loaders, lut, plan
from oracle import yolo_int as Y

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module', params=[8, 6, 4])
def compiled(request, golden_dir):
    k = request.param
    path = os.path.join(golden_dir, f'workload_k{k}.npz')
    K, sd, sc, ma = loaders.load_workload_npz(path)
    assert K == k
    return k, plan.compile_plan(sd, sc, ma, K, taps=True), Y.Workload(path), sd


def test_header_constants_in_sync():
    c = plan.header_constants(os.path.join(REPO, 'alpha_yolo_quant_b200', 'csrc', 'plan_format.h'))
    assert c['AYQ_MAGIC'] == plan.MAGIC and c['AYQ_PLAN_VERSION'] == plan.VERSION
    assert c['AYQ_OP_FIELDS'] == plan.OP_FIELDS and c['AYQ_MAX_OUT'] == plan.MAX_OUT
    for name in ('OP_CONV', 'OP_CONV_P1', 'OP_POOL', 'OP_HEAD', 'OP_NMS', 'EPI_SILU', 'EPI_REQUANT8', 'EPI_REQUANT16',
                 'OUT_IDENT', 'OUT_REQUANT'):
        assert c[name] == getattr(plan, name), name
    # field indices used literally in plan.py
    assert (c['CF_KSIZE'], c['CF_NKC'], c['CF_KC_OFF'], c['CF_W_OFF'], c['CF_BIAS_OFF'], c['CF_TAB_OFF']) == (1, 8, 9, 10, 11, 12)
    assert (c['CF_EPI'], c['CF_CLAMP'], c['CF_LUT_OFF'], c['CF_NOUT'], c['CF_OUT0'], c['CF_OUT_STRIDE']) == (13, 14, 15, 16, 17, 6)
    assert (c['CF_LAYER'], c['CF_ACC_TAP'], c['CF_NAME_OFF']) == (40, 41, 42)
    assert c['CF_OUT0'] + c['CF_OUT_STRIDE'] * c['AYQ_MAX_OUT'] <= c['CF_LAYER']
    assert (c['P1_HOUT'], c['P1_OUT_BUF'], c['P1_W_OFF'], c['P1_BIAS_OFF'], c['P1_TAB_OFF'], c['P1_CLAMP'], c['P1_LUT_OFF'], c['P1_ACC_TAP']) == (1, 3, 4, 5, 6, 7, 8, 9)
    assert (c['PL_IN_BUF'], c['PL_NPLANES'], c['PL_OUT_BUF'], c['PL_H'], c['PL_W']) == (1, 3, 4, 6, 7)
    assert (c['HD_BOX_BUF0'], c['HD_CLS_BUF0'], c['HD_LUT_EXP_OFF'], c['HD_LUT16_OFF'], c['HD_DFLW_OFF'], c['HD_ANCH_OFF'], c['HD_KD'], c['HD_ID']) == (1, 4, 7, 8, 9, 10, 11, 12)


def _op_fields(p, i):
    hdr = struct.unpack_from('<IIiiiiii4Q', p.blob, 0)
    ops_off, data_off = hdr[9], hdr[10]
    return struct.unpack_from(f'<{plan.OP_FIELDS}i', p.blob, ops_off + 4 * plan.OP_FIELDS * i), data_off


def test_blob_header(compiled):
    k, p, wl, sd = compiled
    hdr = struct.unpack_from('<IIiiiiii4Q', p.blob, 0)
    assert hdr[0] == plan.MAGIC and hdr[2] == k and hdr[3] == len(p.bufs) and hdr[4] == p.n_ops
    assert hdr[7] == 8400 and hdr[10] + hdr[11] == len(p.blob)
    assert p.n_ops == 63 + 1 + 1 + 1           # convs (Conv_P1 is its own op) + pool + head + nms
    assert p.n_acc_taps == 63


def test_silu_tables_match_oracle(compiled):
    """(k1,s1,k2,s2) of every fused SiLU epilogue == the oracle's coeffs() (which is pinned to the reference's
    requantize() return values through tests/golden coeff_k / coeff_s)."""
    k, p, wl, sd = compiled
    o = Y.OracleYolov8(wl)
    n_checked = 0
    for name, meta in p.info['layers'].items():
        f, data_off = _op_fields(p, meta['op'])
        if f[0] == plan.OP_CONV and f[13] != plan.EPI_SILU:
            continue
        cout = meta['cout']
        tab_off = f[12] if f[0] == plan.OP_CONV else f[6]
        tab = np.frombuffer(p.blob, np.float32, 4 * cout, data_off + tab_off).reshape(4, cout)
        k1, s1 = Y.coeffs(wl.scales[name], Y.scale(6, k))
        assert np.array_equal(tab[0], k1.astype(np.float32)) and np.array_equal(tab[1], np.ldexp(1.0, -s1).astype(np.float32)), name
        n_checked += 1
    assert n_checked == 57


def test_weights_packed_with_residual_duplication(compiled):
    """C2F_4_conv_1 reads [x0, x1, y1+x1, y2+y1+x1]: 7 segments of 2 planes, weights duplicated per addend."""
    k, p, wl, sd = compiled
    meta = p.info['layers']['C2F_4_conv_1']
    f, data_off = _op_fields(p, meta['op'])
    assert f[8] == 14 and meta['kmacs'] * 128 == meta['macs'] * 224
    w = np.frombuffer(p.blob, np.int8, 14 * 64 * 16, data_off + f[10]).reshape(14, 64, 16)
    ref = sd['cf2_conv_3.0.weight'].numpy()[:, :, 0, 0]          # (64,128)
    chunk = lambda c0: ref[:, c0:c0 + 16].astype(np.int8)
    expect = [chunk(0), chunk(16), chunk(32), chunk(48),          # x0, x1
              chunk(64), chunk(80), chunk(64), chunk(80),         # y1 + x1
              chunk(96), chunk(112), chunk(96), chunk(112), chunk(96), chunk(112)]   # y2 + y1 + x1
    for i, e in enumerate(expect):
        assert np.array_equal(w[i], e), i


def test_anchor_table(compiled):
    k, p, wl, sd = compiled
    a, s = plan.quantised_anchors()
    assert a.shape == (8400, 2) and abs(s - 32767 / 79.5) < 1e-3
    assert a[0].tolist() == [206, 206] and a[-1].tolist() == [round(19.5 * s), round(19.5 * s)]


def test_lut_builders_match_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'golden_k8.npz'))
    assert np.array_equal(lut.cached_array('sigmoid', 6, 8)[1].astype(np.int32), g['lut_sigmoid'])
    assert np.array_equal(lut.cached_array('sigmoid', 12, 16)[1].astype(np.int32), g['lut_sigmoid16'])
    assert np.array_equal(lut.cached_array('exp', plan.DFL_RANGE, 8)[1].astype(np.int32), g['lut_exp'])


def test_rescale_retry_is_global():
    """A power-of-two ratio gives k = 256 > 255: the reference then decrements the shift of EVERY channel
    (utils/rescale_coeff_torch.py:27-30), not only of the offending one."""
    import torch
    k, s = plan.rescale_coeffs(torch.tensor([1.0, 1.5]).reshape(1, 2, 1, 1), 1.0)
    assert k.tolist() == [128.0, 85.0] and s.tolist() == [7.0, 7.0]
    k, s = plan.rescale_coeffs(1.0, 1.0)
    assert k.tolist() == [128.0] and s.tolist() == [7.0]


def test_float_head_plan(golden_dir):
    """head='float' (stage_8_torch.py): same 63 convs, six int32 NCHW accumulator buffers, OP_HEAD_FLOAT + OP_NMS_FLOAT."""
    import os
    from alpha_yolo_quant_b200 import loaders, plan
    K, sd, sc, ma = loaders.load_workload_npz(os.path.join(golden_dir, 'workload_k8.npz'))
    p = plan.compile_plan(sd, sc, ma, K, sigmoid_range=7, head='float')
    q = plan.compile_plan(sd, sc, ma, K)
    assert p.n_ops == q.n_ops and p.op_names[-2:] == ['head(float)', 'coord(float NMS)']
    acc = [b for b in p.bufs if b[4] == 4]
    assert [(b[1] * 16, b[2]) for b in acc] == [(64, 80), (80, 80), (64, 40), (80, 40), (64, 20), (80, 20)]
    assert not [b for b in q.bufs if b[4] == 4]
    hc = plan.header_constants(os.path.join(os.path.dirname(plan.__file__), 'csrc', 'plan_format.h'))
    assert hc['OP_HEAD_FLOAT'] == plan.OP_HEAD_FLOAT and hc['OP_NMS_FLOAT'] == plan.OP_NMS_FLOAT and hc['CF_ACC_BUF'] == 43
    assert hc['AYQ_PLAN_VERSION'] == plan.VERSION


def test_algorithmic_bytes_equal_survey_8d(golden_dir):
    """bench.py's roofline numerator: per conv, the reference conv's input read once + output written once at 1 B / element, independent
    of this plan's layout choices (phase-split copies, duplicated addends, 2-byte class logits).  SURVEY.md 8(d): 19,980,800 +
    15,097,600 B over the 63 convs; minus Conv_P1's share (1,228,800 + 1,638,400) = 32,211,200 B for the 62 tcgen05 convs."""
    from alpha_yolo_quant_b200 import loaders, plan
    K, sd, sc, ma = loaders.load_workload_npz(os.path.join(golden_dir, 'workload_k8.npz'))
    p = plan.compile_plan(sd, sc, ma, K)
    L = [v for v in p.info['layers'].values() if 'alg_in_bytes' in v]
    assert len(L) == 62
    assert sum(v['alg_in_bytes'] for v in L) == 19980800 - 1228800
    assert sum(v['alg_out_bytes'] for v in L) == 15097600 - 1638400
    assert sum(v['alg_in_bytes'] + v['alg_out_bytes'] for v in L) == 32211200
    assert sum(v['macs'] for v in p.info['layers'].values() if 'macs' in v) == 4371456000
