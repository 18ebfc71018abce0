"""GPU parity of the calibration forward (stage_4.py:475-946).  Floating point, so with a tolerance:
the CUDA convolution accumulates in a different order than torch's CPU kernels -- taps agree to rtol 2e-4 (abs-max of tensors
whose values are sums of up to 2304 fp32 products), i.e. to the 4 decimals the reference's text format keeps, give or take
one unit in the last place."""
import os

import numpy as np
import pytest
import torch

from oracle import calib_float as C, synth

pytestmark = pytest.mark.gpu
RTOL = 2e-4


def _random_fused_weights(seed=0):
    """a full 127-key fused state_dict with He-normal weights (the real one is 12 MB and lives only in the build container)"""
    from alpha_yolo_quant_b200.stage_8_torch_full_quant import _conv_shapes
    rng = np.random.default_rng(seed)
    sd = {}
    for prefix, cout, cin, k in _conv_shapes():
        sd[prefix + '.weight'] = (rng.standard_normal((cout, cin, k, k)) * (2.0 / (cin * k * k)) ** 0.5).astype(np.float32)
        sd[prefix + '.bias'] = (rng.standard_normal((cout,)) * 0.1).astype(np.float32)
    sd['dfl.weight'] = np.arange(16, dtype=np.float32).reshape(1, 16, 1, 1)
    return sd


def test_first_taps_match_the_reference_run(golden_dir):
    """bnf_head_k8.npz: the reference's fused weights up to Conv_P3 and its own max_a_all.txt for the six calibration images."""
    from alpha_yolo_quant_b200 import calibration as G
    g = np.load(os.path.join(golden_dir, 'bnf_head_k8.npz'))
    ref = C.parse_max_a_all(str(g['max_a_all_txt']))
    sd = {k: g[k] for k in g.files if k.endswith('.weight') or k.endswith('.bias')}
    m = G.CalibrationModel(sd, 'cuda')
    x = torch.from_numpy(synth.to_input_array([synth.synth_image_u8(1000 + i) for i in range(synth.N_CALIB)]))
    maxim_a = {}
    with pytest.raises(KeyError):                                  # the fixture stops after Conv_P3: the next layer's weights are absent
        m.forward(x, maxim_a)
    names = [n for n, _ in ref[:8]]
    assert list(maxim_a)[:8] == names == ['start', 'conv_p1', 'conv_p2', 'conv_0_c2f', 'conv_b_0_c2f', 'conv_b_1_c2f', 'conv_b_2_c2f', 'conv_p3']
    for n, vals in ref[:8]:
        got = [float(v) for v in maxim_a[n]]
        assert len(got) == len(vals) == synth.N_CALIB
        np.testing.assert_allclose(got, vals, rtol=RTOL, atol=6e-5)           # reference text: 4 decimals


def test_all_64_taps_match_the_oracle_and_text_round_trip():
    from alpha_yolo_quant_b200 import calibration as G
    sd = _random_fused_weights()
    x = synth.to_input_array([synth.synth_image_u8(1000), synth.synth_image_u8(7)])
    o = C.CalibOracle(sd)
    ref = [o.forward(x[i:i + 1]) for i in range(2)]
    m = G.CalibrationModel(sd, 'cuda')
    maxim_a = m.forward(torch.from_numpy(x).cuda(), {})
    assert list(maxim_a) == [n for n, _ in ref[0]] and len(maxim_a) == 64
    for j, (n, _) in enumerate(ref[0]):
        np.testing.assert_allclose([float(v) for v in maxim_a[n]], [ref[0][j][1], ref[1][j][1]], rtol=RTOL, atol=1e-6, err_msg=n)
    # batch-1 calls append like the batched call
    m1 = {}
    m.forward(torch.from_numpy(x[:1]).cuda(), m1)
    m.forward(torch.from_numpy(x[1:]).cuda(), m1)
    assert all(torch.equal(torch.stack(m1[n]), torch.stack(maxim_a[n])) for n in maxim_a)
    # stage_4 -> stage_5 text formats
    parsed = G.parse_max_a_all(G.format_max_a_all(maxim_a))
    txt = G.format_max_a(parsed)
    assert txt.startswith('start: 1.0\n') and len(txt.splitlines()) == 64


def test_all_64_taps_match_the_reference_run_and_timing(golden_dir, capsys):
    """The reference's own fused weights (bnf_full_k8.npz) through the CUDA calibration forward: all 64 taps x 6 images against the
    reference's max_a_all.txt (4 decimals in the text, fp32 summation order differs: rtol 2e-4), and the device time of the batch."""
    from alpha_yolo_quant_b200 import calibration as G
    full = np.load(os.path.join(golden_dir, 'bnf_full_k8.npz'))
    head = np.load(os.path.join(golden_dir, 'bnf_head_k8.npz'))
    ref = C.parse_max_a_all(str(head['max_a_all_txt']))
    m = G.CalibrationModel({k: full[k] for k in full.files}, 'cuda')
    x = torch.from_numpy(synth.to_input_array([synth.synth_image_u8(1000 + i) for i in range(synth.N_CALIB)])).cuda()
    maxim_a = m.forward(x, {})
    assert list(maxim_a) == [n for n, _ in ref] and len(maxim_a) == 64
    for n, vals in ref:
        np.testing.assert_allclose([float(v) for v in maxim_a[n]], vals, rtol=RTOL, atol=6e-5, err_msg=n)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    xb = x.repeat(6, 1, 1, 1)[:32].contiguous()
    m.forward(xb, {})
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(3):
        m.forward(xb, {})
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / 3
    with capsys.disabled():
        print(f'\ncalibration forward (fp32 BN-fused network + 64 fused abs-max taps): {ms:.1f} ms per 32 images = {32 / ms * 1e3:.0f} images/s')
