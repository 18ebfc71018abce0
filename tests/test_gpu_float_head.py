"""GPU parity of the stage_8_torch.py path (BASELINE configs[0] / configs[1]): integer activations
bit-exact against the reference goldens, the float Detect head within the tolerance stated here, and coord()'s NMS exact
when it is fed the very same prediction tensor.

Tolerances (fp32 CUDA expf / division vs torch's CPU vectorised softmax / sigmoid, ~1e-6 relative per operation):
  class probabilities  rtol 2e-5, atol 1e-7
  boxes (xywh, pixels) atol 2e-3   (a DFL expectation of <= 15 bins times a stride of <= 32)
"""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import synth, yolo_int as Y, float_head as FH

pytestmark = pytest.mark.gpu

PROB_RTOL, PROB_ATOL, BOX_ATOL = 2e-5, 1e-7, 2e-3


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def _setup(golden_dir, taps, impl='tma', max_batch=8):
    from alpha_yolo_quant_b200 import engine, loaders, plan
    K, sd, sc, ma = loaders.load_workload_npz(os.path.join(golden_dir, 'workload_k8.npz'))
    p = plan.compile_plan(sd, sc, ma, K, sigmoid_range=7, taps=taps, head='float')
    e = engine.Engine(p, 0, max_batch, lib_path=None if impl == 'tma' else engine.TEST_LIB_PATH)   # cross-check families: test build only
    e.set_conv_impl(impl)
    return p, e


def _images(seeds):
    return torch.from_numpy(synth.to_input_array([synth.synth_image_u8(s) for s in seeds]))


def _match_dets(d, boxes, classes):
    """same detections in the same order; boxes within BOX_ATOL, confidences within the probability tolerance"""
    assert d.shape[0] == boxes.shape[0], (d.shape, boxes.shape)
    assert np.array_equal(d[:, 5], classes[:, 1])
    np.testing.assert_allclose(d[:, 4], classes[:, 0], rtol=PROB_RTOL, atol=PROB_ATOL)
    np.testing.assert_allclose(d[:, :4], boxes, rtol=0, atol=BOX_ATOL)


@pytest.mark.parametrize('impl', ['dp4a', 'tcgen05', 'tma'])
def test_integer_activations_and_float_head_match_reference_goldens(golden_dir, impl):
    """golden_float_k8.npz was recorded from the unmodified stage_8_torch.py: all 63 conv accumulators and 57 silu outputs
    bit-exact (sigmoid range 7), dbox_cls within tolerance on the recorded anchors, per-anchor conf / class everywhere."""
    g = np.load(os.path.join(golden_dir, 'golden_float_k8.npz'))
    p, e = _setup(golden_dir, taps=True, impl=impl)
    n = 3
    dets, counts, dbc = e.forward(_images(range(n)).cuda(), want_dbox_cls=True)
    torch.cuda.synchronize()
    from alpha_yolo_quant_b200 import plan as P
    acc = [e.export_acc_tap(t, n).cpu().numpy() for t in range(p.n_acc_taps)]
    silu_layers = [nm for nm, _ in P.LAYERS if 'silu_buf' in p.info['layers'][nm]]
    silu = [e.export_buffer(p.info['layers'][nm]['silu_buf'], n).cpu().numpy() for nm in silu_layers]
    assert len(acc) == 63 and len(silu) == 57
    dbc = dbc.cpu().numpy()
    for i in range(n):
        pre = f'img{i}_'
        bad = [P.LAYERS[t][0] for t in range(63) if sha(acc[t][i:i + 1]) != g[pre + 'conv_sha'][t]]
        bad += [silu_layers[t] for t in range(57) if sha(silu[t][i:i + 1]) != g[pre + 'silu_sha'][t]]
        assert not bad, (i, bad[:8])
        ref = g[pre + 'dbox_cls_s8']
        np.testing.assert_allclose(dbc[i, :4, ::8], ref[:4], rtol=0, atol=BOX_ATOL)
        np.testing.assert_allclose(dbc[i, 4:, ::8], ref[4:], rtol=PROB_RTOL, atol=PROB_ATOL)
        np.testing.assert_allclose(dbc[i, 4:].max(0), g[pre + 'conf_max'], rtol=PROB_RTOL, atol=PROB_ATOL)
    e.close()


def test_coord_is_exact_on_the_same_prediction_tensor(golden_dir):
    """ayq_coord_float (the coord() drop-in) on the ORACLE's dbox_cls: the NMS arithmetic is fp32 in torchvision's order,
    so the kept set, its order and every output value are identical -- including ties at conf == 1.0 (stable order)."""
    p, e = _setup(golden_dir, taps=False)
    wl = Y.Workload(os.path.join(golden_dir, 'workload_k8.npz'))
    o = FH.OracleFloatHead(wl)
    seeds = [0, 1, 5, 201]
    ref = o.forward(_images(seeds).numpy())
    pred = torch.from_numpy(o.last['dbox_cls']).cuda()
    dets, counts = e.coord_float(pred)
    for i, (b, c) in enumerate(ref):
        k = int(counts[i])
        assert k == b.shape[0], (i, k, b.shape)
        d = dets[i, :k].cpu().numpy()
        assert np.array_equal(d[:, :4], b) and np.array_equal(d[:, 4:6], c), i
    # degenerate inputs: nothing above the confidence threshold -> count 0; one candidate -> itself
    z = torch.zeros((2, 84, 8400), device='cuda')
    z[1, :4, 17] = torch.tensor([100., 120., 30., 40.])
    z[1, 4 + 7, 17] = 0.5
    dets, counts = e.coord_float(z)
    assert counts.tolist() == [0, 1]
    assert dets[1, 0].tolist() == [85., 100., 115., 140., 0.5, 7.]
    e.close()


def test_end_to_end_detections_match_reference_and_oracle(golden_dir):
    """model(img) through the drop-in module: detections of the recorded reference run (6 golden images) and of the oracle
    on unseen seeds -- same rows in the same order, values within tolerance; batching invariant."""
    from alpha_yolo_quant_b200 import stage_8_torch as S
    g = np.load(os.path.join(golden_dir, 'golden_float_k8.npz'))
    sd = S.configure(workload=os.path.join(golden_dir, 'workload_k8.npz'))
    model = S.Yolov8(max_batch=4).to('cuda')
    model.load_state_dict(sd)
    model.eval()
    n = int(g['n_images'])
    with torch.no_grad():
        res = model.forward_batch(_images(range(n)).cuda())
        b0, c0 = model(_images([2]).cuda())
    for i, (b, c) in enumerate(res):
        _match_dets(torch.cat((b, c), 1).cpu().numpy(), g[f'img{i}_boxes'], g[f'img{i}_classes'])
    assert torch.equal(b0, res[2][0]) and torch.equal(c0, res[2][1])
    wl = Y.Workload(os.path.join(golden_dir, 'workload_k8.npz'))
    o = FH.OracleFloatHead(wl)
    seeds = [301, 302]
    ref = o.forward(_images(seeds).numpy())
    with torch.no_grad():
        res = model.forward_batch(_images(seeds).cuda())
    for (b, c), (rb, rc) in zip(res, ref):
        _match_dets(torch.cat((b, c), 1).cpu().numpy(), rb, rc)
    # the free functions of the module
    pred = torch.from_numpy(o.last['dbox_cls'][:1]).cuda()
    out = S.coord(pred)
    assert np.array_equal(out[0][:, :4].cpu().numpy(), ref[0][0])
    with pytest.raises(Exception):
        S.coord(pred.cpu())
