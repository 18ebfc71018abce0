"""GPU parity of the weight quantiser through the C ABI: bit-exact against the reference's conv_quant()
results (golden_wquant_k8.npz) and against the oracle on random tensors, including an all-zero channel and other bit widths."""
import os

import numpy as np
import pytest
import torch

from oracle import weight_quant as WQ

pytestmark = pytest.mark.gpu


def test_conv_quant_matches_reference_goldens(golden_dir):
    from alpha_yolo_quant_b200 import weight_quant as G
    g = np.load(os.path.join(golden_dir, 'golden_wquant_k8.npz'))
    for l in g['layers']:
        q, b, s = G.conv_quant(str(l), torch.from_numpy(g[f'{l}/w']).cuda(), torch.from_numpy(g[f'{l}/b']).cuda(),
                               float(g[f'{l}/scale_input']), bool(g[f'{l}/start']), k=8)
        assert np.array_equal(q.cpu().numpy(), g[f'{l}/qw'].astype(np.int64)), l
        assert np.array_equal(b.cpu().numpy(), g[f'{l}/qb']), l
        assert np.array_equal(s.cpu().numpy(), g[f'{l}/scale_res']), l


@pytest.mark.parametrize('k', [8, 6, 4])
def test_conv_quant_matches_oracle_on_random_tensors(k):
    from alpha_yolo_quant_b200 import weight_quant as G
    rng = np.random.default_rng(k)
    for shape in ((5, 3, 3, 3), (48, 96, 1, 1), (256, 128, 3, 3)):
        w = (rng.standard_normal(shape) * rng.uniform(0.01, 3.0, (shape[0], 1, 1, 1))).astype(np.float32)
        b = rng.standard_normal((shape[0], 1, 1, 1)).astype(np.float32)
        si = float(rng.uniform(1.0, 40.0))
        q, qb, s = G.conv_quant('rand', torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda(), si, False, k=k)
        rq, rb, rs = WQ.conv_quant(w, b, si, k, False)
        assert np.array_equal(q.cpu().numpy(), rq) and np.array_equal(qb.cpu().numpy(), rb) and np.array_equal(s.cpu().numpy(), rs), shape
    w = torch.zeros((2, 4, 1, 1), device='cuda'); w[1] = 0.5
    q, qb, s = G.conv_quant('zero', w, torch.ones((2, 1, 1, 1), device='cuda'), 3.0, False, k=k)
    assert q[0].abs().sum().item() == 0 and int(qb[0, 0, 0, 0]) == 0 and torch.isinf(s[0, 0, 0, 0])
    with pytest.raises(Exception):
        G.conv_quant('cpu', w.cpu(), torch.ones((2, 1, 1, 1)), 3.0)
