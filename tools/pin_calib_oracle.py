"""Build-container check of oracle/calib_float.py against the UNMODIFIED reference's stage_4 results (all 64 taps, all
calibration images): needs the harness work directory (oracle/ref_harness.py, default /tmp/ayq_work/k8) with the full
BN-fused weights (12 MB, not committed) and results/max_a_all.txt.  Also (re)writes the small committed fixture
tests/golden/bnf_head_k8.npz = the fused weights of the layers up to Conv_P3 + the reference's max_a_all.txt / max_a.txt text.

    python tools/pin_calib_oracle.py [/tmp/ayq_work/k8]
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import calib_float as C, synth  # noqa: E402

work = sys.argv[1] if len(sys.argv) > 1 else '/tmp/ayq_work/k8'
res = os.path.join(work, '8_nano', 'results')
sd = {k: v.numpy() for k, v in torch.load(os.path.join(res, 'weights_batchnf.pickle')).items()}
txt = open(os.path.join(res, 'max_a_all.txt')).read()
ref = C.parse_max_a_all(txt)
o = C.CalibOracle(sd)
mine = {}
for i in range(synth.N_CALIB):
    for nm, v in o.forward(synth.to_input_array([synth.synth_image_u8(1000 + i)])):
        mine.setdefault(nm, []).append(v)
assert [n for n, _ in ref] == list(mine.keys()), 'tap names / order differ'
worst = max(abs(round(a, 4) - b) for n, vals in ref for a, b in zip(mine[n], vals))
print(f'{len(ref)} taps x {synth.N_CALIB} images: worst |round4(oracle) - reference| = {worst:g}')
assert worst <= 1.01e-4
head = ('conv0.0', 'conv1.0', 'cf2_conv_0.0', 'cf2_bottle_0.0', 'cf2_bottle_0.2', 'cf2_conv_1.0', 'conv3.0')
np.savez_compressed(os.path.join(REPO, 'tests', 'golden', 'bnf_head_k8.npz'),
                    **{f'{p}.{s}': sd[f'{p}.{s}'] for p in head for s in ('weight', 'bias')},
                    max_a_all_txt=np.array(txt), max_a_txt=np.array(open(os.path.join(res, 'max_a.txt')).read()))
print('wrote tests/golden/bnf_head_k8.npz')
# the FULL fused state_dict (127 tensors, 12 MB of fp32): pins all 64 taps in CI (tests/test_calibration_oracle.py) and on the GPU
np.savez_compressed(os.path.join(REPO, 'tests', 'golden', 'bnf_full_k8.npz'), **{k: v for k, v in sd.items()})
print('wrote tests/golden/bnf_full_k8.npz', os.path.getsize(os.path.join(REPO, 'tests', 'golden', 'bnf_full_k8.npz')), 'bytes')
