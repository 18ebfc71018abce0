"""Pins the calibration-forward oracle (oracle/calib_float.py; stage_4.py:475-946) to the unmodified reference in CI: ALL 64 abs-max
taps of ALL six calibration images against the reference's own max_a_all.txt (4 decimals), with the reference's BN-fused weights
(tests/golden/bnf_full_k8.npz, written by tools/pin_calib_oracle.py from the harness run; the text lives in bnf_head_k8.npz)."""
import os

import numpy as np

from oracle import calib_float as C, synth


def test_all_64_taps_equal_the_reference_text(golden_dir):
    full = np.load(os.path.join(golden_dir, 'bnf_full_k8.npz'))
    head = np.load(os.path.join(golden_dir, 'bnf_head_k8.npz'))
    ref = C.parse_max_a_all(str(head['max_a_all_txt']))
    assert len(ref) == 64
    o = C.CalibOracle({k: full[k] for k in full.files})
    mine = {}
    for i in range(synth.N_CALIB):
        for nm, v in o.forward(synth.to_input_array([synth.synth_image_u8(1000 + i)])):
            mine.setdefault(nm, []).append(v)
    assert [n for n, _ in ref] == list(mine.keys())
    worst = max(abs(round(a, 4) - b) for n, vals in ref for a, b in zip(mine[n], vals))
    assert worst <= 1.01e-4, worst
    # stage_5: per-tap maximum over the images = the reference's max_a.txt
    from alpha_yolo_quant_b200 import calibration as G
    from alpha_yolo_quant_b200.plan import parse_max_a
    want = parse_max_a(str(head['max_a_txt']))
    got = parse_max_a(G.format_max_a({n: [round(float(v), 4) for v in vals] for n, vals in mine.items()}))
    assert got.keys() == want.keys()
    assert max(abs(got[k] - want[k]) for k in want) <= 1.01e-4
