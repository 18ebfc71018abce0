"""Detections -> map_boxes frames: identical to what the reference's map_from_torch_np builds
(golden_det_frame.npz: CSV written by the reference function for five images, one of them without detections)."""
import io
import os

import numpy as np
import pandas as pd


def test_frame_matches_reference_rows(golden_dir):
    from alpha_yolo_quant_b200 import validation as V
    g = np.load(os.path.join(golden_dir, 'golden_det_frame.npz'))
    assert list(g['names']) == list(V.COCO_NAMES)
    n = 5
    dets = np.zeros((n, 300, 6), np.float32)
    counts = np.zeros((n,), np.int32)
    for i in range(n):
        b, c = g[f'boxes{i}'], g[f'classes{i}']
        counts[i] = b.shape[0]
        dets[i, :b.shape[0], :4] = b
        dets[i, :b.shape[0], 4:6] = c
    det, no_pred = V.append(V.empty_frame(), dets, counts, first_index=0, no_pred=[])
    assert no_pred == list(g['no_pred'])
    ref = pd.read_csv(io.StringIO(str(g['csv'])), dtype={'ImageID': str})
    mine = pd.read_csv(io.StringIO(det.to_csv(index=False)), dtype={'ImageID': str})
    assert list(mine.columns) == list(ref.columns) == V.COLUMNS
    assert det.to_csv(index=False) == str(g['csv'])                       # byte-identical CSV
    # appending batch by batch equals one call
    d2, np2 = V.append(V.empty_frame(), dets[:2], counts[:2], 0, [])
    d2, np2 = V.append(d2, dets[2:], counts[2:], 2, np2)
    assert d2.to_csv(index=False) == det.to_csv(index=False) and np2 == no_pred
