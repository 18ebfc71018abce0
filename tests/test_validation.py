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


class _CannedModel:
    """forward_batch stand-in returning the canned per-image outputs in order (the engine itself is covered by the GPU tests)."""

    def __init__(self, outs):
        self.outs, self.pos, self.batches = outs, 0, []

    def forward_batch(self, x):
        assert tuple(x.shape[1:]) == (3, 640, 640)
        n = x.shape[0]
        self.batches.append(n)
        r = self.outs[self.pos:self.pos + n]
        self.pos += n
        return r


def test_validation_driver_equals_reference_loop(golden_dir, tmp_path):
    """validation.run == the reference's loop (stage_8_torch.py:1004-1037) on the same loader: annotation frame, detection
    frame, no_pred, the CSV it writes and the arrays handed to mean_average_precision_for_boxes (golden_validation.npz was
    written by tools/make_validation_golden.py with the reference's own utils/coco.py functions)."""
    from alpha_yolo_quant_b200 import validation as V
    from tests.validation_fixture import synthetic_loader, canned_model_outputs
    g = np.load(os.path.join(golden_dir, 'golden_validation.npz'))
    calls = []

    def fake_map(ann, det, iou):                                   # stands in for map_boxes.mean_average_precision_for_boxes
        calls.append((ann.copy(), det.copy(), iou))
        return 0.1 * len(calls), {}
    for batch_images in (1, 3, 64):
        calls.clear()
        m = _CannedModel(canned_model_outputs())
        out = V.run(synthetic_loader(), m, main_dir=str(tmp_path / f'b{batch_images}'), K=8, batch_images=batch_images, map_fn=fake_map)
        assert sum(m.batches) == 7 and max(m.batches) <= batch_images
        assert out['ann'].to_csv(index=False) == str(g['ann_csv'])
        assert out['det'].to_csv(index=False) == str(g['det_csv'])
        assert out['no_pred'] == list(g['no_pred'])
        assert open(out['csv']).read() == str(g['det_csv']) and out['csv'].endswith('results/det_QUANT_8_channel.csv')
        assert [round(c[2], 2) for c in calls] == [0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85, 0.9, 0.95]
        assert np.array_equal(calls[0][0].astype(str), g['ann_values']) and np.array_equal(calls[0][1].astype(str), g['det_values'])
        assert abs(out['map'] - sum(0.1 * (i + 1) for i in range(10)) / 10) < 1e-12
        txt = open(tmp_path / f'b{batch_images}' / 'results' / 'runs_val' / 'results.txt').read()
        assert 'QUANT MODEL mAP(.50 - .95): ' in txt and txt.endswith('---------------\n\n')


def test_validation_driver_without_map_boxes(tmp_path):
    from alpha_yolo_quant_b200 import validation as V
    from tests.validation_fixture import synthetic_loader, canned_model_outputs
    out = V.run(synthetic_loader(), _CannedModel(canned_model_outputs()), main_dir=str(tmp_path), write=False)
    assert out['map'] is None and out['csv'] is None and out['n_images'] == 7 and len(out['det']) == 18
