"""Build-container only: runs the reference's own frame builders (utils/coco.py map_from_torch_ann_np / map_from_torch_np) and
its loop body (stage_8_torch.py:1004-1024) on a seeded synthetic loader + canned detections, and stores what they produce
(annotation frame CSV, detection frame CSV, the `ann` / `det` value arrays handed to mean_average_precision_for_boxes) in
tests/golden/golden_validation.npz.  tests/test_validation.py replays the same loader through validation.run().

    python tools/make_validation_golden.py
"""
import io
import os
import sys
import types

import numpy as np
import pandas as pd
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches'):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
sys.modules['matplotlib'].patches = sys.modules['matplotlib.patches']
sys.path.insert(0, '/root/reference/quantisation')
from utils import coco  # noqa: E402  (the unmodified reference module)

from tests.validation_fixture import synthetic_loader, canned_model_outputs  # noqa: E402

ann = pd.DataFrame({'ImageID': [], 'LabelName': [], 'XMin': [], 'XMax': [], 'YMin': [], 'YMax': []})
det = pd.DataFrame({'ImageID': [], 'LabelName': [], 'Conf': [], 'XMin': [], 'XMax': [], 'YMin': [], 'YMax': []})
ann_mass, det_mass, no_pred = [], [], []
outs = canned_model_outputs()
for ind, batch in enumerate(synthetic_loader()):                         # stage_8_torch.py:1004-1013
    boxes, classes = outs[ind]
    ann_mass.append((batch['images'], str(ind), batch['boxes'], batch['categories']))
    if isinstance(boxes, torch.Tensor):
        det_mass.append((str(ind), boxes, classes))
    else:
        no_pred.append(str(ind))
for el in ann_mass:                                                      # :1016-1019
    ann = coco.map_from_torch_ann_np(ann, el[0], el[1], el[2], el[3])
for el in det_mass:
    det = coco.map_from_torch_np(det, el[0], no_pred, el[1].cpu().numpy(), el[2].cpu().numpy(), ann=0)
buf_a, buf_d = io.StringIO(), io.StringIO()
ann.to_csv(buf_a, index=False)
det.to_csv(buf_d, index=False)
a = ann[['ImageID', 'LabelName', 'XMin', 'XMax', 'YMin', 'YMax']].values  # :1023-1024
d = det[['ImageID', 'LabelName', 'Conf', 'XMin', 'XMax', 'YMin', 'YMax']].values
np.savez_compressed(os.path.join(REPO, 'tests', 'golden', 'golden_validation.npz'), ann_csv=np.array(buf_a.getvalue()),
                    det_csv=np.array(buf_d.getvalue()), no_pred=np.array(no_pred), ann_values=np.array(a.astype(str)),
                    det_values=np.array(d.astype(str)), versions=np.array(f'pandas {pd.__version__} numpy {np.__version__}'))
print('wrote golden_validation.npz:', len(ann), 'annotation rows,', len(det), 'detection rows, no_pred', no_pred)
