"""Accuracy loop glue: engine detections -> the `map_boxes` frames the reference's validation driver
builds (/root/reference/quantisation/stage_8_torch.py:1004-1037, utils/coco.py:138-158 map_from_torch_np), so that the mAP of
this engine can be compared with the reference's when COCO and the `map_boxes` package are available.

The reference appends one image at a time to a DataFrame (`pd.concat` per image, quadratic); here a whole batch of engine
output (dets (n,300,6), counts (n)) becomes the same rows in one go.  Same columns, same order, same values:
XMin / YMin / XMax / YMax = pixel box / 640 in float32 (as the reference divides its float32 arrays), Conf = the confidence
column, LabelName from the COCO-80 table, ImageID the image index as a string; images without detections go to `no_pred`.
"""
import numpy as np
import pandas as pd

# class index -> name, the table of utils/coco.py:17-98 (COCO-80 in ultralytics order); pinned by tests/test_validation.py
COCO_NAMES = (
    'person', 'bicycle', 'car', 'motorcycle', 'airplane', 'bus', 'train', 'truck', 'boat', 'traffic light', 'fire hydrant',
    'stop sign', 'parking meter', 'bench', 'bird', 'cat', 'dog', 'horse', 'sheep', 'cow', 'elephant', 'bear', 'zebra',
    'giraffe', 'backpack', 'umbrella', 'handbag', 'tie', 'suitcase', 'frisbee', 'skis', 'snowboard', 'sports ball', 'kite',
    'baseball bat', 'baseball glove', 'skateboard', 'surfboard', 'tennis racket', 'bottle', 'wine glass', 'cup', 'fork',
    'knife', 'spoon', 'bowl', 'banana', 'apple', 'sandwich', 'orange', 'broccoli', 'carrot', 'hot dog', 'pizza', 'donut',
    'cake', 'chair', 'couch', 'potted plant', 'bed', 'dining table', 'toilet', 'tv', 'laptop', 'mouse', 'remote', 'keyboard',
    'cell phone', 'microwave', 'oven', 'toaster', 'sink', 'refrigerator', 'book', 'clock', 'vase', 'scissors', 'teddy bear',
    'hair drier', 'toothbrush')
COLUMNS = ['ImageID', 'LabelName', 'Conf', 'XMin', 'XMax', 'YMin', 'YMax']


def empty_frame():
    """stage_8_torch.py:999: the empty `det` frame the loop starts from"""
    return pd.DataFrame({c: [] for c in COLUMNS})


def detections_to_frame(dets, counts, first_index=0, no_pred=None, w=640, h=640):
    """dets (n,300,6) rows [x1,y1,x2,y2,conf,class], counts (n) -- torch (any device) or numpy.  Returns the rows that n calls
    of map_from_torch_np(det, str(ind), no_pred, boxes, classes, ann=0) append for ind = first_index .. first_index+n-1."""
    d = dets.detach().cpu().numpy() if hasattr(dets, 'detach') else np.asarray(dets)
    c = counts.detach().cpu().numpy() if hasattr(counts, 'detach') else np.asarray(counts)
    no_pred = no_pred if no_pred is not None else []
    rows, ids = [], []
    for i in range(d.shape[0]):
        k = int(c[i])
        if k == 0:
            no_pred.append(str(first_index + i))                   # :156-157
            continue
        rows.append(d[i, :k].astype(np.float32))
        ids += [str(int(first_index + i))] * k
    if not rows:
        return empty_frame(), no_pred
    r = np.concatenate(rows, 0)
    out = pd.DataFrame({'XMin': r[:, 0] / np.float32(w), 'YMin': r[:, 1] / np.float32(h), 'XMax': r[:, 2] / np.float32(w),
                        'YMax': r[:, 3] / np.float32(h)})                                    # :144-148 (float32 arrays)
    out['ImageID'] = ids
    out['LabelName'] = [COCO_NAMES[int(j)] for j in r[:, 5]]
    out['Conf'] = r[:, 4]
    return out, no_pred


def append(det_frame, dets, counts, first_index=0, no_pred=None):
    """det = map_from_torch_np(det, ...) for a whole batch: pd.concat([det, new rows], ignore_index=True) (:154)"""
    new, no_pred = detections_to_frame(dets, counts, first_index, no_pred)
    if len(new) == 0:
        return det_frame, no_pred
    return pd.concat([det_frame, new], ignore_index=True), no_pred


def mean_ap(ann, det_frame, thresholds=np.arange(0.5, 1, 0.05)):
    """stage_8_torch.py:1026-1035 (needs the external `map_boxes` package, absent from this image: raises ImportError then)"""
    from map_boxes import mean_average_precision_for_boxes
    det = det_frame[COLUMNS].values
    res = [mean_average_precision_for_boxes(ann, det, round(float(t), 2))[0] for t in thresholds]
    return sum(res) / len(res), res
