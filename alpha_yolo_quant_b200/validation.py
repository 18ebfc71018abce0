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


# ----------------------------------------------------------------------------- annotation frame + driver loop
ANN_COLUMNS = ['ImageID', 'LabelName', 'XMin', 'XMax', 'YMin', 'YMax']


def empty_ann_frame():
    """stage_8_torch.py:998"""
    return pd.DataFrame({c: [] for c in ANN_COLUMNS})


def annotations_to_frame(orig_img_shape, ind, boxes, categories):
    """Rows map_from_torch_ann_np(ann, orig_img, str(ind), boxes, categories) appends (utils/coco.py:226-245): `boxes` is the
    loader's (1, m, 4) COCO xywh tensor in ORIGINAL-image pixels, `categories` (1, m); XMax / YMax = (x + w) / W, (y + h) / H in
    the dtype of the box array, XMin / YMin = x / W, y / H; W, H = orig_img.shape[3], orig_img.shape[2]."""
    b = (boxes[0].numpy() if hasattr(boxes, 'numpy') else np.asarray(boxes)[0]).copy()
    c = categories.numpy() if hasattr(categories, 'numpy') else np.asarray(categories)
    w, h = orig_img_shape[3], orig_img_shape[2]
    b[:, 2] = (b[:, 0].copy() + b[:, 2]) / w
    b[:, 3] = (b[:, 1].copy() + b[:, 3]) / h
    b[:, 0] = b[:, 0] / w
    b[:, 1] = b[:, 1] / h
    out = pd.DataFrame(b, columns=['XMin', 'YMin', 'XMax', 'YMax'])
    out['ImageID'] = str(int(ind))
    out['LabelName'] = [COCO_NAMES[int(c[0][i])] for i in range(b.shape[0])]
    return out


def run(loader, model, main_dir=None, K=8, batch_images=64, resize=None, map_fn=None, write=True, progress=None):
    """The validation driver of stage_8_torch.py:984-1037, re-hosted on the batched engine.

    loader   iterable of {'images': float (1,3,H,W) in [0,1] (ToTensor of the original image), 'boxes': (1,m,4) COCO xywh,
             'categories': (1,m)} -- what `val_dataset.pytorch(batch_size=1, transform=...)` yields (:996)
    model    the drop-in Yolov8 (forward_batch) or anything with forward_batch(x (n,3,640,640)) -> list of (boxes, classes)
    resize   batch_transform of :991-993 (images -> 640x640); default torchvision-free bilinear `interpolate(antialias=True)`,
             which is what transforms.Resize does on a float tensor; pass None-op for already 640x640 inputs
    map_fn   mean_average_precision_for_boxes(ann, det, iou) of the external `map_boxes` package (absent from this image:
             importable or injected; when neither, the mAP part is skipped and None is returned for it)

    The reference calls model(img) once per image (batch 1, :1007); here `batch_images` resized images are stacked and run as one
    engine batch -- element i of forward_batch equals model(x[i:i+1]) (tests/test_gpu_parity.py).  Returns a dict with the
    annotation frame, detection frame, no_pred list, per-threshold APs, their mean (:1035) and the CSV path (:1021).
    """
    import torch
    if resize is None:
        def resize(img):
            if tuple(img.shape[-2:]) == (640, 640):
                return img
            return torch.nn.functional.interpolate(img, size=(640, 640), mode='bilinear', antialias=True, align_corners=False)
    ann, det, no_pred = empty_ann_frame(), empty_frame(), []
    ann_rows = []
    pend, pend_idx = [], []

    def flush():
        nonlocal det, no_pred
        if not pend:
            return
        res = model.forward_batch(torch.cat(pend, 0))
        dets = np.zeros((len(res), 300, 6), np.float32)
        counts = np.zeros((len(res),), np.int32)
        for j, (boxes, classes) in enumerate(res):
            if isinstance(boxes, torch.Tensor):                    # :1009-1012 (else: no_pred)
                k = boxes.shape[0]
                counts[j] = k
                dets[j, :k, :4] = boxes.detach().cpu().numpy()
                dets[j, :k, 4:6] = classes.detach().cpu().numpy()
        det, no_pred = append(det, dets, counts, first_index=pend_idx[0], no_pred=no_pred)   # :1018-1019 for the whole batch
        pend.clear(); pend_idx.clear()

    it = loader if progress is None else progress(loader)
    n_images = 0
    for ind, batch in enumerate(it):
        img = resize(batch['images'].float())
        pend.append(img); pend_idx.append(ind)
        ann_rows.append(annotations_to_frame(tuple(batch['images'].shape), ind, batch['boxes'], batch['categories']))   # :1008, :1017
        n_images += 1
        if len(pend) >= batch_images:
            flush()
    flush()
    if ann_rows:
        ann = pd.concat([ann] + ann_rows, ignore_index=True)
    out = {'ann': ann, 'det': det, 'no_pred': no_pred, 'n_images': n_images, 'csv': None, 'aps': None, 'map': None}
    main_dir = main_dir if main_dir is not None else f'{K}_nano'
    if write:
        import os
        os.makedirs(os.path.join(main_dir, 'results'), exist_ok=True)
        out['csv'] = os.path.join(main_dir, 'results', f'det_QUANT_{K}_channel.csv')          # :1021
        det.to_csv(out['csv'], index=False)
    if map_fn is None:
        try:
            from map_boxes import mean_average_precision_for_boxes as map_fn
        except ImportError:
            map_fn = None
    if map_fn is not None:
        a = ann[ANN_COLUMNS].values                                                          # :1023-1024
        d = det[COLUMNS].values
        aps = []
        for iou_threshold in np.arange(0.5, 1, 0.05):                                        # :1026-1029
            mean_ap_t, _ = map_fn(a, d, round(iou_threshold, 2))
            aps.append(mean_ap_t)
        out['aps'] = aps
        out['map'] = sum(aps) / len(aps)                                                     # :1031
        if write:
            write_run_result(out['map'], main_dir)
    return out


def write_run_result(mAP, main_dir, comments='Default'):
    """utils/write_run_result.py:6-22, stage == 7 branch: appends to {main_dir}/results/runs_val/results.txt"""
    import os
    from datetime import datetime
    cur = datetime.now()
    os.makedirs(os.path.join(main_dir, 'results', 'runs_val'), exist_ok=True)
    with open(os.path.join(main_dir, 'results', 'runs_val', 'results.txt'), 'a') as f:
        f.write(f'DATE: {cur.date().day}.{cur.date().month}.{cur.date().year} '
                f'TIME: {cur.time().hour}:{cur.time().minute}:{cur.time().second}\n')
        f.write(f'Comments: {comments}\n')
        f.write(f'QUANT MODEL mAP(.50 - .95): {mAP}\n')
        f.write('---------------\n')
        f.write('\n')
