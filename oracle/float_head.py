"""CPU oracle of the FLOAT Detect head -- TEST INFRASTRUCTURE (see oracle/yolo_int.py for who may import it).

Restates the tail of /root/reference/quantisation/stage_8_torch.py: the integer backbone /
neck / head convolutions are those of oracle/yolo_int.py with the sigmoid table range 7 (stage_8_torch.py:264,268),
then the six raw head accumulators are dequantised and decoded in floating point:
    stage_8_torch.py:915-922   x / scale                     (per-channel scales of all_scales[...])
    :926-941                   softmax over the 16 DFL bins, self.dfl (weights arange(16)), dist2bbox * strides
    :944-947                   class logits -> sigmoid, dbox_cls = cat(dbox, cls)
    :146-190  coord()          xywh2xyxy, conf > 1e-8, per-anchor max / argmax, class offsets 7680, NMS 0.45, [:300]
    :203-258, :949-957         scale_boxes (gain 1, pad 0) + clip_boxes to [0, 640], convert_res
The NMS is torchvision.ops.nms (not under /root/reference; requirements.txt pins torchvision==0.18.0): greedy over boxes
in descending score order, box j is dropped when inter / (area_i + area_j - inter) > thr, all in fp32.  nms_greedy below
restates that published algorithm; tests/test_oracle_golden.py pins it (and everything else here) against
tests/golden/golden_float_k8.npz recorded from the unmodified reference (oracle/ref_harness.py --float-head).

Floating point: torch CPU ops are used for softmax / sigmoid (the reference calls the same ops), so on the recording
machine the result is bit-identical; the CUDA path is compared with a tolerance stated in tests/test_gpu_float_head.py.
"""
import numpy as np
import torch

from . import yolo_int as Y

F32 = np.float32


def make_anchors_float():
    """make_anchors stage_8_torch.py:97-109 for the three 640x640 levels: anchor (2, 8400), strides (8400,)"""
    pts, st = [], []
    for hw, s in ((80, 8.), (40, 16.), (20, 32.)):
        sx = np.arange(hw, dtype=F32) + F32(0.5)
        yy, xx = np.meshgrid(sx, sx, indexing='ij')
        pts.append(np.stack((xx.reshape(-1), yy.reshape(-1)), 0))
        st.append(np.full((hw * hw,), s, F32))
    return np.concatenate(pts, 1).astype(F32), np.concatenate(st)


def decode_float(box_acc, cls_acc, dfl_weight):
    """box_acc / cls_acc: three (acc int64 (N,C,H,W), scale fp32 (C,)) pairs.  Returns dbox_cls (N,84,8400) fp32."""
    n = box_acc[0][0].shape[0]
    deq = lambda a, s: torch.from_numpy(a.astype(F32)) / torch.from_numpy(np.asarray(s, F32)).reshape(1, -1, 1, 1)   # :915-922
    box = torch.cat([deq(a, s).reshape(n, 64, -1) for a, s in box_acc], 2)                 # :926
    a = box.shape[2]
    box = box.view(n, 4, 16, a).transpose(2, 1).softmax(1)                                  # :930
    w = torch.from_numpy(np.asarray(dfl_weight, F32)).reshape(1, 16, 1, 1)
    dfl = torch.nn.functional.conv2d(box, w).view(n, 4, a)                                  # self.dfl(box) :933
    anchor, strides = make_anchors_float()
    anchor = torch.from_numpy(anchor).unsqueeze(0)
    lt, rb = dfl.chunk(2, 1)                                                                # dist2bbox :112-121
    x1y1 = anchor - lt
    x2y2 = anchor + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * torch.from_numpy(strides)       # :936
    cls = torch.cat([deq(a_, s).reshape(n, 80, -1) for a_, s in cls_acc], 2).sigmoid()      # :939-940
    return torch.cat((dbox, cls), 1).numpy()                                                # :942


def nms_greedy(boxes, scores, thr):
    """torchvision.ops.nms, CPU kernel semantics: stable descending score order; suppress j when IoU > thr (fp32)."""
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = ((x2 - x1).astype(F32) * (y2 - y1).astype(F32)).astype(F32)
    order = np.argsort(-scores, kind='stable')
    keep = []
    while order.size:
        i = order[0]
        keep.append(i)
        rest = order[1:]
        w = np.maximum(F32(0), (np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest])).astype(F32))
        h = np.maximum(F32(0), (np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest])).astype(F32))
        inter = (w * h).astype(F32)
        ovr = (inter / ((areas[i] + areas[rest]).astype(F32) - inter).astype(F32)).astype(F32)
        order = rest[~(ovr > F32(thr))]
    return np.array(keep, dtype=np.int64)


def coord_one(pred, conf_thres=0.00000001, iou_thres=0.45, max_det=300):
    """coord() :146-190 + scale_boxes / clip_boxes / convert_res for ONE image.  pred (84, 8400) fp32.
    Returns (boxes (n,4), classes (n,2)) fp32, or (None, None) when no anchor passes conf_thres."""
    cx, cy, w, h = pred[:4]
    dw = (w / F32(2)).astype(F32)
    dh = (h / F32(2)).astype(F32)
    xyxy = np.stack((cx - dw, cy - dh, cx + dw, cy + dh), 1).astype(F32)                   # xywh2xyxy :124-143
    cls = pred[4:]
    conf = cls.max(0)
    j = cls.argmax(0)
    sel = np.nonzero(conf > F32(conf_thres))[0]                                             # :150,:167,:172
    if sel.size == 0:
        return None, None
    bx, cf, jj = xyxy[sel], conf[sel], j[sel].astype(F32)
    off = (jj * F32(7680)).astype(F32)[:, None]                                             # :182
    keep = nms_greedy((bx + off).astype(F32), cf, iou_thres)[:max_det]                      # :186-188
    out = np.clip(bx[keep], F32(0), F32(640))                                               # gain 1, pad 0, clip_boxes :240-252
    return out.astype(F32), np.stack((cf[keep], jj[keep]), 1).astype(F32)


class OracleFloatHead:
    """stage_8_torch.Yolov8.forward, batched like OracleYolov8 (element i == reference model(img[i:i+1]))."""

    def __init__(self, wl):
        self.int_model = Y.OracleYolov8(wl, sigmoid_range=7)
        self.wl = wl

    def forward(self, img, trace=False):
        box_acc, cls_acc = self.int_model.forward_maps(img, trace, raw_head=True)
        pred = decode_float(box_acc, cls_acc, self.wl.sd['dfl.weight'].reshape(-1))
        self.last = dict(dbox_cls=pred, box_acc=box_acc, cls_acc=cls_acc)
        return [coord_one(pred[i]) for i in range(pred.shape[0])]
