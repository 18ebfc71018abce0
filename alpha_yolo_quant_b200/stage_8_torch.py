"""Drop-in for the names that /root/reference/quantisation/stage_8_torch.py defines and its driver code uses
(BASELINE configs[0] / configs[1]): the same integer backbone / neck / head convolutions
as the full-quant model (sigmoid table range 7, :264,268), then the FLOAT Detect head -- dequantise the six raw
accumulators, softmax-DFL decode, class sigmoid (:915-947) -- and coord() = confidence filter + class-offset
torchvision NMS at IoU 0.45 (:146-190), scale_boxes / clip_boxes / convert_res (:203-258, :949-957).

Reference usage (stage_8_torch.py:964-1010)                        this module
    model = Yolov8().to(device)                                     same
    model.load_state_dict(torch.load(QUANT_WEIGHTS))                same (127 keys)
    boxes, classes = model(img)        # img (1,3,640,640)          same; + model.forward_batch(x)

Every integer activation is bit-exact with the reference; the float tail runs in fp32 CUDA kernels
(head_float_kernel, nms_float_kernel behind include/ayq.h) and agrees with the reference's CPU torch ops to the
tolerance stated in tests/test_gpu_float_head.py.  No CPU path: non-CUDA tensors raise.
"""
import torch

from . import engine as _eng
from . import loaders as _loaders
from . import stage_8_torch_full_quant as _fq
from .lut import scale, create_sigmoid_lookup_table  # noqa: F401  (reference names)
from .loaders import load_scales, max_a  # noqa: F401
from .stage_8_torch_full_quant import requantize, sigmoid_quant, quant_matrix  # noqa: F401  (same functions, :27-41)

K = 8                                    # stage_0.py:7
MAIN_DIR_NAME = f'{K}_nano'              # stage_0.py:14
SIGMOID_RANGE = 7                        # :264
device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')       # :38
all_scales = {}
max_a_dict = {}
lookup = {}
_state = {'configured': False}
_engines = {}


def configure(main_dir=None, k=8, workload=None):
    """Fill the module globals the reference builds at import (:262-264)."""
    global K, MAIN_DIR_NAME, all_scales, max_a_dict, lookup
    sd = None
    if workload is not None:
        K, sd, all_scales, max_a_dict = _loaders.load_workload_npz(workload)
    else:
        K = int(k)
        MAIN_DIR_NAME = main_dir if main_dir is not None else f'{K}_nano'
        all_scales = load_scales(MAIN_DIR_NAME)
        max_a_dict = max_a(f'{MAIN_DIR_NAME}/results/max_a.txt')
    lookup = create_sigmoid_lookup_table(SIGMOID_RANGE, K)
    _state['configured'] = True
    return sd


def silu(x, scale_x, a_input):
    """stage_8_torch.py:267-281 (sigmoid range 7).  Returns (tensor, scale(a_input, K))."""
    return _fq._silu_impl(x, scale_x, a_input, SIGMOID_RANGE, K, lookup)


def make_anchors(feats, strides, grid_cell_offset=0.5):
    """:97-109: anchor centres (2, A) and their strides (1, A) for the given feature maps (torch plumbing on the caller's
    device; the engine's head kernel derives the same anchors from the anchor index)."""
    pts, strd = [], []
    for f, st in zip(feats, strides):
        h, w = f.shape[-2:]
        ys = torch.arange(h, device=f.device, dtype=f.dtype) + grid_cell_offset
        xs = torch.arange(w, device=f.device, dtype=f.dtype) + grid_cell_offset
        pts.append(torch.stack((xs.repeat(h), ys.repeat_interleave(w)), 0))          # row-major: x fastest
        strd.append(torch.full((1, h * w), float(st), device=f.device, dtype=f.dtype))
    return torch.cat(pts, 1), torch.cat(strd, 1)


def dist2bbox(distance, anchor_points, xywh=True, dim=-1):
    """:112-121: (left, top, right, bottom) distances around the anchor -> xywh (default) or xyxy boxes"""
    near, far = torch.chunk(distance, 2, dim)
    lo, hi = anchor_points - near, anchor_points + far
    return torch.cat(((lo + hi) / 2, hi - lo), dim) if xywh else torch.cat((lo, hi), dim)


def xywh2xyxy(x):
    """:124-143: centre / size -> corners along the last dimension"""
    if x.shape[-1] != 4:
        raise AssertionError(f"input shape last dimension expected 4 but input shape is {x.shape}")
    centre, half = x[..., :2], x[..., 2:] / 2
    return torch.cat((centre - half, centre + half), -1)


def clip_boxes(boxes, shape):
    """:240-258: clamp xyxy boxes to the image (shape = (C, H, W)), in place like the reference"""
    boxes[..., 0::2] = boxes[..., 0::2].clamp(0, shape[2])
    boxes[..., 1::2] = boxes[..., 1::2].clamp(0, shape[1])
    return boxes


def scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=None, padding=True, xywh=False):
    """:203-236: undo the letter-box (gain, pad) of img1 -> img0 and clip; for the 640x640 path gain = 1 and pad = (0, 0)"""
    if ratio_pad is not None:
        gain, pad = ratio_pad[0][0], ratio_pad[1]
    else:
        gain = min(img1_shape[0] / img0_shape[1], img1_shape[1] / img0_shape[2])
        pad = (round((img1_shape[1] - img0_shape[2] * gain) / 2 - 0.1), round((img1_shape[0] - img0_shape[1] * gain) / 2 - 0.1))
    if padding:
        boxes[..., 0] -= pad[0]
        boxes[..., 1] -= pad[1]
        if not xywh:
            boxes[..., 2] -= pad[0]
            boxes[..., 3] -= pad[1]
    boxes[..., :4] /= gain
    return clip_boxes(boxes, img0_shape)


def convert_res(data):
    """:255-258: (n, 6) detections -> boxes (n, 4), [conf, class] (n, 2)"""
    return data[:, :4], data[:, -2:]


def coord(prediction):
    """:146-190: prediction (1,84,8400) fp32 CUDA -> [tensor (n,6)] (rows x1 y1 x2 y2 conf class, n <= 300).
    Like the reference only the first image is processed (`return output` sits inside its loop, :190); when no anchor
    passes the confidence threshold the reference falls off the loop and returns None."""
    _fq._need_cuda(prediction, 'coord')
    e = _engines.get(prediction.device.index or 0)
    if e is None:
        raise _eng.AyqError('coord: create a stage_8_torch.Yolov8 (load_state_dict + .to(cuda)) on this device first')
    dets, counts = e.coord_float(prediction[:1])
    k = int(counts[0].item())
    if k == 0:
        return None
    return [dets[0, :k].clone()]


class Yolov8(_fq.Yolov8):
    """stage_8_torch.Yolov8 (:284-961): same state_dict layout and integer convolutions as the full-quant model, float head."""
    _HEAD = 'float'

    def _cfg(self):
        return _state['configured'], all_scales, max_a_dict, K, SIGMOID_RANGE

    def _register(self, index):
        _engines[index] = self._engine
