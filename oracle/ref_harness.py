"""Reference harness -- TEST INFRASTRUCTURE, runs only in the build container.

Executes the UNMODIFIED reference scripts under /root/reference (read-only) on a
synthetic workload and captures (a) the artefacts the hot path consumes
(QUANT_WEIGHTS_K.pickle, bias_scales/, max_a.txt -> tests/golden/workload_k{K}.npz)
and (b) golden outputs of stage_8_torch_full_quant.Yolov8.forward for a set of
seeded synthetic images (-> tests/golden/golden_k{K}.npz).

Nothing here is imported by the product.  /root/reference does not exist on the
GPU box, so this file is never run there; its outputs are committed fixtures.

Recipe follows SURVEY.md Appendix C:
  * package alias  yolov8n_quantisation -> /root/reference   (overlay dir of symlinks,
    with a generated stage_0.py when K != 8, Appendix C.7)
  * sys.modules stubs for deeplake / map_boxes / ultralytics / matplotlib / seaborn
  * cwd with utils/cats_2_640.jpg and {K}_nano/
  * os.utime re-stamp of weights_pickle/* before stage_7 (Python 3.12 gzip handles)
  * Tensor.argsort(stable=True) forced while the reference NMS runs (hard part 3)

Usage:  python oracle/ref_harness.py --k 8 [--work /tmp/ayq_work] [--skip-pipeline]
"""
import argparse
import hashlib
import io
import os
import runpy
import sys
import time
import types

import numpy as np
import torch

REF = '/root/reference'
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import synth  # noqa: E402  (seeded image / weight generators shared with tests)

# execution order of conv_quant() calls in stage_6_full_quant.py (= SURVEY Appendix B order)
LAYER_ORDER = [
    'Conv_P1', 'Conv_P2', 'C2F_2_conv_0', 'C2F_2_bottle_0', 'C2F_2_bottle_1', 'C2F_2_conv_1',
    'Conv_P3', 'C2F_4_conv_0', 'C2F_4_bottle_0', 'C2F_4_bottle_1', 'C2F_4_bottle_2', 'C2F_4_bottle_3', 'C2F_4_conv_1',
    'Conv_P4', 'C2F_6_conv_0', 'C2F_6_bottle_0', 'C2F_6_bottle_1', 'C2F_6_bottle_2', 'C2F_6_bottle_3', 'C2F_6_conv_1',
    'Conv_P5', 'C2F_8_conv_0', 'C2F_8_bottle_0', 'C2F_8_bottle_1', 'C2F_8_conv_1',
    'SPPF_conv_0', 'SPPF_conv_1',
    'C2F_12_conv_0', 'C2F_12_bottle_0', 'C2F_12_bottle_1', 'C2F_12_conv_1',
    'C2F_15_conv_0', 'C2F_15_bottle_0', 'C2F_15_bottle_1', 'C2F_15_conv_1',
    'Conv_16', 'C2F_18_conv_0', 'C2F_18_bottle_0', 'C2F_18_bottle_1', 'C2F_18_conv_1',
    'Conv_19', 'C2F_21_conv_0', 'C2F_21_bottle_0', 'C2F_21_bottle_1', 'C2F_21_conv_1',
    'x_result_5_up_0', 'x_result_5_up_1', 'x_result_5_up_2',
    'x_result_5_down_0', 'x_result_5_down_1', 'x_result_5_down_2',
    'x_result_6_up_0', 'x_result_6_up_1', 'x_result_6_up_2',
    'x_result_6_down_0', 'x_result_6_down_1', 'x_result_6_down_2',
    'x_up_0', 'x_up_1', 'x_up_2', 'x_down_0', 'x_down_1', 'x_down_2',
    'dfl',
]


# --------------------------------------------------------------------------- environment
def make_overlay(work, k):
    """Package alias; for K != 8 an overlay whose stage_0.py is generated (Appendix C.7)."""
    pkg_root = os.path.join(work, 'pkg')
    os.makedirs(pkg_root, exist_ok=True)
    alias = os.path.join(pkg_root, 'yolov8n_quantisation')
    if os.path.islink(alias) or os.path.exists(alias):
        return pkg_root
    if k == 8:
        os.symlink(REF, alias)
        return pkg_root
    os.makedirs(os.path.join(alias, 'quantisation'))
    qsrc = os.path.join(REF, 'quantisation')
    for name in os.listdir(qsrc):
        if name == 'stage_0.py':
            continue
        os.symlink(os.path.join(qsrc, name), os.path.join(alias, 'quantisation', name))
    with open(os.path.join(qsrc, 'stage_0.py')) as f:
        src = f.read()
    assert 'K = 8' in src
    with open(os.path.join(alias, 'quantisation', 'stage_0.py'), 'w') as f:
        f.write(src.replace('K = 8', f'K = {k}', 1))
    return pkg_root


def install_stubs(calib_images):
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches', 'seaborn'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['matplotlib'].patches = sys.modules['matplotlib.patches']

    mb = types.ModuleType('map_boxes')
    mb.mean_average_precision_for_boxes = lambda *a, **k: (0.0, {})
    sys.modules['map_boxes'] = mb

    dl = types.ModuleType('deeplake')

    class _DS:
        def pytorch(self, num_workers=0, batch_size=1, transform=None, shuffle=False):
            tf = transform['images']
            for img_u8 in calib_images:           # HWC uint8 ndarray
                yield {'images': tf(img_u8).unsqueeze(0),
                       'boxes': torch.zeros((1, 1, 4)), 'categories': torch.zeros((1, 1))}

        def __repr__(self):
            return f'<synthetic calibration set, {len(calib_images)} images>'

    dl.load = lambda *a, **k: _DS()
    sys.modules['deeplake'] = dl

    ul = types.ModuleType('ultralytics')

    class YOLO:
        """Stands in for the checkpoint: stage_1 maps weights BY POSITION (stage_1.py:771-779),
        so we hand back stage_1's own freshly-initialised model tensors, re-randomised by
        oracle.synth.synth_float_weights (BN statistics, dfl=arange(16), class-branch tuning)."""

        def __init__(self, path):
            frame = sys._getframe(1)
            m = frame.f_globals['model']
            self._sd = synth.synth_float_weights(m.state_dict(), m)

        def state_dict(self):
            return self._sd

    ul.YOLO = YOLO
    sys.modules['ultralytics'] = ul


def run_stage(pkg_root, name):
    t0 = time.time()
    path = os.path.join(pkg_root, 'yolov8n_quantisation', 'quantisation', name)
    out = io.StringIO()
    real_stdout = sys.stdout
    sys.stdout = out
    try:
        g = runpy.run_path(path, run_name='__main__')
    finally:
        sys.stdout = real_stdout
    print(f'[harness] {name}: {time.time() - t0:.1f}s')
    return g


def run_pipeline(work, pkg_root, k):
    """stage_1 -> 2 -> 4 -> 5 -> 6_full_quant -> 7, unmodified."""
    main_dir = f'{k}_nano'
    run_stage(pkg_root, 'stage_1.py')
    run_stage(pkg_root, 'stage_2.py')
    run_stage(pkg_root, 'stage_4.py')
    run_stage(pkg_root, 'stage_5.py')

    # stage_6_full_quant: silence the Verilog / first-pixel dumps and the sleeps (Appendix C.5);
    # arithmetic untouched.  `from module import *` copies these names at import time, so patch
    # the defining modules first.
    import importlib
    sw = importlib.import_module('yolov8n_quantisation.quantisation.utils.save_weights')
    cp = importlib.import_module('yolov8n_quantisation.quantisation.utils.conv2d_print_fp')
    rt = importlib.import_module('yolov8n_quantisation.quantisation.utils.result_txt')
    noop = lambda *a, **kw: None
    for mod in (sw, cp, rt):
        for nm in dir(mod):
            if nm.startswith('save_txt') or nm in ('conv2d', 'result_txt', 'add_rescale_shift', 'add_silu'):
                setattr(mod, nm, noop)
    real_sleep = time.sleep
    time.sleep = noop
    try:
        run_stage(pkg_root, 'stage_6_full_quant.py')
    finally:
        time.sleep = real_sleep
    import gc
    gc.collect()  # finalise the unclosed gzip handles (Appendix C.4)

    wp = os.path.join(main_dir, 'weights_pickle')
    t = time.time() - 10000
    for i, layer in enumerate(LAYER_ORDER):
        for j, suffix in enumerate(('conv', 'bias')):
            p = os.path.join(wp, f'{layer}_{suffix}.pickle')
            assert os.path.exists(p), p
            os.utime(p, (t + 2 * i + j, t + 2 * i + j))
    run_stage(pkg_root, 'stage_7.py')


# --------------------------------------------------------------------------- capture
def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()[:16]


def capture_goldens(work, pkg_root, k, images_u8, keep_full, dump_dir=None):
    """Import stage_8_torch_full_quant and record every integer tensor of forward()."""
    g = run_stage(pkg_root, 'stage_8_torch_full_quant.py')
    model = g['model']
    fwd_globals = g['silu'].__globals__
    rec = {}
    state = {}

    real_silu = g['silu']
    real_requantize = g['requantize']
    real_coord_quant = g['coord_quant']
    real_conv_forward = torch.nn.Conv2d.forward

    def conv_forward(self, x):
        y = real_conv_forward(self, x)
        state['n_conv'] += 1
        # exactness certificate: fp32 conv == fp64 conv, integer inputs
        if state['certify']:
            y64 = torch.nn.functional.conv2d(x.double(), self.weight.double(),
                                             None if self.bias is None else self.bias.double(),
                                             self.stride, self.padding)
            state['conv_mismatch'] += int((y64 != y.double()).sum())
            state['max_abs_in'] = max(state['max_abs_in'], float(x.abs().max()))
            state['max_abs_acc'] = max(state['max_abs_acc'], float(y.abs().max()))
            state['non_integer_in'] += int((x != x.round()).sum())
        state['conv_out'].append(y)
        return y

    def silu(x, scale_x, a_input):
        r = real_silu(x, scale_x, a_input)
        state['silu_out'].append(r[0].clone())
        return r

    def requantize(arr, old, new, bits, device, bit_size_for_koeff=8):
        r = real_requantize(arr, old, new, bits, device, bit_size_for_koeff)
        state['coeffs'].append((np.asarray(r[1], dtype=np.float64).reshape(-1),
                                np.asarray(r[2], dtype=np.float64).reshape(-1)))
        if not state['in_silu']:
            state['requant_out'].append(r[0].clone())
        return r

    def silu_wrapped(x, scale_x, a_input):
        state['in_silu'] = True
        try:
            return silu(x, scale_x, a_input)
        finally:
            state['in_silu'] = False

    def coord_quant(pred):
        state['dbox_cls'] = pred.clone()
        real_argsort = torch.Tensor.argsort

        def stable_argsort(self, *a, **kw):
            kw['stable'] = True
            return real_argsort(self, *a, **kw)
        torch.Tensor.argsort = stable_argsort
        try:
            return real_coord_quant(pred)
        finally:
            torch.Tensor.argsort = real_argsort

    fwd_globals['silu'] = silu_wrapped
    fwd_globals['requantize'] = requantize
    fwd_globals['coord_quant'] = coord_quant
    torch.nn.Conv2d.forward = conv_forward

    out = {}
    cert = dict(conv_mismatch=0, max_abs_in=0.0, max_abs_acc=0.0, non_integer_in=0)
    try:
        for i, img_u8 in enumerate(images_u8):
            state.update(n_conv=0, conv_out=[], silu_out=[], requant_out=[], in_silu=False, coeffs=[],
                         dbox_cls=None, certify=(i < 2), conv_mismatch=0, max_abs_in=0.0,
                         max_abs_acc=0.0, non_integer_in=0)
            x = synth.to_input_tensor(img_u8)
            t0 = time.time()
            with torch.no_grad():
                boxes, classes = model(x)
            dt = time.time() - t0
            assert state['n_conv'] == 64, state['n_conv']          # 63 convs + dfl
            assert len(state['silu_out']) == 57
            convs = state['conv_out']
            # the six raw head accumulators are conv outputs 47,50,53,56,59,62 (0-based, Appendix B order)
            pre = f'img{i}_'
            # canonical hash form: NCHW int32 bytes
            i32 = lambda t: t.numpy().astype(np.int32)
            out[pre + 'silu_sha'] = np.array([sha(i32(t)) for t in state['silu_out']])
            out[pre + 'conv_sha'] = np.array([sha(i32(t)) for t in convs[:63]])
            out[pre + 'requant_sha'] = np.array([sha(i32(t)) for t in state['requant_out']])
            d = state['dbox_cls'][0].numpy()                        # (84, 8400) fp32, integer valued
            out[pre + 'dbox_sha'] = np.array(sha(d[:4].astype(np.int32)))
            out[pre + 'cls_sha'] = np.array(sha(d[4:].astype(np.int32)))
            out[pre + 'dbox'] = d[:4].astype(np.int32)
            out[pre + 'score_max'] = d[4:].max(0).astype(np.int32)
            out[pre + 'score_arg'] = d[4:].argmax(0).astype(np.int16)
            if boxes is None:
                out[pre + 'boxes'] = np.zeros((0, 4), np.float32)
                out[pre + 'classes'] = np.zeros((0, 2), np.float32)
            else:
                out[pre + 'boxes'] = boxes.numpy().astype(np.float32)
                out[pre + 'classes'] = classes.numpy().astype(np.float32)
            ncand = int((d[4:].max(0) > 8192).sum())
            out[pre + 'ncand'] = np.array(ncand)
            if dump_dir and i in keep_full:
                os.makedirs(dump_dir, exist_ok=True)
                np.savez(os.path.join(dump_dir, f'img{i}.npz'),
                         **{f'silu{j}': i32(t) for j, t in enumerate(state['silu_out'])},
                         **{f'conv{j}': i32(t) for j, t in enumerate(convs[:63])},
                         **{f'requant{j}': i32(t) for j, t in enumerate(state['requant_out'])},
                         dbox_cls=d)
            if i == 0:
                # (k, s) of all 135 requantize() calls in call order (static: scales do not depend on the image)
                out['coeff_len'] = np.array([len(c[0]) for c in state['coeffs']])
                out['coeff_k'] = np.concatenate([c[0] for c in state['coeffs']]).astype(np.int32)
                out['coeff_s'] = np.concatenate([np.broadcast_to(c[1], c[0].shape) for c in state['coeffs']]).astype(np.int32)
            if state['certify']:
                for key in cert:
                    cert[key] = max(cert[key], state[key]) if key.startswith('max') else cert[key] + state[key]
            print(f'[harness] image {i}: {dt:.2f}s ncand={ncand} ndet={len(out[pre + "boxes"])}')
    finally:
        torch.nn.Conv2d.forward = real_conv_forward
    out['n_images'] = np.array(len(images_u8))
    lut = lambda d: np.array([d[key] for key in sorted(d.keys())], dtype=np.float64)
    out['lut_sigmoid'] = lut(g['lookup']).astype(np.int32)
    out['lut_sigmoid16'] = lut(g['lookup_final']).astype(np.int32)
    out['lut_exp'] = lut(g['lookup_exp']).astype(np.int32)
    out['lut_dtype'] = np.array(str(type(next(iter(g['lookup'].values())))))
    for key, v in cert.items():
        out['cert_' + key] = np.array(v)
    print('[harness] exactness certificate:', cert)
    return out, g


def capture_goldens_float(work, pkg_root, k, images_u8):
    """Import the UNMODIFIED stage_8_torch.py (float Detect head + torchvision NMS) and record,
    per image: hashes of the 63 integer conv outputs and the 57 silu() outputs (sigmoid range 7), the float
    prediction tensor dbox_cls (84, 8400) on every 8th anchor, and the returned (boxes, classes)."""
    g = run_stage(pkg_root, 'stage_8_torch.py')
    model = g['model']
    fwd_globals = g['silu'].__globals__
    state = {}
    real_silu = g['silu']
    real_coord = g['coord']
    real_conv_forward = torch.nn.Conv2d.forward

    def conv_forward(self, x):
        y = real_conv_forward(self, x)
        state['conv_out'].append(y)
        return y

    def silu(x, scale_x, a_input):
        r = real_silu(x, scale_x, a_input)
        state['silu_out'].append(r[0].clone())
        return r

    def coord(pred):
        state['dbox_cls'] = pred.clone()         # coord() rewrites the first four rows in place (:160)
        return real_coord(pred)

    fwd_globals['silu'] = silu
    fwd_globals['coord'] = coord
    torch.nn.Conv2d.forward = conv_forward
    out = {}
    try:
        for i, img_u8 in enumerate(images_u8):
            state.update(conv_out=[], silu_out=[], dbox_cls=None)
            x = synth.to_input_tensor(img_u8)
            with torch.no_grad():
                boxes, classes = model(x)
            assert len(state['conv_out']) == 64 and len(state['silu_out']) == 57
            pre = f'img{i}_'
            i32 = lambda t: t.numpy().astype(np.int32)
            for t in state['conv_out'][:63]:
                assert bool((t == t.round()).all())
            out[pre + 'conv_sha'] = np.array([sha(i32(t)) for t in state['conv_out'][:63]])
            out[pre + 'silu_sha'] = np.array([sha(i32(t)) for t in state['silu_out']])
            d = state['dbox_cls'][0].numpy().astype(np.float32)
            out[pre + 'dbox_cls_s8'] = d[:, ::8].copy()
            out[pre + 'conf_max'] = d[4:].max(0)
            out[pre + 'conf_arg'] = d[4:].argmax(0).astype(np.int16)
            if boxes is None:
                out[pre + 'boxes'] = np.zeros((0, 4), np.float32)
                out[pre + 'classes'] = np.zeros((0, 2), np.float32)
            else:
                out[pre + 'boxes'] = boxes.numpy().astype(np.float32)
                out[pre + 'classes'] = classes.numpy().astype(np.float32)
            print(f'[harness] float head image {i}: ndet={len(out[pre + "boxes"])} conf range '
                  f'{float(d[4:].max(0).min()):.3g}..{float(d[4:].max()):.3g}')
    finally:
        torch.nn.Conv2d.forward = real_conv_forward
    out['n_images'] = np.array(len(images_u8))
    import torchvision
    out['versions'] = np.array(f'torch {torch.__version__} torchvision {torchvision.__version__} numpy {np.__version__}')
    return out


def crafted_predictions():
    """Hand-made (84, 8400) prediction tensors for the q_NMS corner cases the forward goldens never reach:
    more than 300 kept rows (`i[:max_det]`, stage_8_torch_full_quant.py:354), more than 1000 candidates with > 300 survivors,
    exactly 300 / 301 survivors, heavy score ties.  Integer valued like the reference's dbox_cls: xywh in 412.1635-per-pixel
    units, class scores 0..32767.  Stored compactly as (dbox int32 (4,8400), score_max int32, score_arg int16)."""
    A = 8400
    rng = np.random.default_rng(2024)
    cases = {}

    def grid_case(n_boxes, cols, pitch_px, size_px, score_fn, cls_fn, jitter=0):
        d = np.zeros((4, A), np.int32)
        sm = np.zeros(A, np.int32)
        sa = np.zeros(A, np.int16)
        idx = rng.permutation(A)[:n_boxes]
        idx.sort()
        for q, a in enumerate(idx):
            cx = (8 + pitch_px * (q % cols)) * 412 + (int(rng.integers(-jitter, jitter + 1)) if jitter else 0)
            cy = (8 + pitch_px * (q // cols)) * 412 + (int(rng.integers(-jitter, jitter + 1)) if jitter else 0)
            d[:, a] = (cx, cy, size_px * 412, size_px * 412)
            sm[a] = score_fn(q)
            sa[a] = cls_fn(q)
        return d, sm, sa

    # A: 420 separated boxes of ONE class, distinct scores -> 420 survivors, the first 300 are returned
    cases['keep420_distinct'] = grid_case(420, 40, 15, 10, lambda q: 32000 - 7 * q, lambda q: 3)
    # B: 1500 candidates of one class on a 50-column grid, scores from a set of 25 values (ties), boxes separated -> top-1000 cut, then 300
    cases['cand1500_ties'] = grid_case(1500, 50, 12, 8, lambda q: 9000 + 100 * (q * 7 % 25), lambda q: 11)
    # C / D: exactly 300 and 301 separated boxes (the boundary of the cap), mixed classes (class offset j * 7680 shifts the boxes)
    cases['keep300_exact'] = grid_case(300, 30, 20, 9, lambda q: 20000 + 3 * q, lambda q: q % 80)
    cases['keep301'] = grid_case(301, 30, 20, 9, lambda q: 20000 + 3 * (q % 97), lambda q: (q * 5) % 80)
    # E: dense overlapping field, 3000 candidates, jittered boxes of 60 px on a 6 px pitch, 40 tied score levels
    cases['dense3000'] = grid_case(3000, 100, 6, 60, lambda q: 8200 + 13 * (q * 11 % 40), lambda q: q % 3, jitter=300)
    # F: 8400 candidates, every anchor a small separated box (100 x 84 grid on a 7 px pitch): > 300 survivors out of the top 1000
    cases['all8400'] = grid_case(8400, 100, 7, 5, lambda q: 8193 + (q * 2654435761 % 20000), lambda q: (q // 7) % 80)
    return cases


def capture_nms_extra(work, pkg_root, k, n_images):
    """(1) crafted predictions through the UNMODIFIED coord_quant() + the forward() tail (scale_boxes / convert_res), with
    argsort(stable=True) forced (the pinned order) AND exactly as shipped (torch's default argsort) -> golden_nms_k{K}.npz;
    (2) the forward goldens' images once more with the shipped argsort, to report how often the reference as users run it
    differs from the pinned order (SURVEY hard part 3)."""
    g = run_stage(pkg_root, 'stage_8_torch_full_quant.py')
    model = g['model']
    fwd_globals = g['silu'].__globals__
    coord_quant, scale_boxes, convert_res = g['coord_quant'], g['scale_boxes'], g['convert_res']
    real_argsort = torch.Tensor.argsort

    def tail(pred, stable):
        """forward() :1265-1275 on a prediction tensor"""
        def stable_argsort(self, *a, **kw):
            kw['stable'] = True
            return real_argsort(self, *a, **kw)
        if stable:
            torch.Tensor.argsort = stable_argsort
        try:
            res = coord_quant(pred.clone())
        finally:
            torch.Tensor.argsort = real_argsort
        if res is None:
            return np.zeros((0, 4), np.float32), np.zeros((0, 2), np.float32)
        for i, p_ in enumerate(res):                                  # :1267-1270 with orig_img (1, 3, 640, 640)
            p_[:, :4] = scale_boxes((640, 640), p_[:, :4], (3, 640, 640))
        b, c = convert_res(p_)
        return b.numpy().astype(np.float32), c.numpy().astype(np.float32)

    out = {}
    names = []
    for name, (d, sm, sa) in crafted_predictions().items():
        pred = torch.zeros((1, 84, 8400), dtype=torch.float32)
        pred[0, :4] = torch.from_numpy(d.astype(np.float32))
        pred[0, 4 + torch.from_numpy(sa.astype(np.int64)), torch.arange(8400)] = torch.from_numpy(sm.astype(np.float32))
        bs, cs = tail(pred, True)
        bu, cu = tail(pred, False)
        out[f'{name}/dbox'] = d
        out[f'{name}/score_max'] = sm
        out[f'{name}/score_arg'] = sa
        out[f'{name}/boxes'] = bs
        out[f'{name}/classes'] = cs
        out[f'{name}/boxes_unpatched'] = bu
        out[f'{name}/classes_unpatched'] = cu
        names.append(name)
        same = bs.shape == bu.shape and np.array_equal(bs, bu) and np.array_equal(cs, cu)
        print(f'[harness] crafted {name}: ncand={(sm > 8192).sum()} kept(stable)={len(bs)} kept(unpatched)={len(bu)} identical={same}')
    out['cases'] = np.array(names)
    # forward goldens with the shipped argsort
    state = {}
    real_coord_quant = g['coord_quant']

    def cq(pred):
        state['pred'] = pred.clone()
        return real_coord_quant(pred)
    fwd_globals['coord_quant'] = cq
    for i in range(n_images):
        x = synth.to_input_tensor(synth.synth_image_u8(i))
        with torch.no_grad():
            boxes, classes = model(x)
        out[f'img{i}_boxes_unpatched'] = np.zeros((0, 4), np.float32) if boxes is None else boxes.numpy().astype(np.float32)
        out[f'img{i}_classes_unpatched'] = np.zeros((0, 2), np.float32) if classes is None else classes.numpy().astype(np.float32)
        print(f'[harness] image {i} with the shipped argsort: ndet={len(out[f"img{i}_boxes_unpatched"])}')
    out['n_images'] = np.array(n_images)
    out['versions'] = np.array(f'torch {torch.__version__} numpy {np.__version__}')
    return out


def capture_weight_quant(work, pkg_root, k, layers):
    """Run the UNMODIFIED stage_6_full_quant.py (text dumps silenced exactly as in run_pipeline) under a profile hook and
    record, for the chosen layers, the arguments and results of conv_quant() (stage_6_full_quant.py:89-126):
    float weights / bias, scale_input, start -> integer weights, integer bias, per-channel scale_res."""
    import importlib
    sw = importlib.import_module('yolov8n_quantisation.quantisation.utils.save_weights')
    cp = importlib.import_module('yolov8n_quantisation.quantisation.utils.conv2d_print_fp')
    rt = importlib.import_module('yolov8n_quantisation.quantisation.utils.result_txt')
    noop = lambda *a, **kw: None
    for mod in (sw, cp, rt):
        for nm in dir(mod):
            if nm.startswith('save_txt') or nm in ('conv2d', 'result_txt', 'add_rescale_shift', 'add_silu'):
                setattr(mod, nm, noop)
    out = {}
    pending = {}

    def prof(frame, event, arg):
        if frame.f_code.co_name != 'conv_quant':
            return
        if event == 'call':
            loc = frame.f_locals
            name = loc['layer_name']
            if name in layers:
                pending[id(frame)] = (name, np.array(loc['conv'], copy=True), np.array(loc['bias_conv'], copy=True),
                                      np.float64(loc['scale_input']) if np.ndim(loc['scale_input']) == 0 else np.array(loc['scale_input'], np.float64),
                                      bool(loc['start']))
        elif event == 'return' and id(frame) in pending:
            name, w, b, si, start = pending.pop(id(frame))
            loc = frame.f_locals
            out[f'{name}/w'] = w
            out[f'{name}/w_dtype'] = np.array(str(w.dtype))
            out[f'{name}/b'] = b
            out[f'{name}/b_dtype'] = np.array(str(b.dtype))
            out[f'{name}/scale_input'] = np.asarray(si, np.float64)
            out[f'{name}/scale_input_type'] = np.array(str(type(loc['scale_input'])))
            out[f'{name}/start'] = np.array(start)
            out[f'{name}/qw'] = np.asarray(loc['conv']).astype(np.int8)
            out[f'{name}/qb'] = np.asarray(loc['bias']).astype(np.int64)
            out[f'{name}/scale_res'] = np.asarray(loc['scale_res'], np.float64)
            out[f'{name}/scale_res_dtype'] = np.array(str(np.asarray(loc['scale_res']).dtype))

    real_sleep = time.sleep
    time.sleep = noop
    sys.setprofile(prof)
    try:
        run_stage(pkg_root, 'stage_6_full_quant.py')
    finally:
        sys.setprofile(None)
        time.sleep = real_sleep
    out['layers'] = np.array(sorted(set(kk.split('/')[0] for kk in out)))
    out['versions'] = np.array(f'numpy {np.__version__}')
    return out


def export_workload(work, k, g):
    """Pack what the hot path reads from disk into one small npz."""
    main_dir = f'{k}_nano'
    sd = torch.load(os.path.join(main_dir, 'results', f'QUANT_WEIGHTS_{k}.pickle'))
    out = {'K': np.array(k)}
    names = []
    for name, t in sd.items():
        a = t.numpy()
        assert (a == np.round(a)).all(), name
        names.append(name)
        if name.endswith('weight'):
            assert np.abs(a).max() <= 127
            out['sd/' + name] = a.astype(np.int8)
        else:
            out['sd/' + name] = a.astype(np.int64)
    out['sd_keys'] = np.array(names)
    scales = g['all_scales']
    for name, t in scales.items():
        out['scale/' + name] = t.numpy().astype(np.float32).reshape(-1)
    out['scale_keys'] = np.array(sorted(scales.keys()))
    with open(os.path.join(main_dir, 'results', 'max_a.txt')) as f:
        out['max_a_txt'] = np.array(f.read())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--k', type=int, default=8)
    ap.add_argument('--work', default='/tmp/ayq_work')
    ap.add_argument('--skip-pipeline', action='store_true')
    ap.add_argument('--n-golden', type=int, default=12)
    ap.add_argument('--dump', default=None, help='scratch dir for full per-layer tensors (not committed)')
    ap.add_argument('--out', default=os.path.join(REPO, 'tests', 'golden'))
    ap.add_argument('--weight-quant', default='', metavar='LAYERS',
                    help='only record conv_quant() goldens of stage_6_full_quant.py for these comma-separated layers')
    ap.add_argument('--export-main-dir', action='store_true',
                    help='only pack the files the hot path reads, AS THE REFERENCE WROTE THEM (stage_7.py:780 torch.save state_dict, '
                         'utils/save_weights.py:24-30 gzip-pickle bias_scales/, stage_5 max_a.txt), into main_dir_k{K}.tar.xz')
    ap.add_argument('--time', type=int, default=0, metavar='N',
                    help='only time the UNMODIFIED stage_8_torch_full_quant forward on N synthetic images (batch 1 per call, all host '
                         'threads) next to the numpy port on the same machine -> profiles/reference_cpu_r2.json')
    ap.add_argument('--nms-extra', action='store_true',
                    help='only record the crafted q_NMS corner cases and the shipped-argsort detections -> golden_nms_k{K}.npz')
    ap.add_argument('--float-head', type=int, default=0, metavar='N',
                    help='only record stage_8_torch.py (float head) goldens for N images -> golden_float_k{K}.npz')
    args = ap.parse_args()
    k = args.k
    work = os.path.join(args.work, f'k{k}')
    os.makedirs(os.path.join(work, 'utils'), exist_ok=True)
    os.makedirs(os.path.join(work, 'input_data'), exist_ok=True)
    cat = os.path.join(work, 'utils', 'cats_2_640.jpg')
    if not os.path.exists(cat):
        os.symlink(os.path.join(REF, 'quantisation', 'utils', 'cats_2_640.jpg'), cat)
    pkg_root = make_overlay(work, k)
    sys.path.insert(0, pkg_root)
    os.chdir(work)
    torch.set_num_threads(os.cpu_count())

    calib = [synth.synth_image_u8(1000 + i).transpose(1, 2, 0).copy() for i in range(synth.N_CALIB)]
    install_stubs(calib)
    if not args.skip_pipeline:
        run_pipeline(work, pkg_root, k)
    if args.weight_quant:
        gold = capture_weight_quant(work, pkg_root, k, set(args.weight_quant.split(',')))
        np.savez_compressed(os.path.join(args.out, f'golden_wquant_k{k}.npz'), **gold)
        print('[harness] wrote weight-quantiser goldens:', list(gold['layers']))
        return
    if args.time:
        import json
        g = run_stage(pkg_root, 'stage_8_torch_full_quant.py')
        model = g['model']
        xs = [synth.to_input_tensor(synth.synth_image_u8(100 + i)) for i in range(args.time + 1)]
        with torch.no_grad():
            model(xs[0])                                            # warm-up
            t0 = time.time()
            for x in xs[1:]:
                model(x)
            dt_ref = time.time() - t0
        from oracle import yolo_int as Y
        o = Y.OracleYolov8(Y.Workload(os.path.join(REPO, 'tests', 'golden', f'workload_k{k}.npz')))
        xa = [synth.to_input_array([synth.synth_image_u8(100 + i)]) for i in range(args.time + 1)]
        o.forward(xa[0])
        t0 = time.time()
        for x in xa[1:]:
            o.forward(x)
        dt_port = time.time() - t0
        out = {'cores': os.cpu_count(), 'images': args.time,
               'reference_unmodified_images_per_s': args.time / dt_ref, 'port_single_process_images_per_s': args.time / dt_port,
               'what': 'unmodified stage_8_torch_full_quant.Yolov8.forward (torch CPU ops, torch.set_num_threads(all cores), batch 1 per call) vs '
                       'oracle/yolo_int.py (numpy, one process) on the same synthetic images in the build container; bench.py times the port '
                       'with one worker process per core on the GPU box, where /root/reference does not exist',
               'versions': f'torch {torch.__version__} numpy {np.__version__}'}
        os.makedirs(os.path.join(REPO, 'profiles'), exist_ok=True)
        json.dump(out, open(os.path.join(REPO, 'profiles', 'reference_cpu_r2.json'), 'w'), indent=1)
        print('[harness]', out)
        return
    if args.export_main_dir:
        import tarfile
        main_dir = f'{k}_nano'
        dst = os.path.join(args.out, f'main_dir_k{k}.tar.xz')
        with tarfile.open(dst, 'w:xz') as tf:
            tf.add(os.path.join(main_dir, 'results', f'QUANT_WEIGHTS_{k}.pickle'))
            tf.add(os.path.join(main_dir, 'results', 'max_a.txt'))
            tf.add(os.path.join(main_dir, 'bias_scales'))
        print('[harness] wrote', dst, os.path.getsize(dst), 'bytes')
        return
    if args.nms_extra:
        gold = capture_nms_extra(work, pkg_root, k, args.n_golden)
        np.savez_compressed(os.path.join(args.out, f'golden_nms_k{k}.npz'), **gold)
        print('[harness] wrote q_NMS corner-case goldens to', args.out)
        return
    if args.float_head:
        images = [synth.synth_image_u8(s) for s in range(args.float_head)]
        gold = capture_goldens_float(work, pkg_root, k, images)
        np.savez_compressed(os.path.join(args.out, f'golden_float_k{k}.npz'), **gold)
        print('[harness] wrote float-head goldens to', args.out)
        return
    images = [synth.synth_image_u8(s) for s in range(args.n_golden)]
    gold, g = capture_goldens(work, pkg_root, k, images, keep_full=(0, 1, 2), dump_dir=args.dump)
    wl = export_workload(work, k, g)
    os.makedirs(args.out, exist_ok=True)
    np.savez_compressed(os.path.join(args.out, f'workload_k{k}.npz'), **wl)
    np.savez_compressed(os.path.join(args.out, f'golden_k{k}.npz'), **gold)
    print('[harness] wrote', args.out)


if __name__ == '__main__':
    main()
