"""Per-tensor diff of the CUDA path against the numpy oracle (diagnostic; run on the GPU box).
    python tools/debug_forward.py [--k 8] [--impl dp4a|tcgen05] [--n 2]"""
import argparse
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from alpha_yolo_quant_b200 import engine, loaders, plan  # noqa: E402
from oracle import synth, yolo_int as Y  # noqa: E402
from tests.test_gpu_parity import REQUANT_ORDER  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--k', type=int, default=8)
    ap.add_argument('--impl', default='tma')
    ap.add_argument('--n', type=int, default=2)
    a = ap.parse_args()
    wpath = os.path.join(REPO, 'tests', 'golden', f'workload_k{a.k}.npz')
    K, sd, sc, ma = loaders.load_workload_npz(wpath)
    p = plan.compile_plan(sd, sc, ma, K, taps=True)
    e = engine.Engine(p, 0, 8, lib_path=None if a.impl == 'tma' else engine.TEST_LIB_PATH)
    e.set_conv_impl(a.impl)
    x = synth.to_input_array([synth.synth_image_u8(s) for s in range(a.n)])
    o = Y.OracleYolov8(Y.Workload(wpath))
    ref = o.forward(x, trace=True)
    tr = o.trace
    dets, counts, dbc = e.forward(torch.from_numpy(x).cuda(), want_dbox_cls=True)
    torch.cuda.synchronize()
    nbad = 0
    silu_layers = [nm for nm, _ in plan.LAYERS if 'silu_buf' in p.info['layers'][nm]]
    si = 0
    for t, (nm, _) in enumerate(plan.LAYERS):
        got = e.export_acc_tap(t, a.n).cpu().numpy()
        d = int((got != tr['conv'][t]).sum())
        msg = f'{nm:20s} acc mism {d:8d}/{got.size}'
        if nm in silu_layers:
            gs = e.export_buffer(p.info['layers'][nm]['silu_buf'], a.n).cpu().numpy()
            ds = int((gs != tr['silu'][si]).sum())
            si += 1
            msg += f'   silu mism {ds:8d}'
            d += ds
        if d:
            nbad += 1
        print(msg, '' if not d else '  <<<<')
    for t, (nm, j) in enumerate(REQUANT_ORDER):
        got = e.export_buffer(p.info['layers'][nm]['requant_bufs'][j][0], a.n).cpu().numpy()
        d = int((got != tr['requant'][t]).sum())
        if d:
            nbad += 1
        print(f'requant {t:2d} {nm:20s} mism {d}', '' if not d else '  <<<<')
    db = int((dbc[:, :4].cpu().numpy() != o.last['dbox']).sum())
    dc = int((dbc[:, 4:].cpu().numpy() != o.last['score']).sum())
    print('dbox mism', db, 'score mism', dc)
    for i, (b, c) in enumerate(ref):
        k = int(counts[i])
        kb = 0 if b is None else b.shape[0]
        ok = k == kb and (k == 0 or (np.array_equal(dets[i, :k, :4].cpu().numpy(), b) and np.array_equal(dets[i, :k, 4:6].cpu().numpy(), c)))
        print(f'img {i}: count {k} (oracle {kb})', 'OK' if ok else 'MISMATCH')
        nbad += 0 if ok else 1
    print('TOTAL BAD', nbad + (db > 0) + (dc > 0))


if __name__ == '__main__':
    main()
