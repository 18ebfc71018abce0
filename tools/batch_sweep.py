"""BASELINE configs[3]: device-resident throughput of the full path (forward + Detect head + q_NMS) for batch sizes 1 .. 4096
on one GPU (batches above --max-batch run as consecutive passes).  Writes one JSON object; copy it to profiles/.

    python tools/batch_sweep.py --out gpurun_out/batch_sweep_r1.json
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from alpha_yolo_quant_b200 import engine, loaders, plan  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--max-batch', type=int, default=512)
ap.add_argument('--out', default=None)
ap.add_argument('--sizes', default='1,2,4,8,16,32,64,128,256,512,1024,2048,4096')
args = ap.parse_args()
K, sd, sc, ma = loaders.load_workload_npz(os.path.join(REPO, 'tests', 'golden', 'workload_k8.npz'))
p = plan.compile_plan(sd, sc, ma, K)
e = engine.Engine(p, 0, args.max_batch)
e.set_conv_impl('tma')
base = torch.from_numpy(bench.synth_batch_u8(256)).float().div_(255.0).cuda()       # 256 distinct images, tiled for larger batches
rows = []
for B in [int(s) for s in args.sizes.split(',')]:
    x = base[:B] if B <= 256 else base.repeat((B + 255) // 256, 1, 1, 1)[:B].contiguous()
    dets = torch.empty((B, 300, 6), device='cuda')
    counts = torch.empty((B,), dtype=torch.int32, device='cuda')
    iters = max(3, min(50, 4096 // B))
    for _ in range(3):
        e.forward_into(x, dets, counts)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(iters):
        e.forward_into(x, dets, counts)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / iters
    rows.append({'batch': B, 'ms_per_batch': ms, 'images_per_s': 1000.0 * B / ms, 'detections': int(counts.sum())})
    print(rows[-1], flush=True)
    del x, dets, counts
out = {'metric': bench.METRIC, 'config': 'BASELINE configs[3]: batch sweep, 1 GPU, fp32 images resident in HBM, passes of <= %d images' % args.max_batch,
       'rows': rows}
if args.out:
    json.dump(out, open(args.out, 'w'), indent=1)
