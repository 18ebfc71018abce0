"""Experiment harness: event-timed passes of one batch (graph path) and, optionally, the per-op table.

    [AYQ_LIB=path/to/variant.so] [AYQ_*=...] python tools/exp.py --batch 256 --steps 10 [--ops] [--dual] [--tag name]

--dual: two engines with batch/2 images each on two streams (tail overlap experiment).
Prints one line: tag, ms per batch, images/s; with --ops the per-op times sorted by plan order.
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
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--steps', type=int, default=10)
ap.add_argument('--k', type=int, default=8)
ap.add_argument('--ops', action='store_true')
ap.add_argument('--dual', action='store_true')
ap.add_argument('--tag', default='exp')
ap.add_argument('--json', default=None)
args = ap.parse_args()
K, sd, sc, ma = loaders.load_workload_npz(os.path.join(REPO, 'tests', 'golden', f'workload_k{args.k}.npz'))
p = plan.compile_plan(sd, sc, ma, K)
B = args.batch
x = (torch.from_numpy(bench.synth_batch_u8(B)).float() / 255.0).cuda()
dets = torch.empty((B, 300, 6), dtype=torch.float32, device='cuda')
counts = torch.empty((B,), dtype=torch.int32, device='cuda')
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if args.dual:
    h = B // 2
    es = [engine.Engine(p, 0, h), engine.Engine(p, 0, h)]
    ss = [torch.cuda.Stream(), torch.cuda.Stream()]
    cur = torch.cuda.current_stream()

    def step():
        e0 = torch.cuda.Event(); e0.record(cur)
        for i in range(2):
            ss[i].wait_event(e0)
            with torch.cuda.stream(ss[i]):
                es[i].forward_into(x[i * h:(i + 1) * h], dets[i * h:(i + 1) * h], counts[i * h:(i + 1) * h])
            e1 = torch.cuda.Event(); e1.record(ss[i]); cur.wait_event(e1)
else:
    e = engine.Engine(p, 0, B)

    def step():
        e.forward_into(x, dets, counts)
for _ in range(3):
    step()
torch.cuda.synchronize()
ev0.record()
for _ in range(args.steps):
    step()
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / args.steps
print(f'{args.tag}: {ms:.3f} ms per {B} images = {B / ms * 1e3:.0f} images/s, detections {int(counts.sum())}', flush=True)
out = {'tag': args.tag, 'ms': ms, 'ips': B / ms * 1e3}
if args.ops and not args.dual:
    e.set_profiling(True)
    for _ in range(3):
        e.forward_into(x, dets, counts)
    torch.cuda.synchronize()
    op_ms, op_calls = e.op_times()
    names = ['absmax'] + list(p.op_names)
    rows = []
    for i in range(len(op_ms)):
        if op_calls[i]:
            rows.append((names[i] or f'op{i - 1}', float(op_ms[i] / op_calls[i]) * 1e3))
    print(' '.join(f'{n}={t:.1f}' for n, t in rows))
    print(f'sum of ops {sum(t for _, t in rows):.1f} us')
    out['ops'] = rows
if args.json:
    json.dump(out, open(args.json, 'w'))
