"""One or more engine passes over a synthetic batch (profiling target: `ncu ... python tools/one_pass.py --batch 64`)."""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from alpha_yolo_quant_b200 import engine, loaders, plan  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=64)
ap.add_argument('--passes', type=int, default=1)
ap.add_argument('--conv', default='tma')
args = ap.parse_args()
K, sd, sc, ma = loaders.load_workload_npz(os.path.join(REPO, 'tests', 'golden', 'workload_k8.npz'))
p = plan.compile_plan(sd, sc, ma, K)
e = engine.Engine(p, 0, args.batch)
e.set_conv_impl(args.conv)
x = (torch.from_numpy(bench.synth_batch_u8(args.batch)).float() / 255.0).cuda()
for _ in range(args.passes):
    dets, counts = e.forward(x)
torch.cuda.synchronize()
print('ok', int(counts.sum()))
