"""BASELINE configs[3]: global-batch sweep through the product-level data-parallel entry (one process, G GPUs): N uint8 host images ->
DataParallelYolo.forward_host -> N results in image order.  End to end (H2D + kernels + D2H inside the timed call).
    python tools/dp_sweep.py [out.json]"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from alpha_yolo_quant_b200 import dataparallel as dp, loaders, plan  # noqa: E402
import bench  # noqa: E402

K, sd, sc, ma = loaders.load_workload_npz(os.path.join(REPO, 'tests', 'golden', 'workload_k8.npz'))
p = plan.compile_plan(sd, sc, ma, K)
n_dev = torch.cuda.device_count()
base = torch.from_numpy(bench.synth_batch_u8(64))
out = {}
for g in [1, 2, 4, 8]:
    if g > n_dev:
        break
    dpy = dp.DataParallelYolo(p, devices=list(range(g)), max_batch=512)
    for n in [1, 8, 64, 256, 1024, 4096]:
        img = base.repeat((n + 63) // 64, 1, 1, 1)[:n].contiguous().pin_memory()
        dets = torch.empty((n, 300, 6)).pin_memory()
        counts = torch.empty((n,), dtype=torch.int32).pin_memory()
        dpy.forward_host(img, dets, counts)
        reps = 3 if n >= 1024 else 10
        t0 = time.perf_counter()
        for _ in range(reps):
            dpy.forward_host(img, dets, counts)
        dt = (time.perf_counter() - t0) / reps
        out[f'{g}gpu_batch{n}'] = n / dt
        print(f'{g} GPU(s), global batch {n:5d}: {dt * 1e3:8.2f} ms = {n / dt:9.0f} images/s end to end (uint8 host images)', flush=True)
        del img, dets, counts
    dpy.close()
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], 'w'), indent=1)
