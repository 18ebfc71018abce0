"""Raw concurrent host->device copy rate of the box: one process, G GPUs, one pinned 315 MB uint8 buffer (256 images) per GPU,
cudaMemcpyAsync on every GPU at once, repeated; aggregate GB/s for G = 1, 2, 4, 8 (as many as visible).  The end-to-end numbers of
bench.py are PCIe / host-memory bound; this is the ceiling they are compared with.   python tools/h2d_probe.py [out.json]"""
import json
import sys
import time

import torch

n_dev = torch.cuda.device_count()
NBYTES = 256 * 3 * 640 * 640
out = {}
for g in [1, 2, 4, 8]:
    if g > n_dev:
        break
    host = [torch.empty(NBYTES, dtype=torch.uint8).pin_memory() for _ in range(g)]
    dev = [torch.empty(NBYTES, dtype=torch.uint8, device=f'cuda:{i}') for i in range(g)]
    streams = [torch.cuda.Stream(device=i) for i in range(g)]
    for rep in range(2):
        for i in range(g):
            with torch.cuda.stream(streams[i]):
                dev[i].copy_(host[i], non_blocking=True)
        for i in range(g):
            streams[i].synchronize()
    reps = 20
    t0 = time.perf_counter()
    for rep in range(reps):
        for i in range(g):
            with torch.cuda.stream(streams[i]):
                dev[i].copy_(host[i], non_blocking=True)
    for i in range(g):
        streams[i].synchronize()
    dt = time.perf_counter() - t0
    out[str(g)] = g * reps * NBYTES / dt / 1e9
    print(f'{g} GPU(s): {out[str(g)]:.1f} GB/s aggregate host->device ({out[str(g)] / g:.1f} per GPU)', flush=True)
    del host, dev
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], 'w'), indent=1)
