"""Multi-GPU parity (run on a box with >= 2 GPUs: gpurun --gpus 2 -- pytest -m gpu tests/test_gpu_multi.py; skipped on one GPU).
The product-level data-parallel entry shards N host images over the visible GPUs; every image's detections must equal what a
single GPU returns for it (which the other GPU tests pin to the reference goldens / the oracle)."""
import os

import numpy as np
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu


def _plan(golden_dir):
    from alpha_yolo_quant_b200 import loaders, plan
    K, sd, sc, ma = loaders.load_workload_npz(os.path.join(golden_dir, 'workload_k8.npz'))
    return plan.compile_plan(sd, sc, ma, K)


def test_sharded_batch_equals_single_gpu_and_goldens(golden_dir):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    from alpha_yolo_quant_b200 import dataparallel as dp, engine
    g = np.load(os.path.join(golden_dir, 'golden_k8.npz'))
    p = _plan(golden_dir)
    G = torch.cuda.device_count()
    seeds = list(range(12)) + [101, 102, 103, 104, 105, 200, 201]            # 19 images: ragged shards on 2, 4 and 8 GPUs
    u8 = torch.from_numpy(np.stack([synth.synth_image_u8(s) for s in seeds])).pin_memory()
    one = engine.Engine(p, 0, 64)
    d1, c1 = one.forward_host(u8)
    one.close()
    dpy = dp.DataParallelYolo(p, devices=list(range(G)), max_batch=64)
    for rep in range(2):                                                     # second call reuses the engines' pipelines
        dN, cN = dpy.forward_host(u8)
        assert torch.equal(cN, c1)
        for i in range(len(seeds)):
            k = int(c1[i])
            assert torch.equal(dN[i, :k], d1[i, :k]), (rep, i)
    f32 = (u8.float() / 255.0).pin_memory()                                  # the reference's own input format
    dF, cF = dpy.forward_host(f32)
    assert torch.equal(cF, c1)
    for i in range(12):                                                      # and the recorded reference results
        k = int(cN[i])
        assert k == g[f'img{i}_boxes'].shape[0]
        assert np.array_equal(dN[i, :k, :4].numpy(), g[f'img{i}_boxes']) and np.array_equal(dN[i, :k, 4:6].numpy(), g[f'img{i}_classes'])
        assert torch.equal(dF[i, :k], dN[i, :k])
    dpy.close()
