"""Target for compute-sanitizer (memcheck / racecheck / synccheck): a 2-image pass through every epilogue variant of the conv kernel
-- K = 8 production plan (MAGIC2 + FAST epilogues, phase-split stores, halo and nine-box feeds), K = 8 plan with accumulator taps
(generic epilogue), K = 6 production plan (WIDE epilogue) -- plus Conv_P1, pool, head and q_NMS, and the host entry.
    compute-sanitizer --tool memcheck python tools/sanitize_pass.py"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from alpha_yolo_quant_b200 import engine, loaders, plan  # noqa: E402
from oracle import synth  # noqa: E402

x = torch.from_numpy(synth.to_input_array([synth.synth_image_u8(1), synth.synth_image_u8(3)])).cuda()
u8 = torch.from_numpy(np.stack([synth.synth_image_u8(1), synth.synth_image_u8(3)])).pin_memory()
for k, taps in ((8, False), (8, True), (6, False)):
    K, sd, sc, ma = loaders.load_workload_npz(os.path.join(REPO, 'tests', 'golden', f'workload_k{k}.npz'))
    p = plan.compile_plan(sd, sc, ma, K, taps=taps)
    e = engine.Engine(p, 0, 2)
    dets, counts = e.forward(x)
    torch.cuda.synchronize()
    if not taps:
        dh, ch = e.forward_host(u8)
        assert torch.equal(ch, counts.cpu())
    print(f'K={k} taps={taps}: detections {counts.tolist()} conv impls all TMA: {bool((e.conv_impls()[e.conv_impls() != -2] == 2).all())}', flush=True)
    e.close()
print('sanitize_pass ok')
