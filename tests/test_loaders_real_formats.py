"""The weight / scale / max_a loaders on the reference's REAL on-disk formats.

tests/golden/main_dir_k8.tar.xz holds files written by the unmodified reference pipeline in the build container
(oracle/ref_harness.py --export-main-dir): `8_nano/results/QUANT_WEIGHTS_8.pickle` (stage_7.py:780, torch.save of the 127-key
state_dict), `8_nano/bias_scales/*_scale.pickle` (utils/save_weights.py:24-30, gzip + pickle of numpy arrays) and
`8_nano/results/max_a.txt` (stage_5).  The compact workload_k8.npz was exported from the same run, so both routes must give
the same tensors and the same compiled plan."""
import os
import tarfile

import numpy as np
import pytest
import torch

from alpha_yolo_quant_b200 import loaders, plan


@pytest.fixture(scope='module')
def main_dir(golden_dir, tmp_path_factory):
    d = tmp_path_factory.mktemp('main_dir')
    with tarfile.open(os.path.join(golden_dir, 'main_dir_k8.tar.xz')) as tf:
        tf.extractall(d, filter='data')
    return str(d / '8_nano')


def test_load_main_dir_equals_fixture_workload(golden_dir, main_dir):
    sd, scales, ma = loaders.load_main_dir(main_dir, 8)
    K, sd2, scales2, ma2 = loaders.load_workload_npz(os.path.join(golden_dir, 'workload_k8.npz'))
    assert len(sd) == 127 and list(sd.keys()) == list(sd2.keys())
    for k in sd:
        assert sd[k].dtype == torch.float32 and torch.equal(sd[k], sd2[k]), k
    assert sorted(scales) == sorted(scales2) and len(scales) == 64
    for k in scales:
        assert scales[k].dtype == torch.float32
        assert torch.equal(scales[k].reshape(-1), scales2[k].reshape(-1)), k
    assert ma == ma2 and len(ma) >= 60
    # reference-named single-file loaders (utils/save_weights.py:32-42, utils/max_a.py:1-7)
    one = loaders.load_scale(main_dir, 'Conv_P1_scale.pickle')
    assert np.asarray(one).reshape(-1).shape == (16,)
    assert loaders.max_a(os.path.join(main_dir, 'results', 'max_a.txt')) == ma


def test_plan_from_main_dir_is_byte_identical(golden_dir, main_dir):
    sd, scales, ma = loaders.load_main_dir(main_dir, 8)
    K, sd2, scales2, ma2 = loaders.load_workload_npz(os.path.join(golden_dir, 'workload_k8.npz'))
    assert plan.compile_plan(sd, scales, ma, 8).blob == plan.compile_plan(sd2, scales2, ma2, K).blob


@pytest.mark.gpu
def test_configure_main_dir_drop_in_on_gpu(golden_dir, main_dir):
    """The reference's driver lines (stage_8_torch_full_quant.py:1278-1294) reading the reference's own files."""
    from alpha_yolo_quant_b200 import stage_8_torch_full_quant as S
    from oracle import synth
    g = np.load(os.path.join(golden_dir, 'golden_k8.npz'))
    S.configure(main_dir=main_dir, k=8)
    model = S.Yolov8().to('cuda')
    model.load_state_dict(torch.load(os.path.join(main_dir, 'results', 'QUANT_WEIGHTS_8.pickle')))
    model.eval()
    for i in (1, 2, 11):
        with torch.no_grad():
            boxes, classes = model(synth.to_input_tensor(synth.synth_image_u8(i)))
        if g[f'img{i}_boxes'].shape[0] == 0:
            assert boxes is None and classes is None
        else:
            assert np.array_equal(boxes.cpu().numpy(), g[f'img{i}_boxes']) and np.array_equal(classes.cpu().numpy(), g[f'img{i}_classes'])
