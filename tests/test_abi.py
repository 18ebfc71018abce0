"""The C-ABI library loads and exports every symbol include/ayq.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

from alpha_yolo_quant_b200 import build, engine

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    build.build()
    return engine.load_library()


def _declared():
    txt = open(os.path.join(REPO, 'include', 'ayq.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(ayq_\w+)\s*\(', txt)))


def test_exports_match_header(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(engine.SIGNATURES) == names


def test_version_and_error_string(lib):
    from alpha_yolo_quant_b200 import plan
    assert lib.ayq_version() == plan.VERSION
    assert isinstance(lib.ayq_last_error(), bytes)


def test_bad_arguments_fail_loudly(lib):
    h = ctypes.c_void_p()
    assert lib.ayq_create(None, 0, 0, ctypes.byref(h)) < 0
    assert b'ayq_create' in lib.ayq_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback(lib, golden_dir):
    """Without a CUDA device the product path must raise, not compute."""
    from alpha_yolo_quant_b200 import loaders, plan
    from alpha_yolo_quant_b200 import stage_8_torch_full_quant as S
    K, sd, sc, ma = loaders.load_workload_npz(os.path.join(golden_dir, 'workload_k8.npz'))
    p = plan.compile_plan(sd, sc, ma, K)
    with pytest.raises(engine.AyqError):
        engine.Engine(p)
    h = ctypes.c_void_p()
    assert lib.ayq_create(p.blob, len(p.blob), 0, ctypes.byref(h)) < 0
    assert b'no CUDA device' in lib.ayq_last_error()
    S.configure(workload=os.path.join(golden_dir, 'workload_k8.npz'))
    m = S.Yolov8()
    m.load_state_dict(sd)
    with pytest.raises(engine.AyqError):
        m(torch.zeros(1, 3, 640, 640))
    with pytest.raises(engine.AyqError):
        S.requantize(torch.zeros(1, 2, 2, 2), 1.0, 2.0, 8, 'cpu')


def test_shim_state_dict_layout(golden_dir):
    import numpy as np
    from alpha_yolo_quant_b200 import stage_8_torch_full_quant as S
    z = np.load(os.path.join(golden_dir, 'workload_k8.npz'))
    m = S.Yolov8()
    assert list(m.state_dict().keys()) == [str(k) for k in z['sd_keys']]
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(z['sd/' + k].shape), k
