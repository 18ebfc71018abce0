"""ctypes binding of libayq.so (include/ayq.h) + a thin torch-facing Engine.

PyTorch is used for device memory and streams only; every arithmetic step of the hot path runs in
the hand-written CUDA kernels behind the C ABI.  There is no CPU fallback: constructing an Engine
without the library or without a CUDA device raises.
"""
import ctypes
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# AYQ_ROLE_PROF=1 (role-level cycle counters of the conv kernel) needs the profiling build of the same sources
LIB_PATH = os.environ.get('AYQ_LIB') or os.path.join(HERE, 'libayq_prof.so' if os.environ.get('AYQ_ROLE_PROF') else 'libayq.so')
# Test build: the product sources + the two cross-check convolution families (dp4a, cp.async-fed tcgen05).  Only tests load it.
TEST_LIB_PATH = os.path.join(HERE, 'libayq_test.so')
MAX_DET, DET_STRIDE, ANCHORS = 300, 6, 8400

_c = ctypes
_vp, _int, _sz = _c.c_void_p, _c.c_int, _c.c_size_t
# name -> (restype, argtypes); must list every symbol include/ayq.h declares (tests/test_abi.py checks)
SIGNATURES = {
    'ayq_last_error': (_c.c_char_p, []),
    'ayq_version': (_int, []),
    'ayq_create': (_int, [_vp, _sz, _int, _c.POINTER(_vp)]),
    'ayq_destroy': (_int, [_vp]),
    'ayq_set_max_batch': (_int, [_vp, _int]),
    'ayq_workspace_bytes': (_sz, [_vp]),
    'ayq_forward': (_int, [_vp, _vp, _int, _vp, _vp, _vp, _vp]),
    'ayq_forward_u8': (_int, [_vp, _vp, _int, _vp, _vp, _vp, _vp]),
    'ayq_forward_host': (_int, [_vp, _vp, _int, _vp, _vp]),
    'ayq_forward_host_u8': (_int, [_vp, _vp, _int, _vp, _vp]),
    'ayq_forward_host_async': (_int, [_vp, _vp, _int, _int, _vp, _vp]),
    'ayq_wait': (_int, [_vp]),
    'ayq_get_conv_impls': (_int, [_vp, _vp, _int]),
    'ayq_get_conv_variants': (_int, [_vp, _vp, _int]),
    'ayq_check_guards': (_int, [_vp]),
    'ayq_export_buffer': (_int, [_vp, _int, _int, _vp, _vp]),
    'ayq_buffer_shape': (_int, [_vp, _int, _c.POINTER(_int), _c.POINTER(_int), _c.POINTER(_int)]),
    'ayq_export_acc_tap': (_int, [_vp, _int, _int, _vp, _vp]),
    'ayq_launches_per_pass': (_int, [_vp]),
    'ayq_set_conv_impl': (_int, [_vp, _int]),
    'ayq_set_profiling': (_int, [_vp, _int]),
    'ayq_get_op_times': (_int, [_vp, _vp, _vp, _int]),
    'ayq_requantize_f32': (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _vp]),
    'ayq_silu_f32': (_int, [_vp, _vp, _vp, _vp, _int, _int, _int, _int, _vp]),
    'ayq_lut_f32': (_int, [_vp, _vp, _vp, _int, _int, _sz, _vp]),
    'ayq_quant_input_f32': (_int, [_vp, _vp, _vp, _vp, _int, _sz, _int, _vp]),
    'ayq_absmax_f32': (_int, [_vp, _vp, _int, _sz, _vp]),
    'ayq_nms': (_int, [_vp, _vp, _int, _vp, _vp, _vp]),
    'ayq_nms_boxes': (_int, [_vp, _vp, _int, _vp, _vp, _vp]),
    'ayq_coord_float': (_int, [_vp, _vp, _int, _vp, _vp, _vp]),
    'ayq_calib_conv_f32': (_int, [_vp, _vp, _vp, _vp, _vp, _int, _int, _int, _int, _int, _int, _int, _vp]),
    'ayq_calib_silu_f32': (_int, [_vp, _sz, _vp]),
    'ayq_calib_maxpool5_f32': (_int, [_vp, _vp, _int, _int, _int, _vp]),
    'ayq_calib_upsample2_f32': (_int, [_vp, _vp, _int, _int, _int, _vp]),
    'ayq_quant_weights_f32': (_int, [_vp, _vp, _int, _sz, _int, _c.c_double, _vp, _vp, _vp, _vp]),
}

_LIBS = {}


class AyqError(RuntimeError):
    pass


def load_library(path=None):
    """dlopen libayq.so (or the library at `path`) and bind every entry point.  Raises if the library is missing (build it with
    `python -m alpha_yolo_quant_b200.build`); never falls back to anything else."""
    path = os.path.abspath(path or LIB_PATH)
    if path in _LIBS:
        return _LIBS[path]
    if not os.path.exists(path):
        raise AyqError(f'{path} not found: build the CUDA extension first (python -m alpha_yolo_quant_b200.build). '
                       'There is no CPU fallback.')
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIBS[path] = lib
    return lib


def check(rc, lib=None):
    if rc < 0:
        raise AyqError(f'libayq error {rc}: {(lib or load_library()).ayq_last_error().decode()}')
    return rc


def _stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t, name):
    if not t.is_cuda:
        raise AyqError(f'{name}: expected a CUDA tensor, got {t.device} (no CPU fallback)')


class Engine:
    """One handle per GPU (not thread-safe), created from a compiled plan (plan.compile_plan)."""

    def __init__(self, plan, device=0, max_batch=512, lib_path=None):
        self.lib = load_library(lib_path)
        if not torch.cuda.is_available():
            raise AyqError('Engine: no CUDA device available; the integer YOLOv8n path has no CPU fallback')
        self.plan = plan
        self.device = torch.device('cuda', device if isinstance(device, int) else torch.device(device).index or 0)
        self._h = _vp()
        blob = plan.blob
        self._ck(self.lib.ayq_create(blob, len(blob), self.device.index, ctypes.byref(self._h)))
        self._ck(self.lib.ayq_set_max_batch(self._h, max_batch))
        self.max_batch = max_batch

    def _ck(self, rc):
        return check(rc, self.lib)

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self.lib.ayq_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- hot path
    def forward(self, img, want_dbox_cls=False):
        """img: CUDA float32 (n,3,640,640) in [0,1], or CUDA uint8 (n,3,640,640) (ToTensor then runs inside Conv_P1).
        Returns (dets (n,300,6), counts (n) int32[, dbox_cls (n,84,8400)])."""
        _require_cuda(img, 'Engine.forward')
        if img.dtype not in (torch.float32, torch.uint8) or img.dim() != 4 or tuple(img.shape[1:]) != (3, 640, 640):
            raise AyqError(f'Engine.forward: expected float32 or uint8 (n,3,640,640), got {img.dtype} {tuple(img.shape)}')
        img = img.contiguous()
        n = img.shape[0]
        with torch.cuda.device(self.device):
            dets = torch.empty((n, MAX_DET, DET_STRIDE), dtype=torch.float32, device=self.device)
            counts = torch.empty((n,), dtype=torch.int32, device=self.device)
            dbc = torch.empty((n, 84, ANCHORS), dtype=torch.float32, device=self.device) if want_dbox_cls else None
            fn = self.lib.ayq_forward_u8 if img.dtype == torch.uint8 else self.lib.ayq_forward
            self._ck(fn(self._h, img.data_ptr(), n, dbc.data_ptr() if dbc is not None else None,
                        dets.data_ptr(), counts.data_ptr(), _stream_ptr(self.device)))
        return (dets, counts, dbc) if want_dbox_cls else (dets, counts)

    def forward_into(self, img, dets, counts):
        """Allocation-free variant for timing loops (img: contiguous CUDA float32 or uint8 (n,3,640,640))."""
        fn = self.lib.ayq_forward_u8 if img.dtype == torch.uint8 else self.lib.ayq_forward
        self._ck(fn(self._h, img.data_ptr(), img.shape[0], None, dets.data_ptr(), counts.data_ptr(), _stream_ptr(self.device)))

    def forward_host(self, img_host, dets_host=None, counts_host=None):
        """img_host: CPU float32 or uint8 tensor (n,3,640,640) (pinned for full speed).  Synchronous."""
        if img_host.is_cuda:
            raise AyqError('forward_host takes host tensors')
        img_host = img_host.contiguous()
        n = img_host.shape[0]
        if dets_host is None:
            dets_host = torch.empty((n, MAX_DET, DET_STRIDE), dtype=torch.float32).pin_memory()
            counts_host = torch.empty((n,), dtype=torch.int32).pin_memory()
        fn = {torch.float32: self.lib.ayq_forward_host, torch.uint8: self.lib.ayq_forward_host_u8}.get(img_host.dtype)
        if fn is None:
            raise AyqError(f'forward_host: float32 or uint8 images, got {img_host.dtype}')
        self._ck(fn(self._h, img_host.data_ptr(), n, dets_host.data_ptr(), counts_host.data_ptr()))
        return dets_host, counts_host

    def forward_host_async(self, img_host, dets_host, counts_host):
        """Enqueue forward_host without waiting (ayq_forward_host_async); the three host tensors must stay alive and untouched
        until wait().  Several calls may be queued (each with its own buffers): the upload of one overlaps the kernels of the
        previous one."""
        if img_host.is_cuda or dets_host.is_cuda or counts_host.is_cuda:
            raise AyqError('forward_host_async takes host tensors')
        if not img_host.is_contiguous():
            raise AyqError('forward_host_async: contiguous host images required (the copy is asynchronous)')
        if img_host.dtype not in (torch.float32, torch.uint8):
            raise AyqError(f'forward_host_async: float32 or uint8 images, got {img_host.dtype}')
        n = img_host.shape[0]
        if tuple(dets_host.shape) != (n, MAX_DET, DET_STRIDE) or dets_host.dtype != torch.float32 or counts_host.dtype != torch.int32 or counts_host.numel() != n:
            raise AyqError('forward_host_async: dets_host float32 (n,300,6) and counts_host int32 (n) required')
        self._ck(self.lib.ayq_forward_host_async(self._h, img_host.data_ptr(), 1 if img_host.dtype == torch.uint8 else 0, n,
                                                 dets_host.data_ptr(), counts_host.data_ptr()))

    def wait(self):
        """Block until every queued forward_host_async call has delivered its results."""
        self._ck(self.lib.ayq_wait(self._h))

    def check_guards(self):
        """Canary bytes overwritten behind the activation buffers (needs AYQ_WS_GUARD=1 at engine creation); 0 = clean."""
        return self._ck(self.lib.ayq_check_guards(self._h))

    def conv_impls(self):
        """Per plan op: 2 / 1 / 0 = the conv kernel family that ran it in the last pass (2 = TMA-fed tcgen05), -2 = not a conv."""
        n = self.plan.n_ops
        out = np.zeros(n, np.int32)
        self._ck(self.lib.ayq_get_conv_impls(self._h, out.ctypes.data, n))
        return out

    def conv_variants(self):
        """Per plan op: the launch-plan variant the load-time tuner picked (0 default, 1 one chain per pipeline, 2 one accumulator
        per epilogue group; -1 not built yet, -2 not a conv).  All variants are bit-identical."""
        n = self.plan.n_ops
        out = np.zeros(n, np.int32)
        self._ck(self.lib.ayq_get_conv_variants(self._h, out.ctypes.data, n))
        return out

    def nms(self, dbox_cls):
        _require_cuda(dbox_cls, 'Engine.nms')
        dbox_cls = dbox_cls.contiguous().float()
        n = dbox_cls.shape[0]
        assert tuple(dbox_cls.shape[1:]) == (84, ANCHORS)
        dets = torch.empty((n, MAX_DET, DET_STRIDE), dtype=torch.float32, device=dbox_cls.device)
        counts = torch.empty((n,), dtype=torch.int32, device=dbox_cls.device)
        self._ck(self.lib.ayq_nms(self._h, dbox_cls.data_ptr(), n, dets.data_ptr(), counts.data_ptr(), _stream_ptr(self.device)))
        return dets, counts

    def coord_float(self, dbox_cls):
        """coord() of stage_8_torch.py on a (n,84,8400) float prediction tensor -> (dets (n,300,6), counts (n))."""
        _require_cuda(dbox_cls, 'Engine.coord_float')
        dbox_cls = dbox_cls.contiguous().float()
        n = dbox_cls.shape[0]
        assert tuple(dbox_cls.shape[1:]) == (84, ANCHORS)
        dets = torch.empty((n, MAX_DET, DET_STRIDE), dtype=torch.float32, device=dbox_cls.device)
        counts = torch.empty((n,), dtype=torch.int32, device=dbox_cls.device)
        self._ck(self.lib.ayq_coord_float(self._h, dbox_cls.data_ptr(), n, dets.data_ptr(), counts.data_ptr(), _stream_ptr(self.device)))
        return dets, counts

    # -- taps / introspection
    def buffer_shape(self, buf):
        c, h, w = _int(), _int(), _int()
        self._ck(self.lib.ayq_buffer_shape(self._h, buf, ctypes.byref(c), ctypes.byref(h), ctypes.byref(w)))
        return c.value, h.value, w.value

    def export_buffer(self, buf, n):
        c, h, w = self.buffer_shape(buf)
        out = torch.empty((n, c, h, w), dtype=torch.int32, device=self.device)
        self._ck(self.lib.ayq_export_buffer(self._h, buf, n, out.data_ptr(), _stream_ptr(self.device)))
        if buf in self.plan.info.get('ps_bufs', ()):
            # phase-split buffer [(y&1)*2 + (x&1)][plane][n][h][w][16] -> (n, C, 2h, 2w)
            out = out.view(n, 2, 2, c // 4, h, w).permute(0, 3, 4, 1, 5, 2).reshape(n, c // 4, 2 * h, 2 * w).contiguous()
        return out

    def export_acc_tap(self, tap, n):
        name, c, h, w = self.plan.info['acc_taps'][tap]
        out = torch.empty((n, c, h, w), dtype=torch.int32, device=self.device)
        self._ck(self.lib.ayq_export_acc_tap(self._h, tap, n, out.data_ptr(), _stream_ptr(self.device)))
        return out

    def set_conv_impl(self, impl):
        self._ck(self.lib.ayq_set_conv_impl(self._h, {'dp4a': 0, 'tcgen05': 1, 'tma': 2}.get(impl, impl)))

    def set_max_batch(self, mb):
        self._ck(self.lib.ayq_set_max_batch(self._h, mb))
        self.max_batch = mb

    def set_profiling(self, on):
        self._ck(self.lib.ayq_set_profiling(self._h, 1 if on else 0))

    def op_times(self):
        n = self.plan.n_ops + 1
        ms = np.zeros(n, np.float32)
        calls = np.zeros(n, np.int32)
        self._ck(self.lib.ayq_get_op_times(self._h, ms.ctypes.data, calls.ctypes.data, n))
        return ms, calls

    @property
    def launches_per_pass(self):
        return self._ck(self.lib.ayq_launches_per_pass(self._h))

    @property
    def workspace_bytes(self):
        return self.lib.ayq_workspace_bytes(self._h)


def unpack_detections(dets, counts):
    """(n,300,6), (n) -> list of (boxes (k,4), classes (k,2)) or (None, None), the reference's return value
    (stage_8_torch_full_quant.py:1267-1275; convert_res :426-429)."""
    out = []
    cnt = counts.tolist()
    for i, k in enumerate(cnt):
        if k == 0:
            out.append((None, None))
        else:
            out.append((dets[i, :k, :4], dets[i, :k, 4:6]))
    return out
