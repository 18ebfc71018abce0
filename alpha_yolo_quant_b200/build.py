"""Builds libayq.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m alpha_yolo_quant_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libayq.so')
STAMP = os.path.join(HERE, '.libayq.stamp')
LIB_PROF = os.path.join(HERE, 'libayq_prof.so')        # same sources with -DAYQ_ROLE_PROF_BUILD (role-level cycle counters)
STAMP_PROF = os.path.join(HERE, '.libayq_prof.stamp')
LIB_TEST = os.path.join(HERE, 'libayq_test.so')        # same sources with -DAYQ_TEST_BUILD: + the dp4a / cp.async-fed cross-check conv kernels
STAMP_TEST = os.path.join(HERE, '.libayq_test.stamp')
SOURCES = ['ayq.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--fmad=false',            # the reference multiplies and adds in separate fp32 ops (SURVEY hard part 1)
              '-Xcompiler', '-fPIC', '-shared', '-cudart', 'static']


def _digest():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join('..', '..', 'include', 'ayq.h')]
    for name in files:
        with open(os.path.join(CSRC, name), 'rb') as f:
            h.update(name.encode())
            h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def build(force=False, verbose=False, prof=False, test=False):
    """prof=True builds the profiling variant libayq_prof.so (AYQ_ROLE_PROF=1 loads it instead of libayq.so); test=True builds
    libayq_test.so, the product sources plus the two cross-check convolution families (tests only, never the product path)."""
    dig = _digest()
    lib, stamp = (LIB_PROF, STAMP_PROF) if prof else ((LIB_TEST, STAMP_TEST) if test else (LIB, STAMP))
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return lib
    cmd = [nvcc_path()] + NVCC_FLAGS + (['-DAYQ_ROLE_PROF_BUILD'] if prof else []) + (['-DAYQ_TEST_BUILD'] if test else []) + \
          (['-Xptxas', '-v'] if verbose else []) + \
          [os.path.join(CSRC, s) for s in SOURCES] + ['-o', lib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    with open(stamp, 'w') as f:
        f.write(dig)
    return lib


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, prof='--prof' in sys.argv, test='--test' in sys.argv))
