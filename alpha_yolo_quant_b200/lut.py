"""Host-side constant tables of the quantised layer library.

Mirrors the reference LUT builders (paths relative to /root/reference/quantisation/):
  create_sigmoid_lookup_table   utils/silu.py:32-50      (+ sigmoid :4-5, quantize :14-19, dequantize :22-29)
  create_exponent_lookup_table  utils/exponent.py:32-50
  scale                         utils/scale.py:4-5

The tables are tiny host constants (255, 256 and 65 535 entries) that end up in the plan blob; the
device only ever indexes them.  Values follow the reference's evaluation under numpy 2.x dtype rules
(a float32 array divided by a python float stays float32, so the sigmoid/exp argument is a float32
scalar and the result is a float32) -- SURVEY.md hard part 4.  Scalars are evaluated one by one, as
the reference does: numpy's scalar and vectorised float32 pow differ in the last ulp for a handful
of the 16-bit entries.
"""
import os

import numpy as np


def scale(a, k):
    """utils/scale.py:4-5"""
    return (2 ** (k - 1) - 1) / a


def _dequantize(i, max_val, bits):
    arr = np.array((i,)).astype(np.float32)
    s = (2 ** (bits - 1) - 1) / max_val
    if s > 0:
        arr /= s
    else:
        arr[...] = 0
    return arr[0]


def _quantize(arr, bits):
    m = 2 ** (bits - 1) - 1
    return np.clip(np.round(arr * (m / 1)), -m, m)[0]


def _write_table(path, title, table):
    d = os.path.dirname(path)
    if d and not os.path.isdir(d):
        return                      # the reference writes relative to its cwd; only mirror it when utils/ exists
    with open(path, 'w') as f:
        f.write(f'// {title}\n\n')
        for key, value in table.items():
            f.write(f'{key} = {value}\n')


def create_sigmoid_lookup_table(max_conv_value, bit_size_act, write_txt=False):
    """dict {i: round(sigmoid(i * max_conv_value / M) * M)} for i in [-M, M], M = 2^(bits-1)-1."""
    m = 2 ** (bit_size_act - 1) - 1
    table = {}
    for i in range(-m, m + 1):
        d = _dequantize(i, max_conv_value, bit_size_act)
        table[i] = _quantize(np.array((1 / (1 + (np.e ** (-d))),)), bit_size_act)
    if write_txt:
        _write_table(f'utils/sigmoid_table_{bit_size_act}_bit.txt', f'SIGMOID TABLE FOR {bit_size_act} BIT', table)
    return table


def create_exponent_lookup_table(max_conv_value, bit_size_act, write_txt=False):
    """dict {i: round(exp(i * max_conv_value / M) * M)} for i in [-(2^bits-1), 0]."""
    top = 2 ** bit_size_act - 1
    table = {}
    for i in range(-top, 1):
        d = _dequantize(i, max_conv_value, bit_size_act)
        table[i] = _quantize(np.array((np.exp(d),)), bit_size_act)
    if write_txt:
        _write_table(f'utils/exponent_table_{bit_size_act}_bit.txt', f'EXPONENT TABLE FOR {bit_size_act} BIT', table)
    return table


def table_to_array(table):
    """dict -> (key_min, float32 array indexed by key - key_min); keys must be contiguous integers."""
    keys = sorted(table.keys())
    assert keys == list(range(keys[0], keys[-1] + 1))
    return keys[0], np.array([table[k] for k in keys], dtype=np.float32)


_CACHE = {}


def cached_array(kind, max_conv_value, bits):
    """Memoised (key_min, array) -- the 16-bit table takes ~1 s to build."""
    key = (kind, float(max_conv_value), int(bits))
    if key not in _CACHE:
        fn = create_sigmoid_lookup_table if kind == 'sigmoid' else create_exponent_lookup_table
        _CACHE[key] = table_to_array(fn(max_conv_value, bits))
    return _CACHE[key]
