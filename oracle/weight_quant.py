"""CPU oracle of the weight quantiser -- TEST INFRASTRUCTURE (see oracle/yolo_int.py for who may import it).

Restates conv_quant() of /root/reference/quantisation/stage_6_full_quant.py:89-126 without its
text dumps, convolution and file writes:
    utils/quant_matrix.py:56-78   per OUTPUT CHANNEL (the leading dimension of the weight): a = max|w|, s = (2^(k-1)-1)/a,
                                  q = int64(round_half_even(clip(w, -a, a) * s)); the scale is kept in a float64 array
    utils/quant_bias.py:2-4       bias_q = int64(bias * bias_scale)   (truncation towards zero)
    stage_6_full_quant.py:90-96   bias_scale = scale_input * conv_scale (the first layer: 127 * conv_scale, `start=True`)
    :122-123                      scale_res = bias_scale as (1, C, 1, 1): what bias_scales/<layer>_scale.pickle holds
dtypes follow numpy >= 2 (NEP 50), under which tests/golden/golden_wquant_k8.npz was recorded from the unmodified script
(oracle/ref_harness.py --weight-quant): a() and the scale are float32, the product w * s is float32, the bias product is
float64 (float32 array times a float64 scalar).  Pinned by tests/test_oracle_golden.py.
"""
import numpy as np


def quant_matrix(matrix, k):
    """utils/quant_matrix.py:56-78 (start=False). matrix (C, ...) float32 -> (int64 array, scales float64 (C, 1))"""
    res = np.zeros(matrix.shape)
    scales = np.zeros((matrix.shape[0], 1))
    for c in range(matrix.shape[0]):
        a = np.abs(matrix[c]).max()                                    # utils/a.py:4-5  (np.float32)
        m = np.clip(matrix[c], -a, a)                                  # new_clip :50-53 (identity for a = max|w|)
        s = (2 ** (k - 1) - 1) / a                                     # utils/scale.py:4-5  (np.float32 under NEP 50)
        scales[c, :] += s
        res[c] = np.int64(np.round(m * s))
    return np.int64(res), scales


def conv_quant(conv, bias_conv, scale_input, k, start=False):
    """-> (weights int64 (C, Cin, kh, kw), bias int64 (1, C, 1, 1), scale_res float64 (1, C, 1, 1))"""
    q, conv_scale = quant_matrix(np.array(conv, copy=True), k)
    conv_scale = np.transpose(conv_scale)                              # (1, C)
    if start:
        input_scale = np.zeros((1, 1)) + (2 ** (k - 1) - 1) / 1        # quant_matrix(input, k, start=True): a = 1
        bias_scale = np.dot(input_scale, conv_scale)
    else:
        bias_scale = scale_input * conv_scale
    b = np.array(bias_conv).transpose(1, 0, 2, 3)                      # (1, C, 1, 1)
    out = np.zeros(b.shape)
    for i in range(b.shape[0]):
        for c in range(b.shape[1]):
            out[i, c, :, :] += np.int64(b[i, c, :, :] * bias_scale[i, c])   # quant_bias
    return q, np.int64(out), np.expand_dims(bias_scale, (2, 3))
