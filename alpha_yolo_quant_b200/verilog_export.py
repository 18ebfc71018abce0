"""Verilog test-vector export: the text dumps of /root/reference/quantisation/utils/save_weights.py
-- weights / biases (:90-110), activations (:113-127) and rescale / shift coefficients (:130-155) in the
`<width>'b<bits>; // <value>` format of bit_converter (:45-71) -- written from arrays the engine produces (per-layer taps via
Engine.export_buffer / export_acc_tap, plan coefficients) or from the weight quantiser.  Same file names, same directory
layout, byte-identical content; the formatting is table driven (one string per possible value), which is what turns the
reference's minutes of per-element Python into seconds.  Pure host code: numpy in, text files out.
"""
import os

import numpy as np


def bit_converter(final_file_name, k, value, element):
    """The literal format of utils/save_weights.py:45-71: `<sign><width>'b<magnitude bits, zero padded to width>`.
    width = 18 for a bias, k for a rescale / shift coefficient (never signed), k - 1 for a weight or an activation
    (sign-magnitude).  A magnitude that does not fit is reported on stdout like the reference does, and printed unpadded."""
    value = int(value)
    magnitude = format(abs(value), 'b')
    if element == 'bias':
        width, signed, what = 18, True, 'BIAS MORE THAN 18 BIT!'
    elif element == 'rescale':
        width, signed, what = k, False, f'RESCALE MORE THAN {k} BIT!'
    else:
        width, signed, what = k - 1, True, f'MORE THAN {k} BIT!'
    if len(magnitude) > width:
        print(f'{what} {magnitude} {final_file_name}')
    sign = '-' if signed and value < 0 else ''
    return f"{sign}{width}'b{magnitude.rjust(width, '0')}"


def _lines(tag, values, k, element, first_index, final_file_name):
    """['tag[i] = code; // v\\n', ...] for a flat integer array, one table lookup per element"""
    v = np.asarray(values).reshape(-1).astype(np.int64)
    if v.size == 0:
        return []
    lo, hi = int(v.min()), int(v.max())
    if hi - lo <= 1 << 20:
        table = [f"{bit_converter(final_file_name, k, x, element)}; // {x}\n" for x in range(lo, hi + 1)]
        return [f'{tag}[{first_index + i}] = {table[x - lo]}' for i, x in enumerate(v.tolist())]
    return [f'{tag}[{first_index + i}] = {bit_converter(final_file_name, k, x, element)}; // {x}\n' for i, x in enumerate(v.tolist())]


def save_txt_weight(conv, bias, file_name, type, k, dir_names):
    """:90-110.  conv (C, Cin, kh, kw) ints, bias (1, C, 1, 1) ints -> quant_weights_yolov8n/<name>.txt"""
    conv, bias = np.asarray(conv), np.asarray(bias)
    final_file_name = f'{file_name}_type_{type}_bit_{k}_shape_{conv.shape}'
    per_ch = conv.shape[2] * conv.shape[3]
    out, i = [], 0
    for batch in range(conv.shape[0]):
        out.append(f'\n//   Batch: {batch}\n\n')
        for channel in range(conv.shape[1]):
            out += _lines('weight', conv[batch, channel], k, 'weight', i, final_file_name)
            i += per_ch
            out.append('\n')
    out.append('\n\n')
    out += _lines('weight_bias', bias, k, 'bias', 0, final_file_name)
    with open(os.path.join(dir_names, 'quant_weights_yolov8n', final_file_name + '.txt'), 'w') as f:
        f.write(''.join(out))


def _activation_path(file_name, type, k, shape, silu):
    return f"quant_activations/{'silu' if silu else 'conv2d'}/{file_name}_type_{type}_bit_{k}_shape_{tuple(shape)}"


def save_txt_activations(arr, file_name, dir_names, type, k, silu=False):
    """:113-127.  arr (N, C, H, W) ints (e.g. Engine.export_buffer(...).cpu().numpy())"""
    arr = np.asarray(arr)
    final_file_name = _activation_path(file_name, type, k, arr.shape, silu)
    per_ch = arr.shape[2] * arr.shape[3]
    out, i = [], 0
    for batch in range(arr.shape[0]):
        for channel in range(arr.shape[1]):
            out.append(f'\n//   Channel: {channel}\n\n')
            out += _lines('pixel', arr[batch, channel], k, 'activ', i, final_file_name)
            i += per_ch
            out.append('\n')
    with open(os.path.join(dir_names, final_file_name + '.txt'), 'w') as f:
        f.write(''.join(out))


def save_txt_rescale_shift(conv, rescale, shift, file_name, dir_names, type, k, silu=False):
    """:130-155: APPENDS the per-channel (or scalar) rescale coefficients and shifts to the activation file of `conv`."""
    conv = np.asarray(conv)
    final_file_name = _activation_path(file_name, type, k, conv.shape, silu)
    r, s = np.asarray(rescale), np.asarray(shift)
    if r.ndim != 4:                                                # scalar coefficients (:147-151)
        r, s = r.reshape(1, 1, 1, 1), s.reshape(1, 1, 1, 1)
    rv, sv = r[0, :, 0, 0], s[0, :, 0, 0]
    out = ['\n'] + [f'rescale[{i}] = {bit_converter(final_file_name, k, x, "rescale")}; // {x}\n' for i, x in enumerate(rv.tolist())]
    out += ['\n'] + [f'shift[{i}] = {bit_converter(final_file_name, k, x, "rescale")}; // {x}\n' for i, x in enumerate(sv.tolist())]
    with open(os.path.join(dir_names, final_file_name + '.txt'), 'a') as f:
        f.write(''.join(out))


def export_layer(engine, plan, layer, n_images, dir_names, k):
    """Dump the SiLU output of `layer` for the first image of the last forward pass, straight from the engine's buffers."""
    meta = plan.info['layers'][layer]
    buf = meta.get('silu_buf', meta.get('ps_buf'))
    if buf is None:
        raise KeyError(f'{layer}: the plan keeps no plain SiLU output (compile it with taps=True)')
    arr = engine.export_buffer(buf, n_images)[:1].cpu().numpy().astype(np.int64)
    save_txt_activations(arr, layer, dir_names, type='act_silu', k=k, silu=True)
    return arr
