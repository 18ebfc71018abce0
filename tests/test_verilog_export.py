"""Verilog test-vector export: byte-identical to the files the reference's utils/save_weights.py writes
(golden_verilog_txt.npz was produced by calling the reference functions in the build container)."""
import os

import numpy as np


def test_text_dumps_are_byte_identical(golden_dir, tmp_path):
    from alpha_yolo_quant_b200 import verilog_export as V
    g = np.load(os.path.join(golden_dir, 'golden_verilog_txt.npz'))
    d = str(tmp_path)
    for sub in ('quant_weights_yolov8n', 'quant_activations/conv2d', 'quant_activations/silu'):
        os.makedirs(os.path.join(d, sub))
    V.save_txt_weight(g['conv'], g['bias'], 'Conv_T', type='Conv2D', k=8, dir_names=d)
    V.save_txt_activations(g['arr'], 'Conv_T', d, type='act_conv', k=8)
    V.save_txt_activations(g['arr'], 'Conv_T', d, type='act_silu', k=8, silu=True)
    V.save_txt_rescale_shift(g['arr'], g['resc'], g['shift'], 'Conv_T', d, type='act_conv', k=8)
    V.save_txt_rescale_shift(g['arr'], np.int64(201), np.int64(17), 'Conv_T', d, type='act_silu', k=8, silu=True)
    V.save_txt_weight(g['conv6'], g['bias'][:, :2], 'Conv_K6', type='Conv2D', k=6, dir_names=d)
    assert len(g['names']) == 4
    for name, text in zip(g['names'], g['texts']):
        assert open(os.path.join(d, str(name))).read() == str(text), name


def test_bit_converter_edges():
    from alpha_yolo_quant_b200 import verilog_export as V
    assert V.bit_converter('f', 8, 5, 'weight') == "7'b0000101"
    assert V.bit_converter('f', 8, -127, 'activ') == "-7'b1111111"
    assert V.bit_converter('f', 8, 0, 'bias') == "18'b" + '0' * 18
    assert V.bit_converter('f', 8, -3, 'bias') == "-18'b" + '0' * 16 + '11'
    assert V.bit_converter('f', 8, 255, 'rescale') == "8'b11111111"
    assert V.bit_converter('f', 4, 7, 'weight') == "3'b111"
