"""Loaders for the artefacts the reference hot path reads.

  load_scales(dir)   utils/save_weights.py:36-42   {layer: fp32 tensor (1,C,1,1)} from {dir}/bias_scales/*_scale.pickle
  max_a(path)        utils/max_a.py:1-7            {tap name: float} from results/max_a.txt
  load_quant_weights stage_8_torch_full_quant.py:1281 (torch.load of stage_7's QUANT_WEIGHTS_{K}.pickle)
  load_workload_npz  the compact fixture format used by this repo's tests (tests/golden/workload_k*.npz)
"""
import gzip
import os
import pickle

import numpy as np
import torch

from .plan import parse_max_a


def load_scale(dir_names, file_name):
    """utils/save_weights.py:32-33"""
    with gzip.open(f'{dir_names}/bias_scales/{file_name}', 'rb') as f:
        return pickle.load(f)


def load_scales(dir_names):
    all_scales = {}
    for file in os.listdir(os.path.join(dir_names, 'bias_scales')):
        file_name = file.split('_scale')[0]
        all_scales[file_name] = torch.from_numpy(np.asarray(load_scale(dir_names, file))).type(torch.float32)
    return all_scales


def max_a(filepath):
    with open(filepath, 'r') as f:
        return parse_max_a(f.read())


def load_quant_weights(path):
    return torch.load(path, map_location='cpu')


def load_main_dir(main_dir, K):
    """(state_dict, all_scales, max_a_dict) from a reference `{K}_nano/` directory."""
    sd = load_quant_weights(os.path.join(main_dir, 'results', f'QUANT_WEIGHTS_{K}.pickle'))
    return sd, load_scales(main_dir), max_a(os.path.join(main_dir, 'results', 'max_a.txt'))


def load_workload_npz(path):
    """(K, state_dict, all_scales, max_a_dict) from the test fixture format (written by oracle/ref_harness.py)."""
    z = np.load(path, allow_pickle=False)
    K = int(z['K'])
    sd = {str(name): torch.from_numpy(z['sd/' + str(name)].astype(np.float32)) for name in z['sd_keys']}
    scales = {str(name): torch.from_numpy(z['scale/' + str(name)].astype(np.float32)).reshape(1, -1, 1, 1) for name in z['scale_keys']}
    return K, sd, scales, parse_max_a(str(z['max_a_txt']))
