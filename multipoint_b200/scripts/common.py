"""Shared plumbing of the three entry points (reference: predict_keypoints.py:30-57,
predict_align_image_pair.py:38-67, export_keypoints.py:23-60): YAML config with the model section
overridden by ``<model-dir>/params.yaml``, device choice, network construction and weight loading.

The reference reads image pairs from HDF5 through h5py, which is not installed here (and the
36 GB dataset is not present), so samples come from ``--input file.npz`` (arrays ``optical`` and
``thermal`` of shape (N,H,W) or (N,1,H,W) in [0,1], optional ``names``) or, without it, from the
seeded synthetic generator.  Results are written as ``.npz`` (the reference writes ``np.save`` dicts
and an h5 group per sample with dataset ``keypoints``; the npz uses the same key layout).
"""
import os

import numpy as np
import torch
import yaml

from .. import synthetic
from ..models import MultiPoint
from ..utils import fix_model_weigth_keys

DEFAULT_PREDICTION = {'allow_gpu': True, 'batchsize': 1, 'detection_threshold': 0.015, 'nms': 4, 'cpu_nms': True, 'topk': 0,
                      'reprojection_threshold': 3,
                      'matching': {'method': 'bfmatcher', 'method_kwargs': {'crossCheck': True}, 'knn_matches': False}}


def load_config(yaml_config, model_dir):
    config = {'prediction': dict(DEFAULT_PREDICTION), 'model': {'type': 'MultiPoint'}}
    if yaml_config and os.path.exists(yaml_config):
        with open(yaml_config, 'r') as f:
            config.update(yaml.load(f, Loader=yaml.FullLoader))
    params = os.path.join(model_dir, 'params.yaml') if model_dir else None
    if params and os.path.exists(params):
        with open(params, 'r') as f:
            config['model'] = yaml.load(f, Loader=yaml.FullLoader)['model']   # overwrite the model params
    return config


def build_network(config, model_dir, version, device):
    model_cfg = {k: v for k, v in config['model'].items() if k in MultiPoint.default_config}
    net = MultiPoint(model_cfg)
    if version != 'none':
        weights = torch.load(os.path.join(model_dir, version + '.model'), map_location=torch.device('cpu'))
        net.load_state_dict(fix_model_weigth_keys(weights))
    return net.to(device).eval()


def load_samples(path, count, seed, height=512, width=640):
    """-> list of dicts shaped like ImagePairDataset items (ImagePairDataset.py:173-241), batched (N,...)."""
    if path:
        z = np.load(path)
        opt, th = z['optical'].astype(np.float32), z['thermal'].astype(np.float32)
        if opt.ndim == 3:
            opt, th = opt[:, None], th[:, None]
        names = [str(n) for n in z['names']] if 'names' in z.files else ['sample_%d' % i for i in range(len(opt))]
        ones = np.ones(opt.shape, bool)
        batch = {'optical': {'image': opt, 'valid_mask': ones, 'is_optical': np.ones((len(opt), 1), bool)},
                 'thermal': {'image': th, 'valid_mask': ones.copy(), 'is_optical': np.zeros((len(opt), 1), bool)}}
    else:
        batch = synthetic.image_pair_batch(seed, count, height, width)
        names = ['synthetic_%d' % i for i in range(count)]
    data = {s: {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in batch[s].items()} for s in ('optical', 'thermal')}
    data['name'] = names
    return data


def require_cuda(config):
    if not (config['prediction'].get('allow_gpu', True) and torch.cuda.is_available()):
        raise RuntimeError("multipoint_b200 entry points need a CUDA device (allow_gpu: true); there is no CPU path")
    return torch.device("cuda:0")
