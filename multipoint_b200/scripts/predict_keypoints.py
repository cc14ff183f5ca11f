"""Entry point mirroring predict_keypoints.py (reference :14-145): batched keypoint prediction for
both spectra with box NMS on the 4-D batch.  ``-e`` (repeatability) and the plotting branches are
out of scope for this tier.

    python -m multipoint_b200.scripts.predict_keypoints -m model_weights/multipoint -v none -b --batchsize 8
"""
import argparse

import numpy as np
import torch

from .. import utils
from .common import build_network, load_config, load_samples, require_cuda


def main(argv=None):
    parser = argparse.ArgumentParser(description='Predict the keypoints of an image')
    parser.add_argument('-y', '--yaml-config', default='configs/config_image_pair_dataset_prediction.yaml', help='YAML config file')
    parser.add_argument('-m', '--model-dir', default='model_weights/multipoint', help='Directory of the model')
    parser.add_argument('-v', '--version', default='latest', help='Model version (name of the param file)')
    parser.add_argument('-i', '--index', default=0, type=int, help='Index of the sample to predict')
    parser.add_argument('-b', dest='batch', action='store_true', help='Predict a batch instead of a single image')
    parser.add_argument('--batchsize', default=None, type=int, help='Override prediction.batchsize')
    parser.add_argument('-mask', dest='mask', action='store_true', help='If set invalid image pixels will be set to 0')
    parser.add_argument('-s', '--seed', default=0, type=int, help='Seed of the random generators')
    parser.add_argument('--input', default=None, help='npz with optical/thermal arrays (default: synthetic)')
    parser.add_argument('-o', '--output', default=None, help='npz to write keypoints to')
    args = parser.parse_args(argv)

    config = load_config(args.yaml_config, args.model_dir)
    device = require_cuda(config)
    print('Predicting on device: {}'.format(device))
    net = build_network(config, args.model_dir, args.version, device)
    pred = config['prediction']
    bs = (args.batchsize or pred['batchsize']) if args.batch else 1
    data = load_samples(args.input, args.index + bs, args.seed)
    data = {s: {k: v[args.index:args.index + bs].to(device) for k, v in data[s].items()} for s in ('optical', 'thermal')}

    result = {}
    with torch.no_grad():
        for s in ('optical', 'thermal'):
            out = net(data[s])
            prob = out['prob'] * data[s]['valid_mask'] if args.mask else out['prob']
            if pred['nms'] > 0:
                # one fused call: dense NMS map + ordered keypoints (the torch.nonzero idiom :181-183)
                cap = pred['topk'] if pred['topk'] > 0 else None    # None = H*W: never truncates, whatever the box size
                dense, kp, sc, cnt = utils.box_nms_keypoints(prob, pred['nms'], pred['detection_threshold'],
                                                             keep_top_k=pred['topk'], kp_cap=cap)
            else:
                dense = prob
                kp, sc, cnt = __import__('multipoint_b200').ops.extract_keypoints(prob[:, 0].contiguous(), pred['detection_threshold'])
            cnt = cnt.cpu().numpy()
            result['prob_' + s] = dense.cpu().numpy()
            for b in range(bs):
                result['keypoints_%s_%d' % (s, b)] = kp[b, :cnt[b]].cpu().numpy()
            print('{}: {} keypoints per image'.format(s, cnt.tolist()))
    if args.output:
        np.savez(args.output, **result)
    return result


if __name__ == "__main__":
    main()
