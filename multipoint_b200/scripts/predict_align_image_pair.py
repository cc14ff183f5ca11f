"""Entry point mirroring predict_align_image_pair.py (reference :19-143): detect, describe and
match one optical/thermal pair and print the reference's three phase timings.  The plotting
branch (:146-254) is out of scope; RANSAC alignment of the matches (:216) is reported as a number.

    python -m multipoint_b200.scripts.predict_align_image_pair -m model_weights/multipoint -v none -i 0
"""
import argparse
import random
import time

import numpy as np
import torch

from .. import utils
from .common import build_network, load_config, load_samples, require_cuda


def main(argv=None):
    parser = argparse.ArgumentParser(description='Predict the keypoints of an image pair and match them')
    parser.add_argument('-y', '--yaml-config', default='configs/config_image_pair_dataset_prediction.yaml', help='YAML config file')
    parser.add_argument('-m', '--model-dir', default='model_weights/multipoint', help='Directory of the model')
    parser.add_argument('-v', '--version', default='latest', help='Model version (name of the param file), none for no weights')
    parser.add_argument('-i', '--index', default=0, type=int, help='Index of the sample to predict')
    parser.add_argument('-s', '--seed', default=0, type=int, help='Seed of the random generators')
    parser.add_argument('--input', default=None, help='npz with optical/thermal arrays (default: synthetic pair)')
    parser.add_argument('-o', '--output', default=None, help='npz to write keypoints / matches to')
    args = parser.parse_args(argv)

    random.seed(args.seed); np.random.seed(args.seed); torch.manual_seed(args.seed)
    config = load_config(args.yaml_config, args.model_dir)
    device = require_cuda(config)
    print('Predicting on device: {}'.format(device))
    net = build_network(config, args.model_dir, args.version, device)
    pred = config['prediction']

    with torch.no_grad():
        torch.cuda.synchronize(); t_start = time.time()
        data = load_samples(args.input, args.index + 1, args.seed)
        data = {s: {k: v[args.index:args.index + 1].to(device) for k, v in data[s].items()} for s in ('optical', 'thermal')}
        torch.cuda.synchronize(); t_1 = time.time()
        out_optical = net(data['optical'])
        out_thermal = net(data['thermal'])
        torch.cuda.synchronize(); t_2 = time.time()
        if pred['nms'] > 0:
            for out, d in ((out_optical, data['optical']), (out_thermal, data['thermal'])):
                out['prob'] = utils.box_nms(out['prob'] * d['valid_mask'], pred['nms'], pred['detection_threshold'],
                                            keep_top_k=pred['topk'], on_cpu=pred['cpu_nms'])
        torch.cuda.synchronize(); t_3 = time.time()
        print('Loading the data took: {} s'.format(t_1 - t_start))
        print('Two forward passes took: {} s'.format(t_2 - t_1))
        print('Box nms: {} s'.format(t_3 - t_2))

        H, W = data['optical']['image'].shape[2:]
        kp_o = utils.extract_keypoints(out_optical['prob'][0], pred['detection_threshold'])
        kp_t = utils.extract_keypoints(out_thermal['prob'][0], pred['detection_threshold'])
        d_o = utils.interpolate_descriptors(kp_o, out_optical['desc'][0], H, W)
        d_t = utils.interpolate_descriptors(kp_t, out_thermal['desc'][0], H, W)
        m = pred['matching']
        q, t, dist = utils.match_descriptors(d_o, d_t, m['method'], m['knn_matches'], **m['method_kwargs']) \
            if len(kp_o) and len(kp_t) else (torch.zeros(0), torch.zeros(0), torch.zeros(0))
        torch.cuda.synchronize(); t_4 = time.time()
        print('Keypoints: {} optical, {} thermal; matches: {}; sampling + matching: {} s'.format(len(kp_o), len(kp_t), len(q), t_4 - t_3))
    result = {'keypoints_optical': kp_o.cpu().numpy(), 'keypoints_thermal': kp_t.cpu().numpy(), 'query': q.cpu().numpy(),
              'train': t.cpu().numpy(), 'distance': dist.cpu().numpy()}
    if len(q) >= 4:
        import cv2
        src = result['keypoints_optical'][result['query'].astype(int)][:, ::-1].astype(np.float32).reshape(-1, 1, 2)
        dst = result['keypoints_thermal'][result['train'].astype(int)][:, ::-1].astype(np.float32).reshape(-1, 1, 2)
        H_est, mask = cv2.findHomography(src, dst, cv2.RANSAC, ransacReprojThreshold=pred['reprojection_threshold'])
        result['H_est'] = H_est if H_est is not None else np.eye(3)
        print('Estimated Homography:'); print(result['H_est'])
    if args.output:
        np.savez(args.output, **result)
    return result


if __name__ == "__main__":
    main()
