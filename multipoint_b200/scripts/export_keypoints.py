"""Entry point mirroring export_keypoints.py (reference :12-106): homographic adaptation label
export.  For every batch: ``homographic_adaptation[_multispectral]`` -> box NMS (batched, or per
sample with -snms) -> ``torch.nonzero`` keypoints, stored per sample name.  The reference writes an
h5 group per name with dataset ``keypoints`` (int64 (K,2)); h5py is not installed, so the same
mapping is written as ``<name>/keypoints`` entries of an npz.  ``-skip`` resumes like :65-71.
With torchrun (one process per GPU) batches are sharded round-robin over the ranks, and with
``--shard-homographies`` the homography samples of each batch are split instead.

    python -m multipoint_b200.scripts.export_keypoints -o labels.npz -m model_weights/multipoint -v none --count 4
"""
import argparse
import os

import numpy as np
import torch

from .. import parallel, utils
from .common import build_network, load_config, load_samples, require_cuda


def main(argv=None):
    parser = argparse.ArgumentParser(description='Export the keypoints for images in a dataset using a base detector')
    parser.add_argument('-y', '--yaml-config', default='configs/config_export_keypoints.yaml', help='YAML config file')
    parser.add_argument('-o', '--output_file', required=True, help='Output file name (.npz)')
    parser.add_argument('-m', '--model-dir', default='model_weights/surf', help='Directory of the model')
    parser.add_argument('-v', '--version', default='none', help='Model version (name of the .model file)')
    parser.add_argument('-snms', '--single-nms', action='store_true', help='Do the nms calculation for each sample separately')
    parser.add_argument('-skip', dest='skip_processed', action='store_true', help='Skip already processed samples')
    parser.add_argument('--input', default=None, help='npz with optical/thermal arrays (default: synthetic)')
    parser.add_argument('--count', default=2, type=int, help='number of synthetic samples')
    parser.add_argument('--single-image', action='store_true', help='dataset returns single images, not pairs')
    parser.add_argument('--shard-homographies', action='store_true', help='multi-GPU: split homography samples, not batches')
    parser.add_argument('-s', '--seed', default=0, type=int)
    args = parser.parse_args(argv)

    config = load_config(args.yaml_config, args.model_dir)
    pred = config['prediction']
    ha_cfg = pred.get('homographic_adaptation', {})
    rank, local_rank, world = parallel.init_distributed()
    require_cuda(config)
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    net = build_network(config, args.model_dir, args.version, device)
    data = load_samples(args.input, args.count, args.seed)
    names = data['name']
    done = dict(np.load(args.output_file)) if (args.skip_processed and os.path.exists(args.output_file)) else {}
    bs = pred['batchsize']
    results = {}
    batches = list(range(0, len(names), bs))
    with torch.no_grad():
        for bi, start in enumerate(batches):
            batch_names = names[start:start + bs]
            if args.skip_processed and all((n + '/keypoints') in done for n in batch_names):
                continue
            shard = None
            if world > 1 and not args.shard_homographies and bi % world != rank:
                continue                                             # batches are independent units
            # the homography stream of a batch depends only on (seed, batch index): the labels are the same for any
            # world size, sharding mode and -skip state (the reference draws from an unseeded global stream)
            np.random.seed((args.seed * 1000003 + bi) % (2 ** 32))
            batch = {s: {k: v[start:start + bs].to(device) for k, v in data[s].items()} for s in ('optical', 'thermal')}
            Hs = masks = None
            if world > 1 and args.shard_homographies:
                cfg = utils._check_ha_config(ha_cfg)
                hw = tuple(batch['optical']['image'].shape[2:])
                Hs, masks = parallel.broadcast_homographies(lambda: utils.sample_adaptation_homographies(hw, cfg, with_masks=False), device=device)
                shard = parallel.adaptation_shard()
            if args.single_image:
                prob_ha = utils.homographic_adaptation(batch['optical'], net, ha_cfg, homographies=Hs, masks=masks, shard=shard)
            else:
                prob_ha = utils.homographic_adaptation_multispectral(batch, net, ha_cfg, homographies=Hs, masks=masks, shard=shard)
            if pred['nms'] > 0:
                if args.single_nms:
                    for i in range(prob_ha.shape[0]):
                        prob_ha[i, 0] = utils.box_nms(prob_ha[i, 0], pred['nms'], pred['detection_threshold'],
                                                      keep_top_k=pred['topk'], on_cpu=pred['cpu_nms'])
                else:
                    prob_ha = utils.box_nms(prob_ha, pred['nms'], pred['detection_threshold'], keep_top_k=pred['topk'],
                                            on_cpu=pred['cpu_nms'])
            for name, prob in zip(batch_names, prob_ha.split(1)):
                if not (args.skip_processed and (name + '/keypoints') in done):
                    results[name + '/keypoints'] = utils.extract_keypoints(prob, pred['detection_threshold']).cpu().numpy()
    if world > 1:
        gathered = [None] * world
        torch.distributed.all_gather_object(gathered, results)
        results = {k: v for part in gathered for k, v in part.items()}
    if rank == 0:
        done.update(results)
        np.savez(args.output_file, **done)
        print('wrote {} samples to {}'.format(len(done), args.output_file))
    return results


if __name__ == "__main__":
    main()
