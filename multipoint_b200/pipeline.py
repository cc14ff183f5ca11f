"""Sync-free extract-and-match chain: images -> backbone (cuDNN) ->
detector-head kernel -> NMS + top-k + ordered keypoints -> channels-last descriptor normalise ->
descriptor sampling -> mutual nearest-neighbour matching, all on one stream with fixed-capacity
keypoint buffers and device-side counts, so nothing synchronises with the host until the caller
reads the result.  It is the batched form of what predict_align_image_pair.py:121-190 does per
sample; each stage is the same kernel the drop-in functions in multipoint_b200.utils call.
"""
import torch

from . import ops


class KeypointPipeline:
    def __init__(self, net, nms=4, detection_threshold=0.015, topk=2048, iou=0.1, metric='l2', cross_check=True,
                 match_threshold=-1.0, algo=None, trust_spectrum_keys=True, dense_nms_map=True):
        """trust_spectrum_keys: data['optical'] really holds optical images and data['thermal'] thermal ones (what
        ImagePairDataset yields), so the encoder routing of MultiPoint.py:107-122 needs no device->host read of
        ``is_optical``.  Set it to False to route every row by its ``is_optical`` flag like the reference."""
        if topk <= 0:
            raise ValueError("KeypointPipeline needs topk > 0 (fixed-capacity keypoint buffers)")
        self.net, self.nms, self.thr, self.topk, self.iou = net, nms, detection_threshold, int(topk), iou
        self.metric, self.cross_check, self.match_threshold, self.algo = metric, cross_check, match_threshold, algo
        self.trust_spectrum_keys = bool(trust_spectrum_keys)
        # dense_nms_map=False: 'prob_nms' (the dense map utils.box_nms returns) is not materialised -- keypoints, scores
        # and counts are the product of this chain, and the sparse top-k NMS then writes nothing of the map's size
        self.dense_nms_map = bool(dense_nms_map)

    @torch.no_grad()
    def extract_from_backbone(self, logits, raw_desc, H, W, valid_mask=None):
        """Hot path proper: backbone outputs -> keypoints + descriptors (no host sync).
        One stream: running the HBM-bound descriptor normalise on a second stream next to the
        (instruction-bound) NMS kernel was measured at 2.02-2.15 ms per step against 1.90 ms in
        sequence -- the NMS loses more from sharing its SMs than the overlap wins."""
        _, desc_nhwc = ops.normalize_descriptors(raw_desc, nchw=False, nhwc=True)
        prob = ops.detector_head(logits, valid_mask)
        B = prob.shape[0]
        dense, kp, scores, counts = ops.box_nms(prob.reshape(B, H, W), self.nms, self.thr, iou=self.iou,
                                                keep_top_k=self.topk, want_keypoints=True, kp_cap=self.topk,
                                                want_dense=self.dense_nms_map)
        split = None
        if raw_desc.shape[1] in (64, 128, 256):
            # the sampler also writes the rows as the tensor-core matcher wants them (bf16 planes + squared norms):
            # match() then needs no pass of its own over the descriptors
            desc, split = ops.sample_descriptors(kp, desc_nhwc, H, W, counts=counts, channels_last=True, split=True)
        else:
            desc = ops.sample_descriptors(kp, desc_nhwc, H, W, counts=counts, channels_last=True)
        out = {'prob': prob, 'keypoints': kp, 'scores': scores, 'counts': counts, 'desc': desc}
        if split is not None:
            out.update(desc_hi=split['hi'], desc_mid=split['mid'], desc_sq_norms=split['sq_norms'], desc_max_norm=split['max_norm'])
        if dense is not None:
            out['prob_nms'] = dense.reshape(B, 1, H, W)
        return out

    @torch.no_grad()
    def extract(self, data):
        H, W = data['image'].shape[-2:]
        logits, raw = self.net.backbone_outputs(data)
        return self.extract_from_backbone(logits, raw, H, W, data.get('valid_mask'))

    @torch.no_grad()
    def match(self, ext_a, ext_b):
        def split_of(e):
            if 'desc_hi' not in e:
                return None
            return {'hi': e['desc_hi'], 'mid': e['desc_mid'], 'sq_norms': e['desc_sq_norms'], 'max_norm': e['desc_max_norm']}
        q, t, d, c = ops.match(ext_a['desc'], ext_b['desc'], metric=self.metric, algo=self.algo, kind='mutual',
                               cross_check=self.cross_check, threshold=self.match_threshold,
                               n1=ext_a['counts'], n2=ext_b['counts'], split1=split_of(ext_a), split2=split_of(ext_b))
        return {'query': q, 'train': t, 'distance': d, 'counts': c}

    def stream(self, host_batches, device):
        """Double-buffered form of __call__ for batches that live in pinned host memory (see stream_pairs)."""
        return stream_pairs(self, host_batches, device)

    @torch.no_grad()
    def __call__(self, data):
        """data = {'optical': {...}, 'thermal': {...}} with (B,1,H,W) images: one batched backbone
        pass over both spectra (the reference notes this at predict_align_image_pair.py:122)."""
        o, t = data['optical'], data['thermal']
        B = o['image'].shape[0]
        both = {'image': torch.cat([o['image'], t['image']])}
        if 'is_optical' in o and 'is_optical' in t:
            both['is_optical'] = torch.cat([o['is_optical'], t['is_optical']])
            if self.trust_spectrum_keys:
                both['n_optical'] = B   # rows [0,B) optical, [B,2B) thermal: the encoders are picked without a host sync
        if 'valid_mask' in o and 'valid_mask' in t:
            both['valid_mask'] = torch.cat([o['valid_mask'], t['valid_mask']])
        ext = self.extract(both)
        ea = {k: v[:B] for k, v in ext.items()}
        eb = {k: v[B:] for k, v in ext.items()}
        return {'optical': ea, 'thermal': eb, 'matches': self.match(ea, eb)}


def stream_pairs(pipe, host_batches, device):
    """Run ``pipe`` over an iterable of host batches ({'optical': {...}, 'thermal': {...}} of pinned tensors, all of
    one shape), yielding one result per batch.  The host->device copy of batch i+1 runs on a copy stream into the
    second of two preallocated device buffers while batch i is being processed, so the PCIe transfer (3 ms for
    64 pairs of 512x640) hides behind the step.  No allocation happens per batch (a fresh 168 MB allocation on
    the copy stream every step costs more than the copy)."""
    device = torch.device(device)
    cur = torch.cuda.current_stream(device)
    copy_stream = torch.cuda.Stream(device=device)
    bufs, uploaded, consumed = [None, None], [None, None], [None, None]

    def upload(batch, slot):
        if bufs[slot] is None:
            bufs[slot] = {s: {k: torch.empty(v.shape, dtype=v.dtype, device=device) for k, v in d.items()} for s, d in batch.items()}
        if consumed[slot] is not None:
            copy_stream.wait_event(consumed[slot])     # the step that read this buffer two batches ago is done
        with torch.cuda.stream(copy_stream):
            for s, d in batch.items():
                for k, v in d.items():
                    bufs[slot][s][k].copy_(v, non_blocking=True)
            uploaded[slot] = torch.cuda.Event()
            uploaded[slot].record(copy_stream)

    it = iter(host_batches)
    batch = next(it, None)
    if batch is None:
        return
    copy_stream.wait_stream(cur)                       # buffers allocated on the compute stream above are ready
    upload(batch, 0)
    i = 0
    while batch is not None:
        slot = i & 1
        batch = next(it, None)
        if batch is not None:
            upload(batch, slot ^ 1)                    # prefetch the next batch
        cur.wait_event(uploaded[slot])
        result = pipe(bufs[slot])
        consumed[slot] = torch.cuda.Event()
        consumed[slot].record(cur)
        yield result
        i += 1


def calibrate_random_init(net, images, sigma=2.0, dustbin_bias=5.0, is_optical=None):
    """Make a randomly initialised MultiPoint produce non-degenerate outputs for benchmarks
    (with random weights the activations collapse to ~1e-5 variance, the
    heatmap is flat ~1/65 > threshold everywhere and all descriptors coincide).  One forward pass in
    training mode with momentum 1 sets every BatchNorm's running statistics to the statistics of
    its own input on ``images`` (data-dependent init); the detector's final BatchNorm affine is then
    set to (sigma, +dustbin_bias on channel 64) so logits ~ sigma*N(0,1) with a dominant dustbin,
    like multipoint_b200.synthetic.logits.  Convolution weights stay random-init."""
    bns = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    old = [(m.momentum, m.training) for m in bns]
    was_training = net.training
    net.train()
    for m in bns:
        m.momentum = 1.0
    with torch.no_grad():
        data = {'image': images}
        if is_optical is not None:
            data['is_optical'] = is_optical
        net.backbone_outputs(data)
        det = net.detector_head_convolutions[-1]
        assert isinstance(det, torch.nn.BatchNorm2d), "needs final_batchnorm"
        det.weight.fill_(sigma)
        det.bias.zero_()
        det.bias[64] = dustbin_bias
    for m, (mom, _) in zip(bns, old):
        m.momentum = mom
    net.train(was_training)
    return net
