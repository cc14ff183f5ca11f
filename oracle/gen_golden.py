"""Generate tests/golden/*.npz by running the REFERENCE's own functions (build container only).

The reference ships no tests, golden vectors or fixtures, so parity is
pinned by freezing the outputs of its own Python functions -- imported unmodified and read-only
from /root/reference through oracle/ref_shim.py -- on seeded inputs.  Run from the repo root:

    python oracle/gen_golden.py

Inputs are either stored (small) or regenerated from ``multipoint_b200.synthetic`` with the seed
and a sha1 checksum stored (large).  Library versions at generation time are recorded in
tests/golden/MANIFEST.json.  kornia is not installed, so ``homographic_adaptation`` itself cannot
run; its golden is produced by a torch restatement of the kornia calls (F.grid_sample,
align_corners=True) around the reference's own host functions -- PARITY UNPINNED for that row.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from ref_shim import import_reference  # noqa: E402

models, utils = import_reference()

import cv2  # noqa: E402
import torch  # noqa: E402
import torchvision  # noqa: E402

from multipoint_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(max(1, os.cpu_count() or 1))


def save(name, **arrays):
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in arrays.items()})


def sparse(dense):
    """dense (B,1,H,W) or (H,W) -> flat indices + values of the non-zeros (row-major)."""
    flat = np.ascontiguousarray(dense).reshape(-1)
    idx = np.flatnonzero(flat)
    return idx.astype(np.int64), flat[idx].astype(np.float32)


# ---------------------------------------------------------------- row 1/2: model tails
def gen_heads():
    net = models.MultiPoint({'multispectral': False, 'descriptor_size': 64})
    net.eval()
    net.detector_head_convolutions = torch.nn.Identity()
    net.descriptor_head_convolutions = torch.nn.Identity()
    lg = syn.logits(11, 2, 8, 10)
    lg_full = syn.logits(12, 1, 64, 80, sigma=3.0, bias=6.0)
    with torch.no_grad():
        prob, none_logits = net.detector_head(torch.from_numpy(lg))
        assert none_logits is None
        prob_full, _ = net.detector_head(torch.from_numpy(lg_full))
        net.set_force_return_logits(True)
        p2, l2 = net.detector_head(torch.from_numpy(lg))
        assert p2 is None and torch.equal(l2, torch.from_numpy(lg))
        dm = syn.descriptor_map(13, 2, 64, 8, 10)
        dm[0, :, 0, 0] = 0.0  # zero-norm cell: clamp_min(eps) path
        desc = net.descriptor_head(torch.from_numpy(dm))
        dm256 = syn.descriptor_map(14, 1, 256, 8, 10)
        desc256 = net.descriptor_head(torch.from_numpy(dm256))
        d2s = utils.depth_to_space(torch.from_numpy(lg[:, :64].copy()), 8)
    # full-size heatmap: keep a strided sample to stay small, plus its float64 sum
    pf = prob_full.numpy()
    save("heads", logits=lg, prob=prob.numpy(), desc_in=dm, desc=desc.numpy(), desc256_in=dm256,
         desc256=desc256.numpy(), depth_to_space=d2s.numpy(),
         full_seed=np.array([12, 1, 64, 80]), full_sigma_bias=np.array([3.0, 6.0]),
         full_checksum=np.array(syn.checksum(lg_full)), full_rows=pf[0, 0, ::37].copy(),
         full_sum=np.array(pf.astype(np.float64).sum()))


# ---------------------------------------------------------------- MagicLeap heatmap
def gen_magicleap():
    semi = syn.logits(21, 2, 8, 10, sigma=3.0, bias=6.0)
    semi[0, 5, 0, 0] = 40.0   # no max subtraction in the reference: exp(40) must stay finite in fp32
    semi[1, :, 3, 3] = -30.0  # nearly empty cell: the +1e-5 in the denominator dominates
    net = models.SuperPointMagicLeap()
    prob = net.generate_heatmap(torch.from_numpy(semi), (2, 1, 64, 80))
    # whole forward on a seeded random-init net: semi/desc/prob of the reference model
    torch.manual_seed(7)
    net = models.SuperPointMagicLeap().eval()
    img = syn.images(22, 1, 64, 80)
    with torch.no_grad():
        o = net({'image': torch.from_numpy(img)})
    save("magicleap", semi=semi, prob=prob.numpy(), model_image=img, model_seed=np.array(7),
         model_logits=o['logits'].numpy(), model_desc=o['desc'].numpy(), model_prob=o['prob'].numpy())


# ---------------------------------------------------------------- evaluation loops
def evaluation_batches():
    """Two batches of two 128x160 pairs with canned network outputs (the net is a lookup, so the reference on
    the CPU and the product on the GPU see identical heatmaps / descriptor maps).  Batch 0: no homography key
    (identity path), thermal = optical shifted by a pixel here and there + descriptor noise, so many matches
    are correct.  Batch 1: sampled homographies on both spectra and a partially masked frame."""
    H, W, D = 128, 160, 64
    batches = []
    rng = np.random.default_rng(77)
    for bi in range(2):
        po = syn.heatmap(300 + bi, 2, H, W, squarings=5, scale=0.5)
        do = syn.descriptor_map(310 + bi, 2, D, H // 8, W // 8)
        do /= np.linalg.norm(do, axis=1, keepdims=True)
        if bi == 0:
            pt = po.copy()
            pt[:, :, :, 80:] = np.roll(po[:, :, :, 80:], 1, axis=2)        # right half moved down one row
            dt = do + 0.05 * rng.standard_normal(do.shape).astype(np.float32)
        else:
            pt = syn.heatmap(320 + bi, 2, H, W, squarings=5, scale=0.5)
            dt = syn.descriptor_map(330 + bi, 2, D, H // 8, W // 8)
        dt /= np.linalg.norm(dt, axis=1, keepdims=True)
        mo = np.ones((2, 1, H, W), np.float32)
        mt = np.ones((2, 1, H, W), np.float32)
        if bi == 1:
            mt[:, :, :10] = 0
            mo[1, :, :, -7:] = 0
        b = {'optical': {'image': np.zeros((2, 1, H, W), np.float32), 'valid_mask': mo, 'stub_prob': po, 'stub_desc': do.astype(np.float32)},
             'thermal': {'image': np.zeros((2, 1, H, W), np.float32), 'valid_mask': mt, 'stub_prob': pt, 'stub_desc': dt.astype(np.float32)}}
        if bi == 1:
            np.random.seed(40)
            cfg = dict(translation=True, rotation=True, scaling=True, perspective=True, scaling_amplitude=0.1,
                       perspective_amplitude_x=0.1, perspective_amplitude_y=0.1, patch_ratio=0.9, max_angle=0.3,
                       allow_artifacts=True)
            b['optical']['homography'] = np.stack([utils.sample_homography(np.array([H, W]), **cfg) for _ in range(2)]).astype(np.float32)
            b['thermal']['homography'] = np.stack([utils.sample_homography(np.array([H, W]), **cfg) for _ in range(2)]).astype(np.float32)
        batches.append(b)
    return batches


def gen_evaluation():
    from multipoint.utils import evaluation as ref_eval
    batches = evaluation_batches()

    def loader():
        return [{s: {k: torch.from_numpy(v.copy()) for k, v in b[s].items()} for s in b} for b in batches]

    def net(d):
        return {'prob': d['stub_prob'].clone(), 'desc': d['stub_desc'].clone()}

    out = {}
    for bi, b in enumerate(batches):
        for s in b:
            for k, v in b[s].items():
                if k != 'image':
                    out["in%d_%s_%s" % (bi, s, k)] = v
    for tag, topk in (("rep_top0", 0), ("rep_top150", 150)):
        cfg = {'prediction': {'detection_threshold': 0.015, 'nms': 4, 'topk': topk, 'cpu_nms': True}}
        mean, rep, nko, nkt = ref_eval.compute_repeatability_multispectral(net, loader(), 'cpu', cfg, distance_thresh=3)
        out[tag] = np.array([mean] + list(rep))
        out[tag + "_nkp"] = np.array([nko, nkt])
    for tag, method, kwargs, topk in (("desc_bf", "bfmatcher", {'crossCheck': True}, 0), ("desc_nn", "nnmatcher", {'threshold': 0.9}, 200)):
        cfg = {'detection_threshold': 0.015, 'nms': 4, 'topk': topk, 'cpu_nms': True, 'reprojection_threshold': 3,
               'matching': {'method': method, 'knn_matches': False, 'method_kwargs': kwargs}}
        res = ref_eval.compute_descriptor_metrics(net, loader(), 'cpu', cfg, threshold_keypoints=4, threshold_warp=4)
        for k, v in res.items():
            out[tag + "_" + k] = np.asarray(v)
    # the point geometry on its own: reference expressions evaluated literally (evaluation.py:166-197, 288-302)
    np.random.seed(41)
    pts = {}
    for i in range(6):
        Hm = utils.sample_homography(np.array([128, 160]))
        if i % 2:
            Hm = Hm.astype(np.float32)                      # the loops pass fp32 matrices from torch
        a = syn.keypoints(500 + i, 180, 128, 160)
        b = syn.keypoints(520 + i, 150, 128, 160)
        wi = utils.warp_keypoints(a, Hm)
        wf = utils.warp_keypoints(a.astype(np.float32), Hm, np.float64)
        wi_f = utils.filter_points(wi, (128, 160))
        dist = np.linalg.norm(np.expand_dims(wi_f, 1) - np.expand_dims(b, 0), ord=None, axis=2)
        d = torch.from_numpy(wf).unsqueeze(1) - torch.from_numpy(b).unsqueeze(0)
        correct = torch.norm(d.float(), dim=-1) <= 4
        pts.update({"pt%d_H" % i: Hm, "pt%d_a" % i: a, "pt%d_b" % i: b, "pt%d_warp_int" % i: wi, "pt%d_warp_f64" % i: wf,
                    "pt%d_min_dist" % i: np.min(dist, axis=1), "pt%d_filtered" % i: wi_f,
                    "pt%d_correct_rows" % i: correct.sum(1).nonzero()[:, 0].numpy(),
                    "pt%d_correct_diag" % i: correct[torch.arange(150), torch.arange(150)].numpy()})
    out.update(pts)
    save("evaluation", **out)


# ---------------------------------------------------------------- row 4: box_nms
def gen_nms():
    out = {}
    cases = []

    def run(tag, prob, size, thr, topk, **kw):
        res = utils.box_nms(torch.from_numpy(prob), size, thr, keep_top_k=topk, on_cpu=True, **kw).numpy()
        idx, val = sparse(res)
        out[tag + "_idx"] = idx
        out[tag + "_val"] = val
        cases.append(tag)
        return res

    # hand cases measured in the survey
    H, W = 24, 32
    p = np.zeros((H, W), np.float32)
    p[10, 10], p[10, 13], p[10, 16] = 0.9, 0.8, 0.7  # chain: A kills B, C survives
    out["chain_in"] = p
    run("chain", p, 4, 0.015, 0)
    p = np.zeros((H, W), np.float32)
    p[9, 11] = p[9, 12] = p[10, 11] = p[10, 12] = 0.5  # 4-way tie -> lowest row-major index
    out["tie4_in"] = p
    run("tie4", p, 4, 0.015, 0)
    p = np.zeros((H, W), np.float32)
    p[5, 5] = np.float32(0.015)  # strict '>' : equal to threshold is rejected
    p[5, 20] = np.nextafter(np.float32(0.015), np.float32(1))
    out["strict_in"] = p
    run("strict", p, 4, 0.015, 0)
    p = np.zeros((H, W), np.float32)
    p[3, 3], p[12, 20], p[20, 8] = 0.3, 0.9, 0.6  # top-2 of 3
    out["top2_in"] = p
    run("top2", p, 4, 0.015, 2)
    # footprint probes for sizes 3, 4, 8 and a non-default iou: one strong centre + one weaker
    # point at every offset, each pair in its own image (4-D call -> per-image NMS)
    for size, iou in [(3, 0.1), (4, 0.1), (8, 0.1), (4, 0.3), (5, 0.1), (2.5, 0.1)]:
        R = int(np.ceil(size))
        S = 2 * R + 1
        fp = np.zeros((S, S), np.uint8)
        batch = np.zeros((S * S, 1, 4 * R + 1, 4 * R + 1), np.float32)
        for k in range(S * S):
            dy, dx = k // S - R, k % S - R
            batch[k, 0, 2 * R, 2 * R] = 0.9
            if (dy, dx) != (0, 0):
                batch[k, 0, 2 * R + dy, 2 * R + dx] = 0.5
        res = utils.box_nms(torch.from_numpy(batch), size, 0.015, iou=iou, on_cpu=True).numpy()
        for k in range(S * S):
            dy, dx = k // S - R, k % S - R
            if (dy, dx) != (0, 0):
                fp[dy + R, dx + R] = res[k, 0, 2 * R + dy, 2 * R + dx] == 0
        out["footprint_s%s_i%s" % (size, iou)] = fp
    # small random maps, 2-D and 4-D, several sizes / top-k, incl. quantised (exact ties)
    small = []
    for seed, size, topk, quant, B in [(21, 4, 0, None, 1), (22, 4, 50, None, 3), (23, 3, 0, 64, 2),
                                       (24, 8, 20, None, 2), (25, 4, 0, 32, 3), (26, 5, 7, 128, 2),
                                       (27, 2.5, 0, None, 1), (28, 4, 30, 64, 4)]:
        hm = syn.heatmap(seed, B, 64, 80, quant=quant)
        tag = "small%d" % seed
        run(tag + "_4d", hm, size, 0.015, topk)
        run(tag + "_2d", hm[0, 0], size, 0.015, topk)
        small.append((seed, size, topk, quant or 0, B))
    out["small_cases"] = np.array(small, np.float64)
    # >1000 candidates per call in 4-D mode takes torchvision's _batched_nms_vanilla branch
    # full size 512x640 (config 2 sizes): dense + top-k 2048, plus a tie-stress map
    full = []
    for seed, topk, quant, B in [(31, 0, None, 1), (32, 2048, None, 2), (33, 2048, 4096, 1)]:
        hm = syn.heatmap(seed, B, 512, 640, quant=quant)
        tag = "full%d" % seed
        run(tag + "_4d", hm, 4, 0.015, topk)
        out[tag + "_checksum"] = np.array(syn.checksum(hm))
        full.append((seed, topk, quant or 0, B))
    out["full_cases"] = np.array(full, np.float64)
    # heatmap produced by the reference's own softmax path (sigma=2, bias=5), full size
    net = models.MultiPoint({'multispectral': False, 'descriptor_size': 64})
    net.eval()
    net.detector_head_convolutions = torch.nn.Identity()
    lg = syn.logits(34, 1, 64, 80)
    with torch.no_grad():
        prob = net.detector_head(torch.from_numpy(lg))[0].numpy()
    out["softmax34_prob_f16hash"] = np.array(syn.checksum(prob))
    res = run("softmax34", prob, 4, 0.015, 2048)
    # keypoint idiom: torch.nonzero((p > thr).float())
    kp = torch.nonzero((torch.from_numpy(res).squeeze() > 0.015).float()).numpy()
    out["softmax34_kp"] = kp.astype(np.int64)
    out["softmax34_prob"] = prob.astype(np.float32)  # 1.3 MB: the one stored full-size input
    save("box_nms", **out)


# ---------------------------------------------------------------- row 5: interpolate_descriptors
def gen_interp():
    out = {}
    for D, seed in [(64, 41), (256, 42)]:
        dm = syn.descriptor_map(seed, 1, D, 64, 80)[0]
        kp = syn.keypoints(seed, 96, 512, 640)
        r = utils.interpolate_descriptors(torch.from_numpy(kp), torch.from_numpy(dm), 512, 640).numpy()
        out["d%d_seed" % D] = np.array([seed])
        out["d%d_checksum" % D] = np.array(syn.checksum(dm))
        out["d%d_kp" % D] = kp
        out["d%d_out" % D] = r
    # small map stored in full, odd sizes
    dm = syn.descriptor_map(43, 1, 32, 5, 7)[0]
    kp = syn.keypoints(43, 40, 40, 56)
    out["small_in"] = dm
    out["small_kp"] = kp
    out["small_out"] = utils.interpolate_descriptors(torch.from_numpy(kp), torch.from_numpy(dm), 40, 56).numpy()
    empty = utils.interpolate_descriptors(torch.zeros((0, 2), dtype=torch.int64), torch.from_numpy(dm), 40, 56)
    out["empty_shape"] = np.array(empty.shape)
    save("interpolate", **out)


# ---------------------------------------------------------------- rows 6-8: matching
def dm_arrays(matches):
    return (np.array([m.queryIdx for m in matches], np.int32), np.array([m.trainIdx for m in matches], np.int32),
            np.array([m.distance for m in matches], np.float32))


def gen_matching():
    out = {}
    cases = []
    for seed, N1, N2, D, noise, dup in [(51, 300, 300, 64, 0.05, 0), (52, 257, 400, 256, 0.3, 0),
                                        (53, 500, 333, 128, 0.6, 0), (54, 200, 200, 64, 0.05, 12),
                                        (55, 1024, 1024, 256, 0.05, 0), (56, 64, 1, 64, 0.1, 0)]:
        a, b = syn.descriptor_sets(seed, N1, N2, D, noise, dup)
        tag = "m%d" % seed
        if a.nbytes + b.nbytes < 400_000:
            out[tag + "_a"], out[tag + "_b"] = a, b
        out[tag + "_checksum"] = np.array(syn.checksum(a) + syn.checksum(b))
        q, t, d = dm_arrays(utils.get_matches(a, b, 'bfmatcher', False, crossCheck=True))
        out[tag + "_bf_q"], out[tag + "_bf_t"], out[tag + "_bf_d"] = q, t, d
        q, t, d = dm_arrays(utils.get_matches(a, b, 'bfmatcher', False, crossCheck=False))
        out[tag + "_bfnc_q"], out[tag + "_bfnc_t"], out[tag + "_bfnc_d"] = q, t, d
        q, t, d = dm_arrays(utils.get_matches(a, b, 'nnmatcher', False))
        out[tag + "_nn_q"], out[tag + "_nn_t"], out[tag + "_nn_d"] = q, t, d
        q, t, d = dm_arrays(utils.get_matches(a, b, 'nnmatcher', False, threshold=1.1))
        out[tag + "_nn11_q"], out[tag + "_nn11_t"], out[tag + "_nn11_d"] = q, t, d
        if N2 >= 2:
            q, t, d = dm_arrays(utils.get_matches(a, b, 'bfmatcher', True))
            out[tag + "_knn_q"], out[tag + "_knn_t"], out[tag + "_knn_d"] = q, t, d
        if N1 * N2 <= 120_000:
            q, t, d = dm_arrays(utils.get_matches(a, b, 'thresholdmatcher', False, threshold=0.9))
            out[tag + "_thr_q"], out[tag + "_thr_t"], out[tag + "_thr_d"] = q, t, d
        cases.append((seed, N1, N2, D, noise, dup))
    out["cases"] = np.array(cases, np.float64)
    # float64 argmax of the similarity on the same data ("truth" triple of the survey)
    a, b = syn.descriptor_sets(55, 1024, 1024, 256, 0.05, 0)
    s = a.astype(np.float64) @ b.astype(np.float64).T
    out["m55_f64_row"] = s.argmax(1).astype(np.int32)
    out["m55_f64_col"] = s.argmax(0).astype(np.int32)
    # error behaviour
    errs = {}
    try:
        utils.get_matches(a, b, 'nope')
    except ValueError as e:
        errs['unknown'] = str(e)
    try:
        utils.NNMatcher(threshold=-1.0)
    except ValueError as e:
        errs['neg'] = str(e)
    out["err_unknown"] = np.array(errs['unknown'])
    out["err_neg"] = np.array(errs['neg'])
    out["empty_nn"] = np.array(len(utils.get_matches(np.zeros((0, 64), np.float32), b[:, :64].copy(), 'nnmatcher')))
    save("matching", **out)


# ---------------------------------------------------------------- row 11: host homography sampling
def gen_homographies():
    out = {}
    cfg_default = dict(utils.homographies.homography_adaptation_default_config['homographies'])
    cfg_export = dict(translation=True, rotation=True, scaling=True, perspective=True, scaling_amplitude=0.2,
                      perspective_amplitude_x=0.2, perspective_amplitude_y=0.2, patch_ratio=0.85,
                      max_angle=1.57, allow_artifacts=True)
    cfg_noart = dict(cfg_export, allow_artifacts=False, translation_overflow=0.05)
    for tag, cfg, shape, seed, n in [("default", cfg_default, (512, 640), 0, 6), ("export", cfg_export, (512, 640), 1, 6),
                                     ("noart", cfg_noart, (64, 80), 2, 6), ("small", cfg_export, (64, 80), 3, 8)]:
        np.random.seed(seed)
        Hs, masks = [], []
        for i in range(n):
            Hm = utils.sample_homography(np.array(shape), **cfg)
            Hs.append(Hm)
            masks.append(utils.compute_valid_mask(tuple(shape), Hm, 3 if tag != "default" else 5, True))
        out[tag + "_H"] = np.stack(Hs)
        out[tag + "_mask"] = np.packbits(np.stack(masks).astype(bool), axis=-1)
        out[tag + "_shape"] = np.array(shape)
        out[tag + "_seed"] = np.array([seed, 3 if tag != "default" else 5])
    # no-erosion / no-border variants on one matrix
    Hm = out["small_H"][0]
    out["small_mask_e0"] = np.packbits(utils.compute_valid_mask((64, 80), Hm, 0, False).astype(bool), axis=-1)
    out["small_mask_e2nb"] = np.packbits(utils.compute_valid_mask((64, 80), Hm, 2, False).astype(bool), axis=-1)
    # warp_keypoints / filter_points helpers
    kp = syn.keypoints(61, 50, 64, 80)
    out["wk_kp"] = kp
    out["wk_out"] = utils.warp_keypoints(kp, Hm)
    out["wk_filtered"] = utils.filter_points(out["wk_out"], (64, 80))
    save("homographies", **out)


# ---------------------------------------------------------------- rows 9-10: adaptation (restated)
def kornia_free_warp(src, M, dsize, mode='bilinear', padding_mode='zeros'):
    """What warp_perspective_tensor (homographies.py:404-425) does through kornia ~0.2-0.4, written
    with torch only: dst_norm_to_dst_norm + homography_warp(grid_sample, align_corners=True)."""
    B, C, H, W = src.shape
    N = torch.tensor([[2.0 / (W - 1), 0, -1], [0, 2.0 / (H - 1), -1], [0, 0, 1]], dtype=torch.float32)
    M_norm = N @ (M @ torch.inverse(N))
    A = torch.inverse(M_norm)
    ys, xs = torch.meshgrid(torch.linspace(-1, 1, dsize[0]), torch.linspace(-1, 1, dsize[1]), indexing='ij')
    grid = torch.stack([xs, ys, torch.ones_like(xs)], -1).reshape(1, -1, 3).expand(B, -1, -1)
    pts = grid @ A.transpose(1, 2)
    z = pts[..., 2:]
    scale = torch.where(z.abs() > 1e-8, 1.0 / z, torch.ones_like(z))
    flow = (pts[..., :2] * scale).reshape(B, dsize[0], dsize[1], 2)
    return torch.nn.functional.grid_sample(src, flow, mode=mode, padding_mode=padding_mode, align_corners=True)


def gen_adaptation():
    out = {}
    hom = utils.homographies
    hom.kornia_available = True
    hom.warp_perspective_tensor = kornia_free_warp  # WarpingModule.forward looks this up at call time

    conv = torch.nn.Conv2d(1, 65, 8, stride=8)
    torch.manual_seed(7)
    torch.nn.init.normal_(conv.weight, std=1.5)
    torch.nn.init.normal_(conv.bias, std=0.5)
    conv.bias.data[64] += 2.0

    def net(data):
        # stub network with the MultiPoint output contract: any dict -> {'prob': (B,1,H,W)}
        with torch.no_grad():
            lg = conv(data['image'])
            p = torch.nn.functional.pixel_shuffle(torch.softmax(lg, 1)[:, :-1], 8)
        return {'prob': p}

    out["stub_w"] = conv.weight.detach().numpy()
    out["stub_b"] = conv.bias.detach().numpy()
    H, W = 64, 80
    batch = syn.image_pair_batch(71, 2, H, W)
    img_o, img_t = torch.from_numpy(batch['optical']['image']), torch.from_numpy(batch['thermal']['image'])
    out["img_o"], out["img_t"] = batch['optical']['image'], batch['thermal']['image']
    base = dict(num=6, erosion_radius=3, mask_border=True, min_count=2, filter_size=0,
                homographies=dict(translation=True, rotation=True, scaling=True, perspective=True,
                                  scaling_amplitude=0.2, perspective_amplitude_x=0.2, perspective_amplitude_y=0.2,
                                  patch_ratio=0.85, max_angle=1.57, allow_artifacts=True))
    with torch.no_grad():
        np.random.seed(5)
        cfg = dict(base, aggregation='prod')
        out["single"] = hom.homographic_adaptation({'image': img_o.clone()}, net, cfg).numpy()
        for agg in ('prod', 'sum'):
            np.random.seed(5)
            cfg = dict(base, aggregation=agg)
            data = {'optical': {'image': img_o.clone(), 'is_optical': torch.ones(2, 1, dtype=torch.bool)},
                    'thermal': {'image': img_t.clone(), 'is_optical': torch.zeros(2, 1, dtype=torch.bool)}}
            out["multi_" + agg] = hom.homographic_adaptation_multispectral(data, net, cfg).numpy()
        # the Gaussian-filter branch (filter_size > 0, homographies.py:54-58,100-102,145-149,177-178) on the same stream
        np.random.seed(5)
        out["single_f5"] = hom.homographic_adaptation({'image': img_o.clone()}, net, dict(base, aggregation='prod', filter_size=5)).numpy()
        np.random.seed(5)
        data = {'optical': {'image': img_o.clone(), 'is_optical': torch.ones(2, 1, dtype=torch.bool)},
                'thermal': {'image': img_t.clone(), 'is_optical': torch.zeros(2, 1, dtype=torch.bool)}}
        out["multi_prod_f5"] = hom.homographic_adaptation_multispectral(data, net, dict(base, aggregation='prod', filter_size=5)).numpy()
        # the filter itself (utils.py:124-157): weights for three sizes / an explicit sigma, and its action on a heatmap
        for k, sg in ((3, None), (5, None), (7, 1.5)):
            f = utils.get_gaussian_filter(k) if sg is None else utils.get_gaussian_filter(k, sg)
            out["gauss_w_%d" % k] = f.weight.detach().numpy()
        f5 = utils.get_gaussian_filter(5)
        p_in = net({'image': img_o})['prob']
        out["gauss_in"] = p_in.numpy()
        out["gauss_out_5"] = f5(torch.nn.ReflectionPad2d(2)(p_in)).detach().numpy()
        # the homographies / masks the calls above consumed (same RNG stream)
        np.random.seed(5)
        Hs, masks = [], []
        for i in range(5):
            Hm = utils.sample_homography(np.array([H, W]), **base['homographies'])
            Hs.append(Hm)
            masks.append(utils.compute_valid_mask((H, W), Hm, 3, True))
        out["H"] = np.stack(Hs)
        out["masks"] = np.stack(masks).astype(np.float32)
        # the normalised 3x3 matrices exactly as the torch restatement rounds them, so the
        # per-pixel arithmetic can be pinned separately from the 3x3 algebra
        Nn = torch.tensor([[2.0 / (W - 1), 0, -1], [0, 2.0 / (H - 1), -1], [0, 0, 1]], dtype=torch.float32)
        Aw, Au = [], []
        for Hm in Hs:
            Mf = torch.from_numpy(Hm.astype(np.float32))
            Aw.append(torch.inverse(Nn @ (Mf @ torch.inverse(Nn))).numpy())
            Au.append(torch.inverse(Nn @ (torch.inverse(Mf) @ torch.inverse(Nn))).numpy())
        out["A_warp"], out["A_unwarp"] = np.stack(Aw), np.stack(Au)
        # plain warps for the kernel-level check
        M = torch.from_numpy(Hs[0].astype(np.float32))[None].repeat(2, 1, 1)
        out["warp_bilinear_reflection"] = kornia_free_warp(img_o, M, (H, W), 'bilinear', 'reflection').numpy()
        out["warp_bilinear_zeros"] = kornia_free_warp(img_o, torch.inverse(M), (H, W), 'bilinear', 'zeros').numpy()
        out["warp_nearest_zeros"] = kornia_free_warp(torch.from_numpy(out["masks"][:1])[None].repeat(2, 1, 1, 1),
                                                     torch.inverse(M), (H, W), 'nearest', 'zeros').numpy()
    out["seed"] = np.array([5])
    save("adaptation", **out)
    # error behaviour (homographies.py:42-46,68,123)
    errs = {}
    for key, cfg in [('num', dict(base, num=0)), ('filter', dict(base, filter_size=4))]:
        try:
            hom.homographic_adaptation({'image': img_o}, net, cfg)
        except ValueError as e:
            errs[key] = str(e)
    try:
        data = {'optical': {'image': img_o, 'is_optical': torch.ones(2, 1, dtype=torch.bool)},
                'thermal': {'image': img_t, 'is_optical': torch.zeros(2, 1, dtype=torch.bool)}}
        hom.homographic_adaptation_multispectral(data, net, dict(base, num=2, aggregation='max'))
    except ValueError as e:
        errs['agg'] = str(e)
    return errs


# ---------------------------------------------------------------- row 3: whole model contract
def gen_model():
    out = {}
    for tag, cfg in [("shipped", {'multispectral': False, 'descriptor_size': 64, 'bn_first': False,
                                  'descriptor_head': True, 'final_batchnorm': True, 'reflection_pad': True,
                                  'normalize_descriptors': True}),
                     ("multi", {'multispectral': True, 'descriptor_size': 256})]:
        torch.manual_seed(0)
        net = models.MultiPoint(cfg)
        net.eval()
        keys = list(net.state_dict().keys())
        shapes = [tuple(v.shape) for v in net.state_dict().values()]
        out[tag + "_keys"] = np.array(keys)
        out[tag + "_shapes"] = np.array([str(s) for s in shapes])
        out[tag + "_nparams"] = np.array(sum(p.numel() for p in net.parameters()))
        img = syn.images(81, 2, 64, 80)
        data = {'image': torch.from_numpy(img), 'is_optical': torch.tensor([[True], [False]])}
        with torch.no_grad():
            o = net(data)
        out[tag + "_prob"] = o['prob'].numpy()
        out[tag + "_desc"] = o['desc'].numpy()
        assert o['logits'] is None
        net.train()
    save("model", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["heads", "magicleap", "nms", "interp", "matching", "homographies", "adaptation", "model", "evaluation"]
    errs = {}
    for w in which:
        r = globals()["gen_" + w]()
        if isinstance(r, dict):
            errs.update(r)
    manifest = {"torch": torch.__version__, "torchvision": torchvision.__version__, "cv2": cv2.__version__,
                "numpy": np.__version__, "reference": "ethz-asl/multipoint @ /root/reference (read-only)",
                "kornia": "not installed -- adaptation goldens use the torch restatement in gen_golden.py",
                "adaptation_errors": errs}
    if not sys.argv[1:]:
        with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
            json.dump(manifest, f, indent=1)
