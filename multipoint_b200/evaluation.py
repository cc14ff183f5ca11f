"""Device-resident mirror of the two evaluation loops that call the hot path:

    compute_repeatability_multispectral   multipoint/utils/evaluation.py:105-200
    compute_descriptor_metrics            multipoint/utils/evaluation.py:209-438

Same signatures, same return values.  What changes is where the per-sample inner loops run: the
reference pulls every heatmap to the host, builds N1 x N2 distance / "correct match" matrices in
numpy / torch and calls the OpenCV matcher three times per pair; here keypoints, warps, nearest
distances, correctness flags, descriptor sampling and the matchers run batched on the GPU through
the C ABI (mp_extract_keypoints_f32, mp_warp_keypoints_i64, mp_points_min_dist2_i64,
mp_points_correct_f32, mp_sample_descriptors_f32, mp_match_f32) and one small device->host copy
per batch brings back counts, flags and matches.  RANSAC (cv2.findHomography) and the final
precision/recall bookkeeping stay on the host, as in the reference.

3x3 inverses are taken by torch on the CPU in fp32 -- what the reference does when it evaluates on
the CPU (on a GPU it would call cuSOLVER and round differently).
"""
import numpy as np
import torch

from . import ops
from .utils import box_nms, data_to_device, match_descriptors, warp_keypoints


def div0(a, b):
    """a / b with 0/0 -> 1 and x/0 -> 0 (evaluation.py:202-207)."""
    with np.errstate(divide='ignore', invalid='ignore'):
        c = np.true_divide(a, b)
        bad = ~np.isfinite(c)
        c[bad] = np.where(a[bad] == 0, 1, 0)
    return c


def compute_mAP(precision, recall):
    """Area under the precision/recall staircase (evaluation.py:98-103)."""
    return np.sum(precision[1:] * (recall[1:] - recall[:-1]))


def _host_homographies(d, key, n):
    if key in d:
        return d[key].detach().to('cpu', torch.float32)
    return torch.eye(3, 3).repeat(n, 1, 1)


def _keypoints(prob, threshold, mask=None):
    """(B,1,H,W) heatmaps -> (kp (B,cap,2) int64, counts (B) int32 device, counts host list), cap = max count."""
    kp, _, cnt = ops.extract_keypoints(prob[:, 0].contiguous(), threshold, None if mask is None else mask[:, 0])
    n = cnt.cpu().tolist()
    cap = max(1, max(n))
    return kp[:, :cap].contiguous(), cnt, n


def _live(counts, cap):
    return torch.arange(cap, device=counts.device)[None, :] < counts[:, None]


def compute_repeatability_multispectral(net, dataloader, device, config, distance_thresh=3, verbose=False):
    """Repeatability of the keypoints between the two spectra (evaluation.py:105-200)."""
    repeatability, n_kp_optical, n_kp_thermal = [], [], []
    pred = config['prediction']
    for data in dataloader:
        thr = pred['detection_threshold']
        B = data['optical']['image'].shape[0]
        h_o = _host_homographies(data['optical'], 'homography', B)
        h_t = _host_homographies(data['thermal'], 'homography', B)
        data = data_to_device(data, device)
        prob_o = net(data['optical'])['prob']
        prob_t = net(data['thermal'])['prob']
        if pred['nms'] > 0:
            prob_o = box_nms(prob_o, pred['nms'], thr, keep_top_k=pred['topk'], on_cpu=pred['cpu_nms'])
            prob_t = box_nms(prob_t, pred['nms'], thr, keep_top_k=pred['topk'], on_cpu=pred['cpu_nms'])
        H, W = prob_o.shape[-2:]
        kp_o, c_o, n_o = _keypoints(prob_o, thr, data['optical']['valid_mask'])
        kp_t, c_t, n_t = _keypoints(prob_t, thr, data['thermal']['valid_mask'])
        n_kp_optical += n_o
        n_kp_thermal += n_t
        dev = kp_o.device
        inv_o, inv_t = h_o.inverse().double().to(dev), h_t.inverse().double().to(dev)
        fwd_o, fwd_t = h_o.double().to(dev), h_t.double().to(dev)
        # optical -> common frame -> thermal frame, truncated to int after each step like warp_keypoints (:166-168)
        w_o = ops.warp_keypoints(ops.warp_keypoints(kp_o, inv_o, c_o), fwd_t, c_o)
        w_t = ops.warp_keypoints(ops.warp_keypoints(kp_t, inv_t, c_t), fwd_o, c_t)
        d2_t = ops.points_min_dist2(w_t, kp_o, H, W, c_t, c_o)     # dist1 (:186): warped thermal vs optical keypoints
        d2_o = ops.points_min_dist2(w_o, kp_t, H, W, c_o, c_t)     # dist2 (:187)
        stats = []
        for d2, cq in ((d2_t, c_t), (d2_o, c_o)):
            ok = (d2 >= 0) & _live(cq, d2.shape[1])                # filter_points survivors
            near = ok & (torch.sqrt(d2.clamp(min=0).double()) <= distance_thresh)
            stats += [ok.sum(1), near.sum(1)]
        n_th, cnt1, n_op, cnt2 = (s.tolist() for s in torch.stack(stats).cpu())
        for b in range(B):
            if n_th[b] + n_op[b] > 0:
                repeatability.append((cnt1[b] + cnt2[b]) / (n_th[b] + n_op[b]))
    return np.mean(repeatability), repeatability, n_kp_optical, n_kp_thermal


def _sorted_by_distance(q, t, d):
    order = np.argsort(d, kind='stable')   # sorted(matches, key=distance) is stable (:284-285)
    return q[order], t[order], d[order]


def _pad_matches(lists, dev):
    cap = max(1, max(len(x[0]) for x in lists))
    mq = torch.zeros((len(lists), cap), dtype=torch.int32)
    mt = torch.zeros((len(lists), cap), dtype=torch.int32)
    for b, (q, t, _) in enumerate(lists):
        mq[b, :len(q)] = torch.from_numpy(q.astype(np.int32))
        mt[b, :len(t)] = torch.from_numpy(t.astype(np.int32))
    nm = torch.tensor([len(x[0]) for x in lists], dtype=torch.int32)
    return mq.to(dev), mt.to(dev), nm.to(dev)


def compute_descriptor_metrics(net, dataloader, device, config, threshold_keypoints, threshold_warp):
    """Matching precision/recall, matching score and homography estimation accuracy (evaluation.py:209-438)."""
    import cv2
    acc = {s: {'tp': [], 'distance': [], 'n_gt': 0, 'm_score': []} for s in ('optical', 'thermal')}
    pts_dist = []
    thr = config['detection_threshold']
    for data in dataloader:
        data = data_to_device(data, device)
        out_o, out_t = net(data['optical']), net(data['thermal'])
        prob_o = out_o['prob'] * data['optical']['valid_mask']
        prob_t = out_t['prob'] * data['thermal']['valid_mask']
        if config['nms'] > 0:
            prob_t = box_nms(prob_t, config['nms'], thr, keep_top_k=config['topk'], on_cpu=config['cpu_nms'])
            prob_o = box_nms(prob_o, config['nms'], thr, keep_top_k=config['topk'], on_cpu=config['cpu_nms'])
        B = data['optical']['image'].shape[0]
        H_o, W_o = data['optical']['image'].shape[2:]
        H_t, W_t = data['thermal']['image'].shape[2:]
        h_o = _host_homographies(data['optical'], 'homography', B)
        h_t = _host_homographies(data['thermal'], 'homography', B)
        gt = torch.bmm(h_t, h_o.inverse())                         # optical -> thermal (:263)
        gt_inv = gt.inverse()

        kp_o, c_o, n_o = _keypoints(prob_o, thr)
        kp_t, c_t, n_t = _keypoints(prob_t, thr)
        dev = kp_o.device
        desc_o = ops.sample_descriptors(kp_o, out_o['desc'].to(dev, torch.float32), H_o, W_o, counts=c_o)
        desc_t = ops.sample_descriptors(kp_t, out_t['desc'].to(dev, torch.float32), H_t, W_t, counts=c_t)

        # ground-truth positions in the other spectrum, float64 like warp_keypoints(..., np.float) (:288-289)
        w_o = ops.warp_keypoints(kp_o, gt.double().to(dev), c_o, as_int=False)
        w_t = ops.warp_keypoints(kp_t, gt_inv.double().to(dev), c_t, as_int=False)

        # the three matcher calls per pair (:272-283, :330-336)
        m_th, m_op, m_cfg = [], [], []
        mcfg = config['matching']
        for b in range(B):
            empty = (np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float32))
            if n_o[b] > 0 and n_t[b] > 0:
                do, dt = desc_o[b, :n_o[b]], desc_t[b, :n_t[b]]
                m_th.append(tuple(x.cpu().numpy() for x in match_descriptors(dt, do, 'bfmatcher', False, crossCheck=True)))
                m_op.append(tuple(x.cpu().numpy() for x in match_descriptors(do, dt, 'bfmatcher', False, crossCheck=True)))
                m_cfg.append(tuple(x.cpu().numpy() for x in match_descriptors(do, dt, mcfg['method'], mcfg['knn_matches'],
                                                                              **mcfg['method_kwargs'])))
            else:
                m_th.append(empty)
                m_op.append(empty)
                m_cfg.append(empty)
        m_op = [_sorted_by_distance(*m) for m in m_op]
        m_th = [_sorted_by_distance(*m) for m in m_th]

        # correct-match flags for the matched pairs and per-row "has a correct partner" (:291-315)
        any_o, tp_o = ops.points_correct(w_o, kp_t, threshold_keypoints, c_o, c_t, *_pad_matches(m_op, dev))
        any_t, tp_t = ops.points_correct(w_t, kp_o, threshold_keypoints, c_t, c_o, *_pad_matches(m_th, dev))
        gt_rows = torch.stack([(any_o.bool() & _live(c_o, any_o.shape[1])).sum(1),
                               (any_t.bool() & _live(c_t, any_t.shape[1])).sum(1)]).cpu()
        # filter_points on the float64 warped points (:319-320)
        inside = []
        for w, c, (Hh, Ww) in ((w_o, c_o, (H_o, W_o)), (w_t, c_t, (H_o, W_o))):
            ok = (w[..., 0] >= 0) & (w[..., 1] >= 0) & (w[..., 0] < Hh) & (w[..., 1] < Ww) & _live(c, w.shape[1])
            inside.append(ok.sum(1))
        inside = torch.stack(inside).cpu()
        tp_o, tp_t = tp_o.cpu().numpy().astype(bool), tp_t.cpu().numpy().astype(bool)
        kp_o_h, kp_t_h = kp_o.cpu().numpy(), kp_t.cpu().numpy()

        for b in range(B):
            for s, matches, tp, row in (('optical', m_op[b], tp_o[b], 0), ('thermal', m_th[b], tp_t[b], 1)):
                n = len(matches[0])
                acc[s]['n_gt'] += int(gt_rows[row, b])
                acc[s]['tp'] += tp[:n].tolist()
                acc[s]['distance'] += matches[2].tolist()
                n_in = int(inside[row, b])
                acc[s]['m_score'].append(float(tp[:n].sum()) / n_in if n_in > 0 else 0.0)

            # homography from the configured matcher's pairs (:338-358); OpenCV points are (x,y)
            q, t, _ = m_cfg[b]
            optical_pts = kp_o_h[b, q][:, ::-1].astype(np.float32).reshape(-1, 1, 2)
            thermal_pts = kp_t_h[b, t][:, ::-1].astype(np.float32).reshape(-1, 1, 2)
            H_est = None
            if len(q) >= 4:
                H_est, _ = cv2.findHomography(optical_pts, thermal_pts, cv2.RANSAC,
                                              ransacReprojThreshold=config['reprojection_threshold'])
            if H_est is not None:
                corners = np.array([[0, 0], [H_o, 0], [0, W_o], [H_o, H_o]])  # sic: the reference's fourth corner (:353)
                d = warp_keypoints(corners, H_est, float) - warp_keypoints(corners, gt[b].numpy(), float)
                pts_dist.append(np.linalg.norm(d, axis=1).sum() / 4)
            else:
                pts_dist.append(999.0)

    out = {}
    curves = {}
    for s in ('optical', 'thermal'):
        tp = np.array(acc[s]['tp'])
        dist = np.array(acc[s]['distance'])
        order = np.argsort(dist)                                   # ascending descriptor distance (:371-380)
        tp, dist = tp[order], dist[order]
        fp = np.logical_not(tp)
        tp_cum, fp_cum = np.cumsum(tp), np.cumsum(fp)
        recall = np.concatenate([[0], div0(tp_cum, acc[s]['n_gt']), [1]])
        precision = np.concatenate([[0], div0(tp_cum, tp_cum + fp_cum), [0]])
        precision = np.maximum.accumulate(precision[::-1])[::-1]
        curves[s] = (precision, recall)
        out.update({'tp_' + s: tp, 'fp_' + s: fp, 'distance_' + s: dist, 'recall_' + s: recall, 'precision_' + s: precision,
                    'nn_map_' + s: compute_mAP(precision, recall), 'm_score_' + s: np.array(acc[s]['m_score'])})
    pts_dist = np.array(pts_dist)
    out['nn_map'] = (out['nn_map_optical'] + out['nn_map_thermal']) * 0.5
    out['m_score'] = (out['m_score_optical'].mean() + out['m_score_thermal'].mean()) * 0.5
    out['pts_dist'] = pts_dist
    out['average_h_error'] = pts_dist.mean()
    out['h_correctness'] = (pts_dist < threshold_warp).sum() / len(pts_dist)
    return out
