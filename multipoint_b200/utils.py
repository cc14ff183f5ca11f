"""Host-side mirror of the reference's ``multipoint.utils`` helpers that sit on the hot path
(reference: multipoint/utils/utils.py, matching.py, homographies.py).  Same names, argument
meaning and error behaviour; the arithmetic runs in this repo's CUDA kernels through the C ABI
(multipoint_b200.ops).  There is no CPU fallback: tensors living on the host are moved to the
current CUDA device, processed there, and the result is returned on the original device, the way
the reference returns ``prob_nms.to(device)`` (utils.py:122).

Deliberate, documented deviations from the reference
  - ``box_nms(on_cpu=...)`` is accepted and ignored (utils.py:94-95 copies to the host only to
    dodge torchvision's slow GPU path).
  - ``homographic_adaptation*`` deep-copy the default config; the reference merges overrides into
    the module-level dict so they persist across calls (homographies.py:40,132).
  - ``get_matches(method='flann')`` raises NotImplementedError: approximate and non-deterministic,
    never selected by a shipped config.
"""
import collections.abc
import copy
from math import pi

import numpy as np
import torch
import torch.nn as nn

from . import ops

# ----------------------------------------------------------------------------- glue (utils.py:10-62,169-175)


def dict_update(d, u):
    """Nested dictionary update (utils.py:10-26)."""
    for k, v in u.items():
        if isinstance(v, collections.abc.Mapping):
            d[k] = dict_update(d.get(k, {}), v)
        else:
            d[k] = v
    return d


def data_to_device(data, device):
    for key in data.keys():
        if type(data[key]) is torch.Tensor:
            data[key] = data[key].to(device)
        elif type(data[key]) is dict:
            data[key] = data_to_device(data[key], device)
    return data


def data_unsqueeze(data, dim):
    for key in data.keys():
        if type(data[key]) is torch.Tensor:
            data[key] = data[key].unsqueeze(dim)
        elif type(data[key]) is dict:
            data[key] = data_unsqueeze(data[key], dim)
    return data


def fix_model_weigth_keys(weights):
    """Strip everything up to the last '__' of each key (DataParallel / renamed checkpoints, utils.py:169-175)."""
    return collections.OrderedDict((key.split('__')[-1], value) for key, value in weights.items())


def _device_for(t):
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("multipoint_b200 needs a CUDA device: the hot path has no CPU implementation")
    return torch.device('cuda', torch.cuda.current_device())


# ----------------------------------------------------------------------------- utils.py:64-76


def depth_to_space(x, block_size):
    """utils.depth_to_space (utils.py:64-69); block 8 on 64 channels is PixelShuffle(8)."""
    dev = x.device
    out = ops.depth_to_space(x.to(_device_for(x), torch.float32), block_size)
    return out.to(dev)


def space_to_depth(x, block_size):
    """utils.space_to_depth (utils.py:71-76): used by the loss only; a pure permute, kept in torch."""
    N, C, H, W = x.size()
    x = x.view(N, C, H // block_size, block_size, W // block_size, block_size)
    x = x.permute(0, 3, 5, 1, 2, 4).contiguous()
    return x.view(N, C * (block_size ** 2), H // block_size, W // block_size)


# ----------------------------------------------------------------------------- utils.py:78-122


def box_nms(prob, size, min_prob, iou=0.1, keep_top_k=0, on_cpu=False):
    """Drop-in for utils.box_nms (utils.py:78-122): greedy IoU NMS of size x size boxes centred on
    every pixel above ``min_prob``, optional top-k, scores scattered into a zero map.
    prob: (H,W) or (B,1,H,W).  Bit-exact against the reference (tests/test_gpu_parity.py)."""
    if not (len(prob.shape) == 2 or len(prob.shape) == 4):
        raise ValueError('The probability must be either 2D (H,W), or 4D (B, 1, H, W)')
    device = prob.device
    p = prob.to(_device_for(prob), torch.float32)
    H, W = p.shape[-2:]
    out = ops.box_nms(p.reshape(-1, H, W), size, min_prob, iou=iou, keep_top_k=keep_top_k)
    return out.reshape(prob.shape).to(device)


def box_nms_keypoints(prob, size, min_prob, iou=0.1, keep_top_k=0, kp_cap=None):
    """box_nms plus the keypoint idiom ``torch.nonzero((nms > thr).float())`` in one pass, without
    the host sync of nonzero: returns (dense (B,1,H,W), keypoints (B,cap,2) int64 (y,x) row-major,
    scores (B,cap), counts (B) int32)."""
    if len(prob.shape) != 4:
        raise ValueError('The probability must be 4D (B, 1, H, W)')
    p = prob.to(_device_for(prob), torch.float32)
    B, _, H, W = p.shape
    dense, kp, sc, cnt = ops.box_nms(p.reshape(B, H, W), size, min_prob, iou=iou, keep_top_k=keep_top_k,
                                     want_keypoints=True, kp_cap=kp_cap)
    return dense.reshape(B, 1, H, W), kp, sc, cnt


def extract_keypoints(prob, threshold, mask=None):
    """``torch.nonzero((prob.squeeze() > threshold).float() [* mask])`` for one (H,W) map
    (predict_align_image_pair.py:170-171, evaluation.py:157-158): (K,2) int64 (y,x), row-major."""
    p = prob.squeeze()
    dev = p.device
    p = p.to(_device_for(p), torch.float32)
    m = None if mask is None else mask.squeeze().to(p.device)[None]
    kp, _, cnt = ops.extract_keypoints(p[None], threshold, m)
    return kp[0, :int(cnt[0])].to(dev)


# ----------------------------------------------------------------------------- utils.py:124-157


def get_gaussian_filter(kernel_size, sigma=None, channels=1):
    """utils.get_gaussian_filter (utils.py:124-157): frozen depthwise Gaussian Conv2d.  Only used
    when filter_size > 0 (off in every shipped config); stays a cuDNN convolution."""
    if sigma is None:
        sigma = 0.3 * ((kernel_size - 1) * 0.5 - 1) + 0.8
    coords = torch.arange(kernel_size)
    xg = coords.repeat(kernel_size).view(kernel_size, kernel_size)
    grid = torch.stack([xg, xg.t()], dim=-1)
    mean = (kernel_size - 1) / 2.
    var = sigma ** 2.
    k = (1. / (2. * np.pi * var)) * torch.exp(-torch.sum((grid - mean) ** 2., dim=-1) / (2 * var))
    k = (k / torch.sum(k)).view(1, 1, kernel_size, kernel_size).repeat(channels, 1, 1, 1)
    f = nn.Conv2d(channels, channels, kernel_size=kernel_size, groups=channels, bias=False)
    f.weight.data = k
    f.weight.requires_grad = False
    return f


# ----------------------------------------------------------------------------- utils.py:159-167


def interpolate_descriptors(keypoints, descriptors_lowres, H, W):
    """Drop-in for utils.interpolate_descriptors (utils.py:159-167): keypoints (K,2) (y,x),
    descriptors_lowres (D,Hc,Wc) -> (K,D) unit-norm rows.  Never mutates ``keypoints`` (the
    reference mutates a float32 input in place, utils.py:160-163)."""
    dev = descriptors_lowres.device
    cdev = _device_for(descriptors_lowres)
    kp = keypoints
    if kp.is_floating_point():
        if kp.numel() and not bool((kp == kp.round()).all()):
            raise NotImplementedError("interpolate_descriptors: fractional keypoints are not supported "
                                      "(every reference call site passes torch.nonzero output)")
        kp = kp.to(torch.int64)
    D = descriptors_lowres.shape[0]
    if kp.shape[0] == 0:
        return torch.zeros((0, D), dtype=torch.float32, device=dev)
    out = ops.sample_descriptors(kp.to(cdev, torch.int64).reshape(1, -1, 2).contiguous(),
                                 descriptors_lowres.to(cdev, torch.float32)[None], H, W)
    return out[0].to(dev)


# ----------------------------------------------------------------------------- matching.py


def _dmatches(q, t, d):
    import cv2
    return [cv2.DMatch(int(a), int(b), float(c)) for a, b, c in zip(q, t, d)]


def match_descriptors(desc_1, desc_2, method='bfmatcher', knn_matches=False, **kwargs):
    """Tensor-level core of get_matches: CUDA tensors (N1,D), (N2,D) -> (query, train, distance)
    CUDA tensors in ascending query order.  One device->host read (the match count)."""
    if method == 'bfmatcher':
        cross = bool(kwargs.pop('crossCheck', False))
        if kwargs:
            raise TypeError('unexpected BFMatcher arguments: %s' % sorted(kwargs))
        if knn_matches:
            if cross:
                raise ValueError('BFMatcher.knnMatch(k=2) is not available with crossCheck=True '
                                 '(OpenCV asserts K == 1 && update == 0)')
            if desc_2.shape[0] < 2:
                # knnMatch(k=2) returns 1-element lists here and the reference's ``for m, n in all_matches``
                # (matching.py:24) raises exactly this
                raise ValueError('not enough values to unpack (expected 2, got %d)' % desc_2.shape[0])
            q, t, d, c = ops.match(desc_1, desc_2, metric='l2', kind='ratio', ratio=0.9)
        else:
            q, t, d, c = ops.match(desc_1, desc_2, metric='l2', kind='mutual', cross_check=cross)
    elif method == 'nnmatcher':
        thr = NNMatcher(**kwargs).nn_thresh
        if knn_matches:
            raise AttributeError("'NNMatcher' object has no attribute 'knnMatch'")
        q, t, d, c = ops.match(desc_1, desc_2, metric='nn', kind='mutual', cross_check=True, threshold=thr)
    elif method == 'thresholdmatcher':
        thr = ThresholdMatcher(**kwargs).threshold
        if knn_matches:
            raise AttributeError("'ThresholdMatcher' object has no attribute 'knnMatch'")
        return ops.match_threshold(desc_1, desc_2, thr)
    elif method == 'flann':
        raise NotImplementedError('flann matching is out of scope (approximate, non-deterministic)')
    else:
        raise ValueError('unknown matching method')
    n = int(c[0])
    return q[0, :n], t[0, :n], d[0, :n]


def get_matches(desc_1, desc_2, method='bfmatcher', knn_matches=False, **kwargs):
    """Drop-in for utils.get_matches (matching.py:4-33): host numpy (N,D) fp32 in, list of
    cv2.DMatch out, computed on the GPU.  Same kwargs: crossCheck (bfmatcher), threshold
    (nnmatcher / thresholdmatcher)."""
    if method not in ('bfmatcher', 'flann', 'nnmatcher', 'thresholdmatcher'):
        raise ValueError('unknown matching method')
    if method == 'nnmatcher':
        return NNMatcher(**kwargs).match(desc_1, desc_2) if not knn_matches else NNMatcher(**kwargs).knnMatch(desc_1, desc_2, 2)
    if method == 'thresholdmatcher':
        return ThresholdMatcher(**kwargs).match(desc_1, desc_2) if not knn_matches else ThresholdMatcher(**kwargs).knnMatch(desc_1, desc_2, 2)
    d1, d2 = _to_cuda_desc(desc_1), _to_cuda_desc(desc_2)
    if d1.shape[0] == 0 or d2.shape[0] == 0:
        return []
    q, t, d = match_descriptors(d1, d2, method, knn_matches, **kwargs)
    return _dmatches(q.cpu().numpy(), t.cpu().numpy(), d.cpu().numpy())


def _to_cuda_desc(d):
    if isinstance(d, torch.Tensor):
        return d.to(_device_for(d), torch.float32).contiguous()
    if not torch.cuda.is_available():
        raise RuntimeError("multipoint_b200 needs a CUDA device: the hot path has no CPU implementation")
    return torch.from_numpy(np.ascontiguousarray(d, dtype=np.float32)).cuda()


class NNMatcher():
    """Drop-in for matching.NNMatcher (matching.py:35-72): mutual nearest neighbour on
    sqrt(2 - 2 clip(a.b)) with a distance threshold; descriptors are assumed unit-norm."""

    def __init__(self, threshold=0.7):
        self.nn_thresh = threshold
        if threshold < 0.0:
            raise ValueError('\'threshold\' should be non-negative')

    def match(self, desc1, desc2):
        assert desc1.shape[1] == desc2.shape[1]
        if desc1.shape[0] == 0 or desc2.shape[0] == 0:
            return []
        q, t, d, c = ops.match(_to_cuda_desc(desc1), _to_cuda_desc(desc2), metric='nn', kind='mutual',
                               cross_check=True, threshold=self.nn_thresh)
        n = int(c[0])
        return _dmatches(q[0, :n].cpu().numpy(), t[0, :n].cpu().numpy(), d[0, :n].cpu().numpy())


class ThresholdMatcher():
    """Drop-in for matching.ThresholdMatcher (matching.py:74-99): every pair under the threshold."""

    def __init__(self, threshold=0.4):
        self.threshold = threshold
        if threshold < 0.0:
            raise ValueError('\'threshold\' should be non-negative')

    def match(self, desc1, desc2):
        assert desc1.shape[1] == desc2.shape[1]
        if desc1.shape[0] == 0 or desc2.shape[0] == 0:
            return []
        q, t, d = ops.match_threshold(_to_cuda_desc(desc1), _to_cuda_desc(desc2), self.threshold)
        return _dmatches(q.cpu().numpy(), t.cpu().numpy(), d.cpu().numpy())


# ----------------------------------------------------------------------------- homographies.py (host side)

homography_adaptation_default_config = {
    'num': 100,
    'aggregation': 'prod',
    'homographies': {
        'translation': True, 'rotation': True, 'scaling': True, 'perspective': True,
        'scaling_amplitude': 0.15, 'perspective_amplitude_x': 0.15, 'perspective_amplitude_y': 0.15,
        'patch_ratio': 0.9, 'max_angle': pi, 'allow_artifacts': True,
    },
    'erosion_radius': 5,
    'mask_border': True,
    'min_count': 2,
    'filter_size': 0,
}


def sample_homography(image_shape, perspective=True, scaling=True, rotation=True, translation=True,
                      n_scales=10, n_angles=25, scaling_amplitude=0.2, perspective_amplitude_x=0.1,
                      perspective_amplitude_y=0.1, patch_ratio=0.8, max_angle=pi / 2,
                      allow_artifacts=True, translation_overflow=0.1):
    """Random homography exactly as the reference draws it (homographies.py:191-329): the corners
    of a centred patch are perturbed by up to four transforms applied in a shuffled order, each
    consuming the numpy *global* RNG in the reference's order, then
    ``cv2.getPerspectiveTransform(unit corners * (W,H), patch corners * (W,H))`` on float32 points.
    Stays on the host: the RNG stream and OpenCV's solver define the result."""
    import cv2
    rng = np.random  # the global stream, like the reference

    def perspective_step(pts):
        lo, hi = -pts.min(axis=0), 1.0 - pts.max(axis=0)
        hi[1] = min(abs(lo[1]), abs(hi[1]))
        lo[1] = -hi[1]
        amp = np.array([perspective_amplitude_x, perspective_amplitude_y])
        if allow_artifacts:
            a_min, a_max = -amp, amp
        else:
            a_min, a_max = np.maximum(-amp, lo), np.minimum(amp, hi)
        dy = rng.uniform(a_min[1], a_max[1])
        dx_left = rng.uniform(a_min[0], a_max[0])
        dx_right = rng.uniform(a_min[0], a_max[0])
        pts += np.array([[dx_left, dy], [dx_left, -dy], [dx_right, dy], [dx_right, -dy]])
        return pts

    def scale_step(pts):
        scales = rng.uniform(-scaling_amplitude, scaling_amplitude, n_scales) + 1.0
        centre = pts.mean(axis=0)
        cand = np.expand_dims(pts - centre, 0) * np.expand_dims(np.expand_dims(scales, 1), 1) + centre
        if allow_artifacts:
            ok = np.arange(n_scales)
        else:
            ok = [i for i in range(n_scales) if cand[i, ...].max() < 1.0 and cand[i, ...].min() >= 0.0]
        return cand[rng.choice(ok)]

    def translation_step(pts):
        lo, hi = -pts.min(axis=0), 1.0 - pts.max(axis=0)
        if allow_artifacts:
            lo -= translation_overflow
            hi += translation_overflow
        pts += np.array([rng.uniform(lo[0], hi[0]), rng.uniform(lo[1], hi[1])])
        return pts

    def rotation_step(pts):
        angles = np.append(rng.uniform(-max_angle, max_angle, n_angles), 0)  # 0 = fallback when nothing fits
        centre = pts.mean(axis=0)
        rot = np.reshape(np.stack([np.cos(angles), -np.sin(angles), np.sin(angles), np.cos(angles)], axis=1), [-1, 2, 2])
        cand = np.matmul(np.tile(np.expand_dims(pts - centre, axis=0), [n_angles + 1, 1, 1]), rot) + centre
        if allow_artifacts:
            ok = np.arange(n_angles)
        else:
            ok = [i for i in range(len(angles)) if cand[i, ...].max() < 1.0 and cand[i, ...].min() >= 0.0]
        return cand[rng.choice(ok)]

    corners = np.array([[0., 0.], [0., 1.], [1., 1.], [1., 0.]])
    patch = (1 - patch_ratio) * 0.5 + patch_ratio * corners
    steps = [fn for enabled, fn in ((perspective, perspective_step), (scaling, scale_step),
                                    (translation, translation_step), (rotation, rotation_step)) if enabled]
    order = np.arange(len(steps))
    rng.shuffle(order)
    for idx in order:
        patch = steps[idx](patch)
    wh = image_shape[::-1]  # (H,W) -> (W,H): points are (x,y)
    corners *= wh
    patch *= wh
    return cv2.getPerspectiveTransform(corners.astype(np.float32), patch.astype(np.float32))


def compute_valid_mask(image_shape, homography, erosion_radius=0, mask_border=False):
    """Valid-pixel mask of a warped image (homographies.py:375-402): nearest-neighbour warp of ones,
    optional one-pixel zero frame, erosion by a (2r+1)^2 box.  Host / OpenCV: the raster rule of
    cv2.warpPerspective defines the result."""
    import cv2
    mask = cv2.warpPerspective(np.ones(image_shape), homography, image_shape[::-1], flags=cv2.INTER_NEAREST)
    if erosion_radius > 0:
        if mask_border:
            framed = np.zeros((image_shape[0] + 2, image_shape[1] + 2))
            framed[1:-1, 1:-1] = mask
            mask = framed
        kernel = np.ones((erosion_radius * 2 + 1, erosion_radius * 2 + 1), np.float32)
        mask = cv2.erode(mask, kernel, iterations=1)
        if mask_border:
            mask = mask[1:-1, 1:-1]
    return mask


def warp_keypoints(keypoints, homography, return_type=int):
    """homographies.py:331-346: (N,2) (y,x) points through a 3x3 matrix, truncated to return_type."""
    import cv2
    if len(keypoints) > 0:
        warped = cv2.perspectiveTransform(np.array([keypoints[:, ::-1]], dtype=np.float64), homography)
        return warped[0, :, ::-1].astype(return_type)
    return keypoints


def warp_points_pytorch(points, homography):
    """homographies.py:348-356."""
    h = torch.cat([points.flip(-1), torch.ones([points.shape[0], points.shape[1], 1], dtype=torch.float32, device=points.device)], -1)
    w = torch.bmm(homography, h.permute([0, 2, 1])).permute([0, 2, 1])
    return (w[:, :, :2] / w[:, :, 2:]).flip(-1)


def filter_points(points, shape):
    """homographies.py:358-373: drop points outside [0,H) x [0,W)."""
    points = points[points[:, 0] >= 0]
    points = points[points[:, 1] >= 0]
    points = points[points[:, 0] < shape[0]]
    points = points[points[:, 1] < shape[1]]
    return points


# ----------------------------------------------------------------------------- homographies.py:404-432


def normalized_warp_matrix(M, src_hw, dsize):
    """The matrix kornia's homography_warp hands to grid_sample for warp_perspective_tensor(src, M)
    (homographies.py:424-425): M_norm = N_dst @ (M @ N_src^-1) with N = pixel -> [-1,1] using
    (size-1) denominators, then A = inverse(M_norm).  fp32 torch ops on the host, (n,3,3) in/out.
    PARITY UNPINNED against kornia itself (not installed, no version pinned)."""
    M = M.detach().to('cpu', torch.float32).reshape(-1, 3, 3)

    def norm_mat(h, w):
        return torch.tensor([[2.0 / (w - 1), 0.0, -1.0], [0.0, 2.0 / (h - 1), -1.0], [0.0, 0.0, 1.0]], dtype=torch.float32)

    n_src, n_dst = norm_mat(*src_hw), norm_mat(*dsize)
    m_norm = n_dst @ (M @ torch.inverse(n_src))
    return torch.inverse(m_norm)


def warp_perspective_tensor(src, M, dsize, mode='bilinear', padding_mode='zeros'):
    """Drop-in for homographies.warp_perspective_tensor (:404-425): src (B,C,H,W), M (B,3,3)
    pixel-space homographies (source -> destination); output pixel p samples the source at M^-1 p.
    dsize must equal the source size (every reference call site passes image_shape[2:])."""
    if not torch.is_tensor(src):
        raise TypeError("Input src type is not a torch.Tensor. Got {}".format(type(src)))
    if not torch.is_tensor(M):
        raise TypeError("Input M type is not a torch.Tensor. Got {}".format(type(M)))
    if not len(src.shape) == 4:
        raise ValueError("Input src must be a BxCxHxW tensor. Got {}".format(src.shape))
    if not (len(M.shape) == 3 or M.shape[-2:] == (3, 3)):
        raise ValueError("Input M must be a Bx3x3 tensor. Got {}".format(src.shape))
    B, C, H, W = src.shape
    if tuple(dsize) != (H, W):
        raise NotImplementedError("warp_perspective_tensor: dsize must equal the source size")
    dev = src.device
    cdev = _device_for(src)
    s = src.to(cdev, torch.float32)
    A = normalized_warp_matrix(M, (H, W), (H, W)).to(cdev)
    if A.shape[0] == 1 and B > 1:
        A = A.expand(B, 3, 3).contiguous()
    same = bool((A == A[:1]).all())
    if same:  # one matrix for the whole batch (how homographic adaptation calls it)
        out = ops.warp(s.reshape(B * C, H, W), A[:1].contiguous(), mode, padding_mode)[0]
    else:
        out = torch.stack([ops.warp(s[b], A[b:b + 1].contiguous(), mode, padding_mode)[0] for b in range(B)])
    return out.reshape(B, C, H, W).to(dev)


class WarpingModule(nn.Module):
    """homographies.py:427-432."""

    def forward(self, src, M, dsize, mode='bilinear', padding_mode='zeros'):
        return warp_perspective_tensor(src, M, dsize, mode, padding_mode)


# ----------------------------------------------------------------------------- homographies.py:38-189


def _check_ha_config(user_config):
    config = dict_update(copy.deepcopy(homography_adaptation_default_config), user_config)
    if config['num'] < 1:
        raise ValueError('num must be larger than 0 for the homographic adaptation')
    if config['filter_size'] % 2 == 0 and config['filter_size'] != 0:
        raise ValueError('The filter_size must be uneven')
    return config


def invert_homographies(Hs):
    """cv::invert for 3x3 doubles (cofactors times 1/det, zeros when det == 0), vectorised over a
    batch in the same operation order, so the result is bit-identical to what cv2.warpPerspective
    inverts internally (checked against cv2.invert in tests/test_host.py)."""
    S = np.asarray(Hs, np.float64).reshape(-1, 9)
    a, b, c, d, e, f, g, h, i = (S[:, k] for k in range(9))
    det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g)
    with np.errstate(divide='ignore', invalid='ignore'):
        r = 1.0 / det
        T = np.stack([(e * i - f * h) * r, (c * h - b * i) * r, (b * f - c * e) * r,
                      (f * g - d * i) * r, (a * i - c * g) * r, (c * d - a * f) * r,
                      (d * h - e * g) * r, (b * g - a * h) * r, (a * e - b * d) * r], axis=1)
    T[det == 0.0] = 0.0
    return T.reshape(-1, 3, 3)


def compute_valid_masks(image_shape, homographies, erosion_radius=0, mask_border=False, device=None):
    """Batched, device-side compute_valid_mask: (n,3,3) homographies ->
    (n,H,W) uint8 CUDA tensor, bit-identical to the per-homography cv2 path above, one launch."""
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    Minv = torch.from_numpy(invert_homographies(homographies)).to(dev)
    return ops.valid_masks(Minv, image_shape[0], image_shape[1], erosion_radius, mask_border)


def sample_adaptation_homographies(image_hw, config, with_masks=True):
    """Pre-sample the num-1 homographies (and valid masks) in the reference's RNG order
    (homographies.py:77-80 / :162-165: sample_homography then compute_valid_mask per iteration;
    the mask consumes no random numbers).  Returns (H (n,3,3) float64, masks (n,H,W) uint8) with
    the masks rastered by cv2 on the host, or (H, None) with ``with_masks=False``: the adaptation
    then builds them on the device (compute_valid_masks), which is what the product path does."""
    n = config['num'] - 1
    Hs = np.zeros((n, 3, 3), np.float64)
    masks = np.zeros((n,) + tuple(image_hw), np.uint8) if with_masks else None
    for i in range(n):
        Hs[i] = sample_homography(np.array(image_hw), **config['homographies'])
        if with_masks:
            masks[i] = compute_valid_mask(tuple(image_hw), Hs[i], config['erosion_radius'], config['mask_border']) != 0
    return Hs, masks


def _net_prob(net, images, extra, chunk_imgs):
    """Run ``net`` on (n*B,1,H,W) images in chunks; returns (n*B,H,W) heatmaps."""
    outs = []
    for s in range(0, images.shape[0], chunk_imgs):
        batch = {'image': images[s:s + chunk_imgs]}
        for k, v in extra.items():
            batch[k] = v[s:s + chunk_imgs]
        outs.append(net(batch)['prob'][:, 0])
    return torch.cat(outs) if len(outs) > 1 else outs[0]


def _adaptation_core(images, is_optical, net, config, second, homographies, masks, rank, world, fused, normalized=None):
    """Identity pass + this rank's share of the sampled homographies.
    fused=True  -> the finished heatmap (B,H,W) in one aggregate launch (single process).
    fused=False -> partial accumulators (prob_sum, count_sum); rank 0's include the identity pass."""
    dev = _device_for(images)
    images = images.to(dev, torch.float32)
    B, _, H, W = images.shape
    n_total = config['num'] - 1
    agg = config['aggregation'] if second is not None else 'none'
    if second is not None and agg not in ('prod', 'sum'):
        raise ValueError('Unknown aggregation: ' + str(config['aggregation']))
    if config['filter_size'] > 0:
        gauss = get_gaussian_filter(config['filter_size']).to(dev)
        pad = nn.ReflectionPad2d(int((config['filter_size'] - 1) / 2))
        post = lambda p: gauss(pad(p[:, None]))[:, 0]  # noqa: E731
    else:
        post = lambda p: p  # noqa: E731

    def run(imgs, opt_col):
        extra = {} if opt_col is None else {'is_optical': opt_col.to(dev).repeat(imgs.shape[0] // B, 1)}
        return post(_net_prob(net, imgs, extra, chunk_imgs=max(B, 32)))

    # identity pass (homographies.py:51-66 / :143-154); only rank 0 contributes it
    prob0 = None
    img_b = opt_b = None
    if second is not None:
        img_b, opt_b = second
        img_b = img_b.to(dev, torch.float32)
    if rank == 0:
        prob0 = run(images, is_optical)
        if second is not None:
            pb0 = run(img_b, opt_b)
            prob0 = prob0 * pb0 if agg == 'prod' else prob0 + pb0
        prob0 = prob0.contiguous()

    if homographies is None:
        homographies, masks = sample_adaptation_homographies((H, W), config, with_masks=False)
    # round-robin over num units, unit 0 being the identity pass (rank 0's) and unit i + 1 sample i: rank 0 then gets one
    # sample fewer than the busiest rank instead of the identity pass on top of a full share
    mine = [i for i in range(n_total) if (i + 1) % world == rank]
    tables = ops.linspace_tables(H, W, dev)
    n = len(mine)
    empty = torch.zeros((0, B, H, W), device=dev)
    if n == 0:
        return ops.ha_aggregate(prob0, empty, None if second is None else empty, torch.zeros((0, H, W), dtype=torch.uint8, device=dev),
                                torch.zeros((0, 3, 3), device=dev), agg, config['min_count'], init=(rank == 0), finish=fused, tables=tables)
    H_mine = np.asarray(homographies, np.float64)[mine]
    if normalized is not None:   # (A_warp, A_unwarp) given: the per-pixel arithmetic on its own (parity tests)
        A_warp = torch.as_tensor(np.asarray(normalized[0], np.float32)[mine]).to(dev)
        A_unwarp = torch.as_tensor(np.asarray(normalized[1], np.float32)[mine]).to(dev)
    else:
        Hm = torch.from_numpy(H_mine.astype(np.float32))
        A_warp = normalized_warp_matrix(Hm, (H, W), (H, W)).to(dev)
        A_unwarp = normalized_warp_matrix(torch.inverse(Hm), (H, W), (H, W)).to(dev)  # torch.inverse(homography) :112,:180
    if masks is None:   # built on the device, this rank's share only
        mk = compute_valid_masks((H, W), H_mine, config['erosion_radius'], config['mask_border'], dev)
    elif torch.is_tensor(masks):
        mk = masks.to(dev, torch.uint8)[mine].contiguous()
    else:
        mk = torch.from_numpy(np.ascontiguousarray(np.asarray(masks)[mine])).to(dev, torch.uint8)
    # The samples go through warp -> net -> aggregate in chunks, so the warped images and their heatmaps take
    # O(chunk * B) memory instead of O(num * B); the accumulators carry on from chunk to chunk in the reference's
    # summation order (MP_HA_INIT on the first chunk, MP_HA_FINISH on the last), so the result does not depend on it.
    chunk = max(1, int(config.get('sample_chunk', 0)) or max(1, 64 // max(B, 1)))
    prob_acc = count_acc = out = None
    both_spectra = torch.cat([images[:, 0], img_b[:, 0]]) if second is not None else None
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        nc = c1 - c0
        pw_b = None
        if second is None:
            warped = ops.warp(images[:, 0], A_warp[c0:c1], 'bilinear', 'reflection', tables).reshape(nc * B, 1, H, W)  # :171
            pw_a = run(warped, is_optical).reshape(nc, B, H, W).contiguous()
        else:
            # both spectra are warped by the same matrices (:81, :86): one launch computes every sampling position once
            warped = ops.warp(both_spectra, A_warp[c0:c1], 'bilinear', 'reflection', tables, groups=2)
            pw_a = run(warped[0].reshape(nc * B, 1, H, W), is_optical).reshape(nc, B, H, W).contiguous()
            pw_b = run(warped[1].reshape(nc * B, 1, H, W), opt_b).reshape(nc, B, H, W).contiguous()
        del warped
        first, last = c0 == 0, c1 == n
        res = ops.ha_aggregate(prob0 if first else None, pw_a, pw_b, mk[c0:c1], A_unwarp[c0:c1], agg, config['min_count'],
                               init=(rank == 0 and first), finish=(fused and last), prob_acc=prob_acc, count_acc=count_acc,
                               tables=tables)
        if fused and last:
            out = res
        else:
            prob_acc, count_acc = res
    return out if fused else (prob_acc, count_acc)


def adaptation_finish(prob_sum, count_sum, aggregation, min_count):
    """out = prob/count, sqrt | *0.5, zero where count < min_count (homographies.py:116-126,184-187)
    on already summed accumulators (the step after the multi-GPU all-reduce)."""
    B, H, W = prob_sum.shape
    dev = prob_sum.device
    empty = torch.zeros((0, B, H, W), device=dev)
    return ops.ha_aggregate(None, empty, None if aggregation in (None, 'none') else empty,
                            torch.zeros((0, H, W), dtype=torch.uint8, device=dev), torch.zeros((0, 3, 3), device=dev),
                            aggregation, min_count, init=False, finish=True, prob_acc=prob_sum, count_acc=count_sum)


def _adaptation(images, is_optical, net, config, second=None, homographies=None, masks=None, shard=None, normalized=None):
    """Shared body of the two adaptation entry points.  ``shard=(rank, world, all_reduce)`` splits
    the sampled homographies round-robin across ranks (every rank must be given the same
    ``homographies`` / ``masks``, see parallel.broadcast_homographies) and sums the two accumulators
    with ``all_reduce`` before the finish.  Without ``shard`` everything runs in one fused launch
    in the reference's summation order."""
    if shard is None or shard[1] == 1:
        return _adaptation_core(images, is_optical, net, config, second, homographies, masks, 0, 1, True, normalized)[:, None]
    rank, world, all_reduce = shard
    if homographies is None:
        raise ValueError("sharded homographic adaptation needs pre-sampled homographies shared by all ranks")
    prob_sum, count_sum = _adaptation_core(images, is_optical, net, config, second, homographies, masks, rank, world, False, normalized)
    all_reduce(prob_sum)
    all_reduce(count_sum)
    agg = config['aggregation'] if second is not None else 'none'
    return adaptation_finish(prob_sum, count_sum, agg, config['min_count'])[:, None]


def homographic_adaptation(data, net, homographic_adaptation_config={}, homographies=None, masks=None, shard=None,
                           normalized_matrices=None):
    """Drop-in for utils.homographic_adaptation (homographies.py:130-189).  ``net`` is any callable
    dict -> {'prob': (B,1,H,W)}.  Optional ``homographies`` / ``masks`` replace the host sampling
    (tests, multi-GPU broadcast); ``shard`` see _adaptation; ``normalized_matrices=(A_warp, A_unwarp)`` (n,3,3)
    replaces the 3x3 normalisation algebra of warp_perspective_tensor by given matrices (parity tests pin the
    per-pixel arithmetic with it)."""
    config = _check_ha_config(homographic_adaptation_config)
    device = data['image'].device
    out = _adaptation(data['image'], data.get('is_optical'), net, config, None, homographies, masks, shard, normalized_matrices)
    return out.to(device)


def homographic_adaptation_multispectral(data, net, homographic_adaptation_config={}, homographies=None, masks=None, shard=None,
                                         normalized_matrices=None):
    """Drop-in for utils.homographic_adaptation_multispectral (homographies.py:38-128)."""
    config = _check_ha_config(homographic_adaptation_config)
    device = data['optical']['image'].device
    if config['aggregation'] not in ('prod', 'sum'):
        raise ValueError('Unknown aggregation: ' + config['aggregation'])
    out = _adaptation(data['optical']['image'], data['optical'].get('is_optical'), net, config,
                      (data['thermal']['image'], data['thermal'].get('is_optical')), homographies, masks, shard, normalized_matrices)
    return out.to(device)
