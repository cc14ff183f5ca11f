"""ctypes/numpy front end of the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of bench.py may import this module, and only as the checker.  Nothing under
``multipoint_b200/`` imports it.  The arithmetic lives in ``mp_oracle.c``; this file
only marshals numpy arrays and restates the tiny host-side matrix algebra of the
homography warp (kornia's ``dst_norm_to_dst_norm``; PARITY UNPINNED).

Reference sites are cited per function (paths relative to ethz-asl/multipoint).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    """Compile mp_oracle.c with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, "libmp_oracle.so")
    src = os.path.join(_HERE, "mp_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libmp_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.mpo_extract_keypoints.restype = ctypes.c_int64
        _LIB.mpo_match_threshold.restype = ctypes.c_int64
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


# --- row 1: MultiPoint.detector_head (multipoint/models/MultiPoint.py:150-158) ---
def detector_head(logits):
    logits = _f32(logits)
    B, C, Hc, Wc = logits.shape
    assert C == 65
    out = np.empty((B, 1, Hc * 8, Wc * 8), np.float32)
    lib().mpo_detector_head(_p(logits, _f32p), B, Hc, Wc, _p(out, _f32p))
    return out


# --- SuperPointMagicLeap.generate_heatmap (SuperPointMagicLeap.py:68-85) ---
def heatmap_magicleap(semi):
    semi = _f32(semi)
    B, C, Hc, Wc = semi.shape
    assert C == 65
    out = np.empty((B, 1, Hc * 8, Wc * 8), np.float32)
    lib().mpo_heatmap_magicleap(_p(semi, _f32p), B, Hc, Wc, _p(out, _f32p))
    return out


# --- row 2: MultiPoint.descriptor_head tail (MultiPoint.py:160-166) ---
def normalize_descriptors(x):
    x = _f32(x)
    B, D = x.shape[:2]
    HW = int(np.prod(x.shape[2:]))
    out = np.empty_like(x)
    lib().mpo_normalize_descriptors(_p(x, _f32p), B, D, HW, _p(out, _f32p))
    return out


# --- row 4: utils.box_nms (multipoint/utils/utils.py:78-122) ---
def box_nms(prob, size, min_prob, iou=0.1, keep_top_k=0, literal=False):
    prob = _f32(prob)
    if prob.ndim not in (2, 4):
        raise ValueError('The probability must be either 2D (H,W), or 4D (B, 1, H, W)')
    H, W = prob.shape[-2:]
    B = 1 if prob.ndim == 2 else prob.shape[0]
    out = np.empty_like(prob)
    fn = lib().mpo_box_nms_literal if literal else lib().mpo_box_nms
    rc = fn(_p(prob, _f32p), B, H, W, ctypes.c_double(size), ctypes.c_double(min_prob),
            ctypes.c_double(iou), int(keep_top_k), _p(out, _f32p))
    assert rc == 0
    return out


def nms_footprint(size, iou=0.1):
    R = max(1, int(np.ceil(size)))
    fp = np.zeros((2 * R + 1, 2 * R + 1), np.uint8)
    lib().mpo_nms_footprint(ctypes.c_double(size), ctypes.c_double(iou), R, _p(fp, _u8p))
    return fp


# --- row 4b: torch.nonzero((p > thr).float()) idiom (predict_align_image_pair.py:170-171) ---
def extract_keypoints(prob, thr):
    prob = _f32(prob)
    H, W = prob.shape
    kp = np.empty((H * W, 2), np.int64)
    n = lib().mpo_extract_keypoints(_p(prob, _f32p), H, W, ctypes.c_double(thr), _p(kp, _i64p),
                                    ctypes.c_int64(H * W))
    return kp[:n].copy()


# --- row 5: utils.interpolate_descriptors (utils.py:159-167) ---
def interpolate_descriptors(keypoints, desc, H, W):
    kp = np.ascontiguousarray(keypoints, dtype=np.int64).reshape(-1, 2)
    desc = _f32(desc)
    D, Hc, Wc = desc.shape
    out = np.empty((kp.shape[0], D), np.float32)
    lib().mpo_interpolate_descriptors(_p(kp, _i64p), ctypes.c_int64(kp.shape[0]), _p(desc, _f32p), D,
                                      Hc, Wc, int(H), int(W), _p(out, _f32p))
    return out


# --- rows 6-8: matching (multipoint/utils/matching.py) ---
_MODE = {'nn': 0, 'bf': 1}


def nearest(d1, d2, mode='nn', f64=False):
    d1, d2 = _f32(d1), _f32(d2)
    N1, D = d1.shape
    N2 = d2.shape[0]
    i12, i21 = np.empty(N1, np.int32), np.empty(N2, np.int32)
    b12, s12 = np.empty(N1), np.empty(N1)
    b21, s21 = np.empty(N2), np.empty(N2)
    lib().mpo_nearest(_p(d1, _f32p), N1, _p(d2, _f32p), N2, D, _MODE[mode], int(f64),
                      _p(i12, _i32p), _p(b12, _f64p), _p(s12, _f64p),
                      _p(i21, _i32p), _p(b21, _f64p), _p(s21, _f64p))
    return dict(idx12=i12, best12=b12, second12=s12, idx21=i21, best21=b21, second21=s21)


def match_mutual(d1, d2, mode='nn', f64=False, cross_check=True, threshold=-1.0):
    """NNMatcher.match (matching.py:41-72) for mode='nn' with threshold>=0;
    cv2.BFMatcher(NORM_L2, crossCheck=..).match (matching.py:7,31) for mode='bf'."""
    d1, d2 = _f32(d1), _f32(d2)
    N1, D = d1.shape if d1.ndim == 2 else (0, 0)
    N2 = d2.shape[0]
    q, t, dist = np.empty(N1, np.int32), np.empty(N1, np.int32), np.empty(N1, np.float32)
    n = lib().mpo_match_mutual(_p(d1, _f32p), N1, _p(d2, _f32p), N2, D, _MODE[mode], int(f64),
                               int(cross_check), ctypes.c_double(threshold), _p(q, _i32p),
                               _p(t, _i32p), _p(dist, _f32p))
    return q[:n].copy(), t[:n].copy(), dist[:n].copy()


def match_ratio(d1, d2, mode='bf', f64=False, ratio=0.9):
    """get_matches(knn_matches=True) (matching.py:21-28)."""
    d1, d2 = _f32(d1), _f32(d2)
    N1, D = d1.shape
    N2 = d2.shape[0]
    q, t, dist = np.empty(N1, np.int32), np.empty(N1, np.int32), np.empty(N1, np.float32)
    n = lib().mpo_match_ratio(_p(d1, _f32p), N1, _p(d2, _f32p), N2, D, _MODE[mode], int(f64),
                              ctypes.c_double(ratio), _p(q, _i32p), _p(t, _i32p), _p(dist, _f32p))
    return q[:n].copy(), t[:n].copy(), dist[:n].copy()


def match_threshold(d1, d2, threshold=0.4, f64=False):
    """ThresholdMatcher.match (matching.py:74-99)."""
    d1, d2 = _f32(d1), _f32(d2)
    N1, D = d1.shape
    N2 = d2.shape[0]
    cap = N1 * N2
    q, t, dist = np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap, np.float32)
    n = lib().mpo_match_threshold(_p(d1, _f32p), N1, _p(d2, _f32p), N2, D, int(f64),
                                  ctypes.c_double(threshold), _p(q, _i32p), _p(t, _i32p),
                                  _p(dist, _f32p), ctypes.c_int64(cap))
    return q[:n].copy(), t[:n].copy(), dist[:n].copy()


# --- rows 9-10: warp_perspective_tensor (multipoint/utils/homographies.py:404-425) ---
def linspace_table(n):
    """torch.linspace(-1, 1, n) in fp32 as ATen computes it: step=(end-start)/(n-1);
    start + i*step for the lower half, end - (n-1-i)*step for the upper half."""
    if n == 1:
        return np.array([-1.0], np.float32)
    step = np.float32(np.float32(2.0) / np.float32(n - 1))
    i = np.arange(n)
    lo = np.float32(-1.0) + step * i.astype(np.float32)
    hi = np.float32(1.0) - step * (n - 1 - i).astype(np.float32)
    return np.where(i < n // 2, lo, hi).astype(np.float32)


def normal_transform_pixel(h, w):
    """kornia normal_transform_pixel: pixel -> [-1,1] with (size-1) denominators."""
    return np.array([[2.0 / (w - 1), 0.0, -1.0], [0.0, 2.0 / (h - 1), -1.0], [0.0, 0.0, 1.0]], np.float32)


def warp_matrix(M, H, W):
    """The 3x3 fp32 matrix A (dst_norm -> src_norm) that warp_perspective_tensor hands to
    grid_sample for a pixel-space homography M (src -> dst), homographies.py:424-425:
    M_norm = N @ (M @ N^-1); A = inverse(M_norm)."""
    N = normal_transform_pixel(H, W)
    Ninv = np.linalg.inv(N).astype(np.float32)
    Mn = (N @ (np.asarray(M, np.float32) @ Ninv)).astype(np.float32)
    return np.linalg.inv(Mn).astype(np.float32)


def warp(src, A, mode='bilinear', padding='zeros'):
    src = _f32(src)
    H, W = src.shape[-2:]
    N = int(np.prod(src.shape[:-2])) if src.ndim > 2 else 1
    out = np.empty_like(src)
    A = _f32(A).reshape(9)
    xs, ys = linspace_table(W), linspace_table(H)
    lib().mpo_warp(_p(src, _f32p), N, H, W, _p(A, _f32p), _p(xs, _f32p), _p(ys, _f32p),
                   {'bilinear': 0, 'nearest': 1}[mode], {'zeros': 0, 'reflection': 1}[padding],
                   _p(out, _f32p))
    return out


def ha_aggregate(prob0, probw_a, probw_b, masks, Ainv, aggregation, min_count):
    prob0, probw_a, masks, Ainv = _f32(prob0), _f32(probw_a), _f32(masks), _f32(Ainv)
    probw_b = _f32(probw_b) if probw_b is not None else None
    n, B, H, W = probw_a.shape
    out = np.empty((B, H, W), np.float32)
    count = np.empty((B, H, W), np.float32)
    xs, ys = linspace_table(W), linspace_table(H)
    lib().mpo_ha_aggregate(_p(prob0, _f32p), _p(probw_a, _f32p), _p(probw_b, _f32p), _p(masks, _f32p),
                           _p(Ainv, _f32p), n, B, H, W, _p(xs, _f32p), _p(ys, _f32p),
                           {'none': 0, 'prod': 1, 'sum': 2}[aggregation], int(min_count),
                           _p(out, _f32p), _p(count, _f32p))
    return out, count


def homographic_adaptation(images, net_prob, homographies, masks, min_count=2, images_b=None,
                           aggregation='none', A_warp=None, A_unwarp=None):
    """homographic_adaptation (homographies.py:130-189) and, with images_b, the multispectral
    variant (:38-128), for pre-sampled pixel-space homographies (n,3,3) and valid masks (n,H,W).
    ``net_prob(batch_of_images, spectrum) -> (B,H,W)`` heatmaps.  filter_size=0 only.
    A_warp / A_unwarp (n,3,3) override the normalised matrices (to separate the per-pixel
    arithmetic from how a given LAPACK rounds the 3x3 algebra)."""
    images = _f32(images)
    B, H, W = images.shape
    n = len(homographies)
    pa0 = net_prob(images, 0)
    if images_b is not None:
        pb0 = net_prob(_f32(images_b), 1)
        prob0 = pa0 * pb0 if aggregation == 'prod' else pa0 + pb0
    else:
        prob0 = pa0
    pw_a = np.empty((n, B, H, W), np.float32)
    pw_b = np.empty((n, B, H, W), np.float32) if images_b is not None else None
    Ainv = np.empty((n, 3, 3), np.float32)
    for i in range(n):
        Mf = np.asarray(homographies[i], np.float32)
        A = warp_matrix(Mf, H, W) if A_warp is None else _f32(A_warp[i])
        pw_a[i] = net_prob(warp(images, A, 'bilinear', 'reflection'), 0)
        if images_b is not None:
            pw_b[i] = net_prob(warp(_f32(images_b), A, 'bilinear', 'reflection'), 1)
        # torch.inverse(homography) on the fp32 matrix (homographies.py:112,180)
        Ainv[i] = warp_matrix(np.linalg.inv(Mf).astype(np.float32), H, W) if A_unwarp is None else A_unwarp[i]
    out, _ = ha_aggregate(prob0, pw_a, pw_b, masks, Ainv,
                          aggregation if images_b is not None else 'none', min_count)
    return out


# --- compute_valid_mask (homographies.py:375-402) ---
def invert3x3(M):
    M = np.ascontiguousarray(M, np.float64).reshape(9)
    out = np.empty(9, np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    lib().mpo_invert3x3(M.ctypes.data_as(dp), out.ctypes.data_as(dp))
    return out.reshape(3, 3)


def valid_mask(image_shape, homography, erosion_radius=0, mask_border=False, inverse=None):
    """-> (H,W) uint8; ``inverse`` overrides the restated cv::invert (pass cv2.invert's output to isolate the warp)."""
    H, W = int(image_shape[0]), int(image_shape[1])
    Minv = invert3x3(homography) if inverse is None else np.ascontiguousarray(inverse, np.float64)
    out = np.empty((H, W), np.uint8)
    lib().mpo_valid_mask(Minv.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), H, W, int(erosion_radius), int(bool(mask_border)),
                         out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    return out


# --- point geometry of the evaluation loops (evaluation.py:148-199, 253-358) ---
_i64p = ctypes.POINTER(ctypes.c_int64)


def warp_keypoints(kp, homography, as_int=True):
    """homographies.py:331-346 (cv2.perspectiveTransform on flipped float64 points); as_int -> .astype(int)."""
    kp = np.ascontiguousarray(kp, np.int64).reshape(-1, 2)
    m = np.ascontiguousarray(homography, np.float64).reshape(9)
    f = np.empty(kp.shape, np.float64)
    i = np.empty(kp.shape, np.int64)
    lib().mpo_warp_keypoints(kp.ctypes.data_as(_i64p), len(kp), m.ctypes.data_as(_f64p), f.ctypes.data_as(_f64p), i.ctypes.data_as(_i64p))
    return i if as_int else f


def points_min_dist2(q, t, H, W):
    q = np.ascontiguousarray(q, np.int64).reshape(-1, 2)
    t = np.ascontiguousarray(t, np.int64).reshape(-1, 2)
    out = np.empty(len(q), np.int64)
    lib().mpo_points_min_dist2(q.ctypes.data_as(_i64p), len(q), t.ctypes.data_as(_i64p), len(t), int(H), int(W), out.ctypes.data_as(_i64p))
    return out


def points_correct(qw, t, thr, mq=None, mt=None):
    qw = np.ascontiguousarray(qw, np.float64).reshape(-1, 2)
    t = np.ascontiguousarray(t, np.int64).reshape(-1, 2)
    mq = np.ascontiguousarray(mq if mq is not None else [], np.int32)
    mt = np.ascontiguousarray(mt if mt is not None else [], np.int32)
    row_any = np.empty(len(qw), np.uint8)
    tp = np.empty(len(mq), np.uint8)
    i32 = ctypes.POINTER(ctypes.c_int32)
    lib().mpo_points_correct(qw.ctypes.data_as(_f64p), len(qw), t.ctypes.data_as(_i64p), len(t), ctypes.c_float(thr),
                             row_any.ctypes.data_as(_u8p), mq.ctypes.data_as(i32), mt.ctypes.data_as(i32), len(mq), tp.ctypes.data_as(_u8p))
    return row_any, tp


# --- row 3 glue: ReLU / BatchNorm(eval) / MaxPool / pad between the backbone's convolutions (MultiPoint.py:61-90) ---
def relu_bn_pad(x, mean, var, eps, weight, bias, conv_bias=None, bn_first=False, pool=False, pad=1, reflect=True):
    x = _f32(x)
    B, C, H, W = x.shape
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    out = np.empty((B, C, Ho + 2 * pad, Wo + 2 * pad), np.float32)
    cb = None if conv_bias is None else _f32(conv_bias)
    mean, var, weight, bias = _f32(mean), _f32(var), _f32(weight), _f32(bias)
    lib().mpo_relu_bn_pad(_p(x, _f32p), B, C, H, W, None if cb is None else _p(cb, _f32p), _p(mean, _f32p), _p(var, _f32p),
                          ctypes.c_float(eps), _p(weight, _f32p), _p(bias, _f32p), int(bn_first), int(pool), int(pad), int(reflect),
                          _p(out, _f32p))
    return out


def conv1_relu_bn_pad(img, w, conv_bias, mean, var, eps, weight, bias, bn_first=False, in_reflect=True, pad=1, out_reflect=True):
    img = _f32(img)
    B, _, H, W = img.shape
    w = _f32(w)
    C = w.shape[0]
    out = np.empty((B, C, H + 2 * pad, W + 2 * pad), np.float32)
    cb = None if conv_bias is None else _f32(conv_bias)
    mean, var, weight, bias = _f32(mean), _f32(var), _f32(weight), _f32(bias)
    lib().mpo_conv1_relu_bn_pad(_p(img, _f32p), B, H, W, _p(w, _f32p), None if cb is None else _p(cb, _f32p), C, _p(mean, _f32p),
                                _p(var, _f32p), ctypes.c_float(eps), _p(weight, _f32p), _p(bias, _f32p), int(bn_first), int(in_reflect),
                                int(pad), int(out_reflect), _p(out, _f32p))
    return out
