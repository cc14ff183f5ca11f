"""Device-tensor front end of the C ABI: one function per entry point of multipoint_b200.h.

torch is plumbing only (device memory, current stream).  Every function takes CUDA tensors and
raises if handed anything else: there is no CPU or eager fallback on the product path.
"""
import ctypes

import numpy as np
import torch

from . import _lib

METRIC = {'nn': 0, 'l2': 1}
ALGO = {'tensor': 0, 'simt': 1}
AGG = {None: 0, 'none': 0, 'prod': 1, 'sum': 2}
HA_INIT, HA_FINISH, HA_STAGED = 1, 2, 4


def _cuda(t, dtype, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("multipoint_b200.ops: %s must be a CUDA tensor (no CPU fallback)" % name)
    if t.dtype != dtype:
        raise TypeError("multipoint_b200.ops: %s must be %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def detector_head(logits, valid_mask=None):
    """(B,65,Hc,Wc) fp32 logits -> (B,1,8Hc,8Wc) heatmap; optional (B,1,H,W) valid mask.  The mask is a BINARY mask
    (the reference multiplies prob by a bool valid_mask): any non-zero entry, of any dtype, keeps the pixel -- the same
    rule extract_keypoints applies."""
    logits = _cuda(logits, torch.float32, "logits")
    if logits.dim() != 4 or logits.shape[1] != 65:
        raise ValueError("logits must be (B,65,Hc,Wc), got %s" % (tuple(logits.shape),))
    B, _, Hc, Wc = logits.shape
    prob = torch.empty((B, 1, Hc * 8, Wc * 8), dtype=torch.float32, device=logits.device)
    mask = None
    if valid_mask is not None:
        if not valid_mask.is_cuda:
            raise RuntimeError("valid_mask must be a CUDA tensor")
        mask = (valid_mask != 0).to(torch.uint8).contiguous()   # 0/1 bytes whatever the dtype (a 0/255 byte mask must not scale by 255)
        if mask.numel() != prob.numel():
            raise ValueError("valid_mask must have %d elements" % prob.numel())
    with torch.cuda.device(logits.device):
        _lib.check(_lib.load().mp_detector_head_f32(_ptr(logits), B, Hc, Wc, _ptr(mask), _ptr(prob), _stream(logits)),
                   "mp_detector_head_f32")
    return prob


def heatmap_magicleap(semi):
    """SuperPointMagicLeap.generate_heatmap: (B,65,Hc,Wc) -> (B,1,8Hc,8Wc), exp(x)/(sum exp + 1e-5), no dustbin."""
    semi = _cuda(semi, torch.float32, "semi")
    if semi.dim() != 4 or semi.shape[1] != 65:
        raise ValueError("semi must be (B,65,Hc,Wc), got %s" % (tuple(semi.shape),))
    B, _, Hc, Wc = semi.shape
    prob = torch.empty((B, 1, Hc * 8, Wc * 8), dtype=torch.float32, device=semi.device)
    with torch.cuda.device(semi.device):
        _lib.check(_lib.load().mp_heatmap_magicleap_f32(_ptr(semi), B, Hc, Wc, _ptr(prob), _stream(semi)),
                   "mp_heatmap_magicleap_f32")
    return prob


def depth_to_space(x, block_size):
    x = _cuda(x, torch.float32, "x")
    N, C, H, W = x.shape
    bs = int(block_size)
    if C % (bs * bs) != 0:
        raise ValueError("channels %d not divisible by block_size^2" % C)
    out = torch.empty((N, C // (bs * bs), H * bs, W * bs), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mp_depth_to_space_f32(_ptr(x), N, C // (bs * bs), H, W, bs, _ptr(out), _stream(x)),
                   "mp_depth_to_space_f32")
    return out


def normalize_descriptors(x, nchw=True, nhwc=False):
    """F.normalize(x, p=2, dim=1) of (B,D,Hc,Wc); returns (nchw, nhwc) with None for the one not requested."""
    x = _cuda(x, torch.float32, "x")
    B, D = x.shape[:2]
    HW = int(np.prod(x.shape[2:]))
    o1 = torch.empty_like(x) if nchw else None
    o2 = torch.empty((B,) + tuple(x.shape[2:]) + (D,), dtype=torch.float32, device=x.device) if nhwc else None
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().mp_normalize_descriptors_f32(_ptr(x), B, D, HW, _ptr(o1), _ptr(o2), _stream(x)),
                   "mp_normalize_descriptors_f32")
    return o1, o2


def box_nms(prob, size, min_prob, iou=0.1, keep_top_k=0, want_keypoints=False, kp_cap=None, want_dense=True):
    """prob (B,H,W) fp32 -> dense NMS map (B,H,W) [+ keypoints (B,cap,2) int64, scores (B,cap), counts (B)].
    ``want_dense=False`` (with want_keypoints) skips the dense map -- it is returned as None: on the sparse top-k path
    nothing of its size is then written at all."""
    prob = _cuda(prob, torch.float32, "prob")
    B, H, W = prob.shape
    dev = prob.device
    if not want_dense and not want_keypoints:
        raise ValueError("box_nms: nothing requested (want_dense=False needs want_keypoints=True)")
    out = torch.empty_like(prob) if want_dense else None
    lib = _lib.load()
    kp = sc = cnt = None
    cap = 0
    if want_keypoints:
        cap = int(kp_cap) if kp_cap is not None else (int(keep_top_k) if keep_top_k > 0 else H * W)
        kp = torch.zeros((B, cap, 2), dtype=torch.int64, device=dev)
        sc = torch.zeros((B, cap), dtype=torch.float32, device=dev)
        cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
    nbytes = lib.mp_box_nms_workspace_bytes(B, H, W)
    ws = _ws(nbytes, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mp_box_nms_f32(_ptr(prob), B, H, W, float(size), float(min_prob), float(iou), int(keep_top_k),
                                      _ptr(out), _ptr(kp), _ptr(sc), _ptr(cnt), cap, _ptr(ws), nbytes, _stream(prob)),
                   "mp_box_nms_f32")
    return (out, kp, sc, cnt) if want_keypoints else out


def extract_keypoints(prob, threshold, mask=None, kp_cap=None):
    """torch.nonzero((prob > thr).float() [* mask]) per image: keypoints (B,cap,2) int64, scores, counts."""
    prob = _cuda(prob, torch.float32, "prob")
    B, H, W = prob.shape
    dev = prob.device
    cap = int(kp_cap) if kp_cap is not None else H * W
    m = None
    if mask is not None:
        m = (mask != 0).to(torch.uint8).contiguous()
    kp = torch.zeros((B, cap, 2), dtype=torch.int64, device=dev)
    sc = torch.zeros((B, cap), dtype=torch.float32, device=dev)
    cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    nbytes = lib.mp_extract_keypoints_workspace_bytes(B, H, W)
    ws = _ws(nbytes, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mp_extract_keypoints_f32(_ptr(prob), _ptr(m), B, H, W, float(threshold), _ptr(kp), _ptr(sc),
                                                _ptr(cnt), cap, _ptr(ws), nbytes, _stream(prob)),
                   "mp_extract_keypoints_f32")
    return kp, sc, cnt


def transpose_descriptors(desc):
    """(B,D,Hc,Wc) -> channels-last (B,Hc,Wc,D) copy (mp_transpose_descriptors_f32)."""
    desc = _cuda(desc, torch.float32, "desc")
    B, D, Hc, Wc = desc.shape
    out = torch.empty((B, Hc, Wc, D), dtype=torch.float32, device=desc.device)
    with torch.cuda.device(desc.device):
        _lib.check(_lib.load().mp_transpose_descriptors_f32(_ptr(desc), B, D, Hc * Wc, _ptr(out), _stream(desc)),
                   "mp_transpose_descriptors_f32")
    return out


def sample_descriptors(keypoints, desc, H, W, counts=None, channels_last=False, transpose=None, split=False):
    """keypoints (B,K,2) int64 (y,x); desc (B,D,Hc,Wc) or, channels_last, (B,Hc,Wc,D) -> (B,K,D) unit rows.
    An NCHW map is copied to channels-last first when the gather is large enough to pay for it (``transpose=None``:
    K*16 >= Hc*Wc, i.e. the four corners of all keypoints touch at least as many descriptor rows as the copy moves;
    True / False force it): the strided NCHW gather moves 4x the sectors of the contiguous-row one (measured 1.02 ms
    against 0.17 ms at 128 x 2048 keypoints, D = 256).  Results are bit-identical either way.
    ``split=True`` also returns the rows in the matcher's operand form -- a dict with 'hi', 'mid' (B,K,D) bf16,
    'sq_norms' (B,K) and 'max_norm' (B) int32 float bits -- for ``match(..., split1=, split2=)``."""
    keypoints = _cuda(keypoints, torch.int64, "keypoints")
    desc = _cuda(desc, torch.float32, "desc")
    B, K = keypoints.shape[:2]
    if not channels_last:
        if transpose is None:
            transpose = desc.shape[1] in (64, 128, 256) and K * 16 >= desc.shape[2] * desc.shape[3]
        if transpose:
            desc, channels_last = transpose_descriptors(desc), True
    if channels_last:
        _, Hc, Wc, D = desc.shape
    else:
        _, D, Hc, Wc = desc.shape
    if counts is not None:
        counts = _cuda(counts, torch.int32, "counts")
    out = torch.empty((B, K, D), dtype=torch.float32, device=desc.device)
    if split:
        if not channels_last or D not in (64, 128, 256):
            raise NotImplementedError("sample_descriptors(split=True) needs a channels-last map with D in {64, 128, 256}")
        sp = {'hi': torch.empty((B, K, D), dtype=torch.bfloat16, device=desc.device),
              'mid': torch.empty((B, K, D), dtype=torch.bfloat16, device=desc.device),
              'sq_norms': torch.empty((B, K), dtype=torch.float32, device=desc.device),
              # the rows are unit-norm by construction (or zero): an upper bound of the largest norm, as float bits
              'max_norm': torch.full((B,), 0x3F800008, dtype=torch.int32, device=desc.device)}   # 1.000001f
        with torch.cuda.device(desc.device):
            _lib.check(_lib.load().mp_sample_descriptors_split_f32(_ptr(keypoints), _ptr(counts), B, K, _ptr(desc), D, Hc, Wc, 1,
                                                                   int(H), int(W), _ptr(out), _ptr(sp['hi']), _ptr(sp['mid']),
                                                                   _ptr(sp['sq_norms']), _stream(desc)),
                       "mp_sample_descriptors_split_f32")
        return out, sp
    with torch.cuda.device(desc.device):
        _lib.check(_lib.load().mp_sample_descriptors_f32(_ptr(keypoints), _ptr(counts), B, K, _ptr(desc), D, Hc, Wc,
                                                         1 if channels_last else 0, int(H), int(W), _ptr(out),
                                                         _stream(desc)), "mp_sample_descriptors_f32")
    return out


def _match_args(d1, d2, n1, n2):
    d1 = _cuda(d1, torch.float32, "desc_1")
    d2 = _cuda(d2, torch.float32, "desc_2")
    if d1.dim() == 2:
        d1, d2 = d1[None], d2[None]
    P, N1, D = d1.shape
    if d2.shape[0] != P or d2.shape[2] != D:
        raise ValueError("descriptor sets disagree: %s vs %s" % (tuple(d1.shape), tuple(d2.shape)))
    if n1 is not None:
        n1 = _cuda(n1, torch.int32, "n1")
    if n2 is not None:
        n2 = _cuda(n2, torch.int32, "n2")
    return d1.contiguous(), d2.contiguous(), n1, n2, P, N1, d2.shape[1], D


def default_algo(D):
    return 'tensor' if (D % 64 == 0 and D <= 256) else 'simt'


def nearest(d1, d2, metric='nn', algo=None, n1=None, n2=None, want_scores=True):
    """Exact nearest neighbours both ways.  Returns dict idx12 (P,N1), idx21 (P,N2) [+ best/second sims]."""
    d1, d2, n1, n2, P, N1, N2, D = _match_args(d1, d2, n1, n2)
    dev = d1.device
    algo = algo or default_algo(D)
    i12 = torch.empty((P, N1), dtype=torch.int32, device=dev)
    i21 = torch.empty((P, N2), dtype=torch.int32, device=dev)
    outs = [torch.empty((P, n), dtype=torch.float32, device=dev) if want_scores else None for n in (N1, N1, N2, N2)]
    lib = _lib.load()
    nbytes = lib.mp_match_workspace_bytes(P, N1, N2, D)
    ws = _ws(nbytes, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mp_nearest_f32(_ptr(d1), _ptr(n1), N1, _ptr(d2), _ptr(n2), N2, P, D, METRIC[metric], ALGO[algo],
                                      _ptr(i12), _ptr(outs[0]), _ptr(outs[1]), _ptr(i21), _ptr(outs[2]), _ptr(outs[3]),
                                      _ptr(ws), nbytes, _stream(d1)), "mp_nearest_f32")
    return dict(idx12=i12, idx21=i21, best12=outs[0], second12=outs[1], best21=outs[2], second21=outs[3])


def match(d1, d2, metric='l2', algo=None, kind='mutual', cross_check=True, threshold=-1.0, ratio=0.9, n1=None, n2=None,
          split1=None, split2=None):
    """Match lists for P pairs: query (P,N1), train (P,N1) int32, dist (P,N1) fp32, counts (P) int32.
    split1 / split2: the operand dicts of ``sample_descriptors(split=True)`` for d1 / d2 (tensor path; no prep pass)."""
    d1, d2, n1, n2, P, N1, N2, D = _match_args(d1, d2, n1, n2)
    dev = d1.device
    algo = algo or default_algo(D)
    q = torch.empty((P, N1), dtype=torch.int32, device=dev)
    t = torch.empty((P, N1), dtype=torch.int32, device=dev)
    dist = torch.empty((P, N1), dtype=torch.float32, device=dev)
    cnt = torch.zeros((P,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    nbytes = lib.mp_match_workspace_bytes(P, N1, N2, D)
    ws = _ws(nbytes, dev)
    if split1 is not None and split2 is not None and ALGO[algo] == 0:
        for sp, N in ((split1, N1), (split2, N2)):
            if tuple(sp['hi'].shape) != (P, N, D) or not (sp['hi'].is_contiguous() and sp['mid'].is_contiguous() and sp['sq_norms'].is_contiguous()):
                raise ValueError("match: split operands must be contiguous (P,N,D) planes of the same descriptor sets")
        with torch.cuda.device(dev):
            _lib.check(lib.mp_match_split_f32(_ptr(d1), _ptr(split1['hi']), _ptr(split1['mid']), _ptr(split1['sq_norms']), _ptr(split1['max_norm']),
                                              _ptr(n1), N1, _ptr(d2), _ptr(split2['hi']), _ptr(split2['mid']), _ptr(split2['sq_norms']),
                                              _ptr(split2['max_norm']), _ptr(n2), N2, P, D, METRIC[metric],
                                              {'mutual': 0, 'ratio': 1}[kind], int(bool(cross_check)), float(threshold), float(ratio),
                                              _ptr(q), _ptr(t), _ptr(dist), _ptr(cnt), _ptr(ws), nbytes, _stream(d1)),
                       "mp_match_split_f32")
        return q, t, dist, cnt
    with torch.cuda.device(dev):
        _lib.check(lib.mp_match_f32(_ptr(d1), _ptr(n1), N1, _ptr(d2), _ptr(n2), N2, P, D, METRIC[metric], ALGO[algo],
                                    {'mutual': 0, 'ratio': 1}[kind], int(bool(cross_check)), float(threshold), float(ratio),
                                    _ptr(q), _ptr(t), _ptr(dist), _ptr(cnt), _ptr(ws), nbytes, _stream(d1)),
                   "mp_match_f32")
    return q, t, dist, cnt


def match_threshold(d1, d2, threshold):
    """ThresholdMatcher: all pairs with sqrt(2-2clip(a.b)) < threshold, row-major.  Synchronises once."""
    d1 = _cuda(d1, torch.float32, "desc_1")
    d2 = _cuda(d2, torch.float32, "desc_2")
    N1, D = d1.shape
    N2 = d2.shape[0]
    dev = d1.device
    lib = _lib.load()
    nbytes = 16 * (N1 + 1) + 512
    ws = _ws(nbytes, dev)
    total = ctypes.c_int64(0)
    with torch.cuda.device(dev):
        # pass 1: count only (cap = 0)
        _lib.check(lib.mp_match_threshold_f32(_ptr(d1), N1, _ptr(d2), N2, D, float(threshold), None, None, None, 0,
                                              ctypes.byref(total), _ptr(ws), nbytes, _stream(d1)), "mp_match_threshold_f32")
        n = int(total.value)
        q = torch.empty((n,), dtype=torch.int32, device=dev)
        t = torch.empty((n,), dtype=torch.int32, device=dev)
        dist = torch.empty((n,), dtype=torch.float32, device=dev)
        if n:
            _lib.check(lib.mp_match_threshold_f32(_ptr(d1), N1, _ptr(d2), N2, D, float(threshold), _ptr(q), _ptr(t),
                                                  _ptr(dist), n, ctypes.byref(total), _ptr(ws), nbytes, _stream(d1)),
                       "mp_match_threshold_f32")
    return q, t, dist


def _linspace_f32(n):
    """linspace(-1, 1, n) in fp32 by ATen's scalar formula (one rounding per operation):
    step = 2/(n-1); -1 + step*i below the midpoint, 1 - step*(n-1-i) from it on.  torch.linspace on
    the CPU is vectorised and rounds differently per SIMD width, so the table is built here to be
    the same bits on every host."""
    if n == 1:
        return np.array([-1.0], np.float32)
    step = np.float32(np.float32(2.0) / np.float32(n - 1))
    i = np.arange(n)
    lo = np.float32(-1.0) + step * i.astype(np.float32)
    hi = np.float32(1.0) - step * (n - 1 - i).astype(np.float32)
    return np.where(i < n // 2, lo, hi).astype(np.float32)


def linspace_tables(H, W, device):
    """Normalised destination grid of kornia's create_meshgrid: (xs (W), ys (H)) device tensors."""
    return torch.from_numpy(_linspace_f32(W)).to(device), torch.from_numpy(_linspace_f32(H)).to(device)


def warp(src, A, mode='bilinear', padding='zeros', tables=None, groups=1):
    """src (N,H,W) planes shared by all matrices; A (n,3,3) normalised dst->src; out (n,N,H,W).
    ``groups`` > 1 splits the planes into that many equal groups, each with an output block of its own:
    out (groups, n, N // groups, H, W) -- one launch (one pass of coordinate arithmetic) for both spectra of a pair."""
    src = _cuda(src, torch.float32, "src")
    A = _cuda(A, torch.float32, "A")
    N, H, W = src.shape
    n = A.shape[0]
    if groups < 1 or N % groups:
        raise ValueError("warp: {} planes do not split into {} groups".format(N, groups))
    G = N // groups if N else 1
    xs, ys = tables if tables is not None else linspace_tables(H, W, src.device)
    out = torch.empty((n, N, H, W) if groups == 1 else (groups, n, G, H, W), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        _lib.check(_lib.load().mp_warp_groups_f32(_ptr(src), N, max(G, 1), n, H, W, _ptr(A), _ptr(xs), _ptr(ys),
                                                  {'bilinear': 0, 'nearest': 1}[mode], {'zeros': 0, 'reflection': 1}[padding],
                                                  _ptr(out), _stream(src)), "mp_warp_groups_f32")
    return out


def valid_masks(Minv, H, W, erosion_radius=0, mask_border=False, device=None):
    """(n,3,3) float64 INVERTED homographies (as cv2.warpPerspective inverts them) -> (n,H,W) uint8 valid masks."""
    if not torch.is_tensor(Minv):
        Minv = torch.from_numpy(np.ascontiguousarray(np.asarray(Minv, np.float64).reshape(-1, 3, 3)))
    if not Minv.is_cuda:
        if device is None:
            raise RuntimeError("valid_masks needs a CUDA tensor or an explicit device")
        Minv = Minv.to(device)
    Minv = _cuda(Minv, torch.float64, "Minv").reshape(-1, 3, 3).contiguous()
    n = Minv.shape[0]
    out = torch.empty((n, int(H), int(W)), dtype=torch.uint8, device=Minv.device)
    with torch.cuda.device(Minv.device):
        _lib.check(_lib.load().mp_valid_mask_u8(_ptr(Minv), n, int(H), int(W), int(erosion_radius), int(bool(mask_border)),
                                                _ptr(out), _stream(Minv)), "mp_valid_mask_u8")
    return out


def ha_aggregate(prob0, probw_a, probw_b, masks, Ainv, aggregation, min_count, init=True, finish=True,
                 prob_acc=None, count_acc=None, tables=None, staged=False):
    """Unwarp + accumulate + finish of homographic adaptation; see mp_ha_aggregate_f32.  ``staged=True`` selects the
    TMA-staged kernel (bit-identical, measured slower at 512x640)."""
    probw_a = _cuda(probw_a, torch.float32, "probw_a")
    n, B, H, W = probw_a.shape
    dev = probw_a.device
    if probw_b is not None:
        probw_b = _cuda(probw_b, torch.float32, "probw_b")
    masks = _cuda(masks, torch.uint8, "masks")
    Ainv = _cuda(Ainv, torch.float32, "Ainv")
    if prob0 is not None:
        prob0 = _cuda(prob0, torch.float32, "prob0")
    xs, ys = tables if tables is not None else linspace_tables(H, W, dev)
    out = torch.empty((B, H, W), dtype=torch.float32, device=dev) if finish else None
    if not finish or not init:
        if prob_acc is None:
            prob_acc = torch.zeros((B, H, W), dtype=torch.float32, device=dev)
            count_acc = torch.zeros((B, H, W), dtype=torch.float32, device=dev)
    flags = (HA_INIT if init else 0) | (HA_FINISH if finish else 0) | (HA_STAGED if staged else 0)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().mp_ha_aggregate_f32(_ptr(prob0), _ptr(probw_a), _ptr(probw_b), _ptr(masks), _ptr(Ainv), n, B,
                                                   H, W, _ptr(xs), _ptr(ys), AGG[aggregation], int(min_count), flags,
                                                   _ptr(prob_acc), _ptr(count_acc), _ptr(out), _stream(probw_a)),
                   "mp_ha_aggregate_f32")
    return out if finish else (prob_acc, count_acc)


# ----------------------------------------------------------------------------- evaluation point geometry
def _counts(c, P, name):
    if c is None:
        return None
    c = _cuda(c, torch.int32, name)
    if c.numel() != P:
        raise ValueError("%s must have %d entries" % (name, P))
    return c


def warp_keypoints(kp, homographies, counts=None, as_int=True):
    """kp (P,cap,2) int64 (y,x), homographies (P,3,3) float64 -> warped (P,cap,2): int64 (truncated, like the
    reference's default) or float64."""
    kp = _cuda(kp, torch.int64, "kp")
    P, cap = kp.shape[:2]
    Hm = _cuda(homographies, torch.float64, "homographies").reshape(P, 9)
    counts = _counts(counts, P, "counts")
    out = torch.zeros((P, cap, 2), dtype=torch.int64 if as_int else torch.float64, device=kp.device)
    with torch.cuda.device(kp.device):
        _lib.check(_lib.load().mp_warp_keypoints_i64(_ptr(kp), _ptr(counts), P, cap, _ptr(Hm), None if as_int else _ptr(out),
                                                     _ptr(out) if as_int else None, _stream(kp)), "mp_warp_keypoints_i64")
    return out


def points_min_dist2(q, t, H, W, nq=None, nt=None):
    """q (P,capq,2), t (P,capt,2) int64 -> (P,capq) int64: exact squared distance to the nearest target,
    -1 for queries outside the (H,W) frame, INT64_MAX without targets."""
    q = _cuda(q, torch.int64, "q")
    t = _cuda(t, torch.int64, "t")
    P, capq = q.shape[:2]
    capt = t.shape[1]
    out = torch.full((P, capq), -1, dtype=torch.int64, device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(_lib.load().mp_points_min_dist2_i64(_ptr(q), _ptr(_counts(nq, P, "nq")), capq, _ptr(t), _ptr(_counts(nt, P, "nt")),
                                                       capt, P, int(H), int(W), _ptr(out), _stream(q)), "mp_points_min_dist2_i64")
    return out


def points_correct(qw, t, threshold, nq=None, nt=None, mq=None, mt=None, nm=None):
    """qw (P,capq,2) float64 warped points, t (P,capt,2) int64 -> (row_any (P,capq) uint8, tp (P,capm) uint8 or None)."""
    qw = _cuda(qw, torch.float64, "qw")
    t = _cuda(t, torch.int64, "t")
    P, capq = qw.shape[:2]
    capt = t.shape[1]
    row_any = torch.zeros((P, capq), dtype=torch.uint8, device=qw.device)
    tp = None
    capm = 0
    if mq is not None:
        mq = _cuda(mq, torch.int32, "mq")
        mt = _cuda(mt, torch.int32, "mt")
        capm = mq.shape[1]
        tp = torch.zeros((P, capm), dtype=torch.uint8, device=qw.device)
    with torch.cuda.device(qw.device):
        _lib.check(_lib.load().mp_points_correct_f32(_ptr(qw), _ptr(_counts(nq, P, "nq")), capq, _ptr(t), _ptr(_counts(nt, P, "nt")), capt,
                                                     P, float(threshold), _ptr(row_any), _ptr(mq), _ptr(mt), _ptr(_counts(nm, P, "nm")),
                                                     capm, _ptr(tp), _stream(qw)), "mp_points_correct_f32")
    return row_any, tp


# ----------------------------------------------------------------------------- row 3: glue between the backbone convolutions
def relu_bn_pad(x, scale, shift, bn_first=False, pool=False, pad=1, reflect=True, conv_bias=None):
    """[+ conv_bias] -> ReLU -> eval BatchNorm (folded scale/shift) [-> MaxPool2d(2,2)] [-> pad 1] in one pass.
    x (B,C,H,W) fp32 -> (B,C,Ho+2*pad,Wo+2*pad)."""
    x = _cuda(x, torch.float32, "x")
    scale = _cuda(scale, torch.float32, "scale")
    shift = _cuda(shift, torch.float32, "shift")
    B, C, H, W = x.shape
    if scale.numel() != C or shift.numel() != C:
        raise ValueError("scale/shift must have %d entries" % C)
    if conv_bias is not None:
        conv_bias = _cuda(conv_bias, torch.float32, "conv_bias")
        if conv_bias.numel() != C:
            raise ValueError("conv_bias must have %d entries" % C)
    Ho, Wo = (H // 2, W // 2) if pool else (H, W)
    out = torch.empty((B, C, Ho + 2 * pad, Wo + 2 * pad), dtype=torch.float32, device=x.device)
    step = max(1, 65535 // C)   # the kernel takes B*C <= 65535 planes per launch
    lib = _lib.load()
    with torch.cuda.device(x.device):
        for b0 in range(0, B, step):
            nb = min(step, B - b0)
            _lib.check(lib.mp_relu_bn_pad_f32(_ptr(x[b0:b0 + nb]), nb, C, H, W, _ptr(conv_bias), _ptr(scale), _ptr(shift), int(bool(bn_first)), int(bool(pool)),
                                              int(pad), int(bool(reflect)), _ptr(out[b0:b0 + nb]), _stream(x)), "mp_relu_bn_pad_f32")
    return out


def conv1_relu_bn_pad(image, weight, conv_bias, scale, shift, bn_first=False, in_reflect=True, pad=1, out_reflect=True):
    """[pad 1] -> Conv2d(1 -> C, 3x3) -> +bias -> ReLU/BatchNorm(eval) [-> pad 1] for the encoders' first layer.
    image (B,1,H,W) fp32 -> (B,C,H+2*pad,W+2*pad)."""
    image = _cuda(image, torch.float32, "image")
    weight = _cuda(weight, torch.float32, "weight")
    scale = _cuda(scale, torch.float32, "scale")
    shift = _cuda(shift, torch.float32, "shift")
    B, cin, H, W = image.shape
    C = weight.shape[0]
    if cin != 1 or tuple(weight.shape[1:]) != (1, 3, 3):
        raise ValueError("conv1_relu_bn_pad needs a (B,1,H,W) image and (C,1,3,3) weights")
    if conv_bias is not None:
        conv_bias = _cuda(conv_bias, torch.float32, "conv_bias")
    out = torch.empty((B, C, H + 2 * pad, W + 2 * pad), dtype=torch.float32, device=image.device)
    with torch.cuda.device(image.device):
        _lib.check(_lib.load().mp_conv1_relu_bn_pad_f32(_ptr(image), B, H, W, _ptr(weight), _ptr(conv_bias), _ptr(scale), _ptr(shift), C,
                                                        int(bool(bn_first)), int(bool(in_reflect)), int(pad), int(bool(out_reflect)),
                                                        _ptr(out), _stream(image)), "mp_conv1_relu_bn_pad_f32")
    return out
