"""Randomised parity: the CUDA path against the CPU oracle on random SHAPES and parameters (odd widths, tiny and
ragged inputs, every box size, every descriptor width, every matcher mode), seeded, a few seconds by default.
MP_FUZZ_ITERS=<n> runs n cases per kernel instead (the soak recorded in profiles/r2_fuzz_soak.txt).
Bit-exact for maps, keypoints and match indices; floating point within the tolerance written at the assert."""
import os

import numpy as np
import pytest
import torch

from multipoint_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
ITERS = int(os.environ.get("MP_FUZZ_ITERS", "10"))


@pytest.fixture(scope="module")
def ops():
    from multipoint_b200 import ops as _ops
    return _ops


@pytest.fixture(scope="module")
def utils():
    from multipoint_b200 import utils as _utils
    return _utils


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def test_fuzz_detector_head(ops, oracle):
    rng = np.random.RandomState(1001)
    for it in range(ITERS):
        B, Hc, Wc = int(rng.randint(1, 4)), int(rng.randint(1, 40)), int(rng.randint(1, 50))
        lg = syn.logits(5000 + it, B, Hc, Wc, sigma=float(rng.choice([0.01, 1.0, 2.0, 6.0])), bias=float(rng.choice([0.0, 5.0])))
        got = ops.detector_head(cu(lg)).cpu().numpy()
        np.testing.assert_allclose(got, oracle.detector_head(lg), rtol=1e-5, atol=1e-9, err_msg=str((B, Hc, Wc)))


def test_fuzz_box_nms(utils, ops, oracle):
    rng = np.random.RandomState(1002)
    for it in range(ITERS):
        B = int(rng.randint(1, 4))
        H, W = int(rng.randint(1, 150)), int(rng.randint(1, 200))
        if rng.rand() < 0.5:
            W = (W + 3) // 4 * 4                      # the vectorised path
        size = float(rng.choice([2.5, 3, 4, 5, 8]))
        thr = float(rng.choice([0.001, 0.015, 0.05]))
        quant = float(rng.choice([0, 0, 0.25]))        # planted score ties
        hm = syn.heatmap(6000 + it, B, H, W, quant=(quant or None))
        topk = int(rng.choice([0, 0, 1, 7, 300, 2048]))
        tag = str((B, H, W, size, thr, quant, topk))
        want = oracle.box_nms(hm, size, thr, keep_top_k=topk)
        got = utils.box_nms(cu(hm), size, thr, keep_top_k=topk).cpu().numpy()
        np.testing.assert_array_equal(got, want, err_msg=tag)
        # the fused keypoint output: torch.nonzero order of the same map
        cap = topk if topk > 0 else H * W
        dense, kp, sc, cnt = ops.box_nms(cu(hm[:, 0]), size, thr, keep_top_k=topk, want_keypoints=True, kp_cap=cap)
        np.testing.assert_array_equal(dense.cpu().numpy(), want[:, 0], err_msg=tag)
        _, kp2, sc2, cnt2 = ops.box_nms(cu(hm[:, 0]), size, thr, keep_top_k=topk, want_keypoints=True, kp_cap=cap, want_dense=False)
        for b in range(B):
            ref = oracle.extract_keypoints(want[b, 0], 0.0)
            assert int(cnt[b]) == len(ref) == int(cnt2[b]), tag
            np.testing.assert_array_equal(kp[b, :len(ref)].cpu().numpy(), ref, err_msg=tag)
            np.testing.assert_array_equal(kp2[b, :len(ref)].cpu().numpy(), ref, err_msg=tag)
            np.testing.assert_array_equal(sc[b, :len(ref)].cpu().numpy(), want[b, 0][ref[:, 0], ref[:, 1]], err_msg=tag)
            np.testing.assert_array_equal(sc2[b, :len(ref)].cpu().numpy(), want[b, 0][ref[:, 0], ref[:, 1]], err_msg=tag)


def test_fuzz_descriptors(ops, oracle):
    rng = np.random.RandomState(1003)
    for it in range(ITERS):
        B, D = int(rng.randint(1, 4)), int(rng.choice([3, 32, 64, 100, 128, 256]))
        Hc, Wc = int(rng.randint(1, 20)), int(rng.randint(1, 24))
        H, W = 8 * Hc, 8 * Wc
        K = int(rng.randint(0, 70))
        raw = rng.randn(B, D, Hc, Wc).astype(np.float32)
        if rng.rand() < 0.3:
            raw[:, :, 0, 0] = 0.0                     # a zero descriptor: x / max(|x|, 1e-12)
        nchw, nhwc = ops.normalize_descriptors(cu(raw), nchw=True, nhwc=True)
        want = oracle.normalize_descriptors(raw)
        np.testing.assert_allclose(nchw.cpu().numpy(), want, rtol=1e-6, atol=1e-12)
        np.testing.assert_array_equal(nhwc.permute(0, 3, 1, 2).cpu().numpy(), nchw.cpu().numpy())
        kp = np.stack([rng.randint(0, H, (B, K)), rng.randint(0, W, (B, K))], axis=2).astype(np.int64)
        if K:
            kp[:, 0] = (0, 0)
            kp[:, -1] = (H - 1, W - 1)
        counts = rng.randint(0, K + 1, (B,)).astype(np.int32)
        dn = nchw.cpu().numpy()
        for layout, src in (("nchw", nchw), ("nhwc", nhwc)):
            got = ops.sample_descriptors(cu(kp), src, H, W, counts=cu(counts), channels_last=(layout == "nhwc")).cpu().numpy()
            for b in range(B):
                n = int(counts[b])
                ref = oracle.interpolate_descriptors(kp[b, :n], dn[b], H, W)
                np.testing.assert_allclose(got[b, :n], ref, rtol=1e-5, atol=1e-7, err_msg=str((layout, B, D, Hc, Wc, K)))
                assert not got[b, n:].any()


def _sets(rng, N1, N2, D, ties):
    a = rng.randn(N1, D).astype(np.float32)
    a /= np.maximum(np.linalg.norm(a, axis=1, keepdims=True), 1e-12)
    b = rng.randn(N2, D).astype(np.float32)
    m = min(N1, N2)
    if m:
        b[:m] = a[rng.permutation(N1)[:m]] + 0.05 * rng.randn(m, D).astype(np.float32)   # planted partners
    b /= np.maximum(np.linalg.norm(b, axis=1, keepdims=True), 1e-12)
    if ties and N2 > 3:
        b[N2 - 1] = b[0]                              # exact duplicates: ties to the lower index
        b[N2 // 2] = b[1]
    return a.astype(np.float32), b.astype(np.float32)


def test_fuzz_matching(ops, oracle):
    rng = np.random.RandomState(1004)
    for it in range(ITERS):
        D = int(rng.choice([16, 64, 64, 128, 256, 256, 100]))
        N1, N2 = int(rng.randint(1, 700)), int(rng.randint(2, 700))
        a, b = _sets(rng, N1, N2, D, ties=rng.rand() < 0.5)
        metric, mode = [('nn', 'nn'), ('l2', 'bf')][int(rng.randint(2))]
        tag = str((D, N1, N2, metric))
        for algo in (['tensor', 'simt'] if D % 64 == 0 else ['simt']):
            # nearest neighbours both ways = the fp64 argmin over the fp32 inputs, ties to the lowest index
            got = ops.nearest(cu(a), cu(b), metric=metric, algo=algo, want_scores=False)
            want = oracle.nearest(a, b, mode=mode, f64=True)
            np.testing.assert_array_equal(got['idx12'][0].cpu().numpy(), want['idx12'], err_msg=tag + algo)
            np.testing.assert_array_equal(got['idx21'][0].cpu().numpy(), want['idx21'], err_msg=tag + algo)
            # mutual matches (crossCheck / NNMatcher with a threshold)
            thr = float(rng.choice([-1.0, 0.7])) if metric == 'nn' else -1.0
            q, t, d, c = ops.match(cu(a), cu(b), metric=metric, algo=algo, kind='mutual', cross_check=True, threshold=thr)
            wq, wt, wd = oracle.match_mutual(a, b, mode=mode, f64=True, cross_check=True, threshold=thr)
            n = int(c[0])
            assert n == len(wq), tag + algo
            np.testing.assert_array_equal(q[0, :n].cpu().numpy(), wq, err_msg=tag + algo)
            np.testing.assert_array_equal(t[0, :n].cpu().numpy(), wt, err_msg=tag + algo)
            np.testing.assert_allclose(d[0, :n].cpu().numpy(), wd, rtol=2e-5, atol=2e-6, err_msg=tag + algo)
            # Lowe ratio on the two nearest (knn_matches=True)
            if metric == 'l2':
                q, t, d, c = ops.match(cu(a), cu(b), metric='l2', algo=algo, kind='ratio', cross_check=False, ratio=0.9)
                wq, wt, wd = oracle.match_ratio(a, b, mode='bf', f64=True, ratio=0.9)
                n = int(c[0])
                got_pairs = set(zip(q[0, :n].cpu().numpy().tolist(), t[0, :n].cpu().numpy().tolist()))
                want_pairs = set(zip(wq.tolist(), wt.tolist()))
                # a ratio within rounding of 0.9 may fall either side (documented near-tie): everything else agrees
                assert len(got_pairs ^ want_pairs) <= max(1, len(want_pairs) // 200), tag + algo


def test_fuzz_warp_and_masks(ops, utils, oracle):
    rng = np.random.RandomState(1005)
    for it in range(max(2, ITERS // 2)):
        H, W = int(rng.randint(8, 120)), int(rng.randint(8, 160))
        N, n = int(rng.randint(1, 5)), int(rng.randint(1, 7))     # n >= 4 takes the gather-array path
        np.random.seed(7000 + it)
        cfg = utils._check_ha_config({'num': n + 1})
        Hs, _ = utils.sample_adaptation_homographies((H, W), cfg, with_masks=False)
        A = utils.normalized_warp_matrix(torch.from_numpy(Hs.astype(np.float32)), (H, W), (H, W))
        src = rng.rand(N, H, W).astype(np.float32)
        for mode, pad in (('bilinear', 'reflection'), ('bilinear', 'zeros'), ('nearest', 'zeros')):
            got = ops.warp(cu(src), A.cuda(), mode, pad).cpu().numpy()
            for i in range(n):
                want = oracle.warp(src, A[i].numpy(), mode, pad)
                np.testing.assert_allclose(got[i], want, rtol=1e-6, atol=1e-7, err_msg=str((H, W, N, n, mode, pad, i)))
        if N % 2 == 0:
            both = ops.warp(cu(src), A.cuda(), 'bilinear', 'reflection', groups=2).cpu().numpy()
            np.testing.assert_array_equal(both[0], ops.warp(cu(src[:N // 2]), A.cuda(), 'bilinear', 'reflection').cpu().numpy())
            np.testing.assert_array_equal(both[1], ops.warp(cu(src[N // 2:]), A.cuda(), 'bilinear', 'reflection').cpu().numpy())
        # valid masks: cv2.warpPerspective(INTER_NEAREST) + border + erosion, bit for bit
        r, border = int(rng.choice([0, 2, 5])), bool(rng.rand() < 0.5)
        Minv = torch.from_numpy(utils.invert_homographies(Hs)).cuda()
        got = ops.valid_masks(Minv, H, W, r, border).cpu().numpy()
        for i in range(n):
            want = oracle.valid_mask((H, W), Hs[i], r, border)
            np.testing.assert_array_equal(got[i] != 0, want != 0, err_msg=str((H, W, r, border, i)))


def test_fuzz_homographic_adaptation(utils, ops, oracle):
    """The whole entry point (sample order, masks built on the device, both spectra warped in one launch, chunked
    accumulation, finish) against the C oracle's chain with the same stub network and the same normalised matrices."""
    rng = np.random.RandomState(1006)
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(1, 65, 8, stride=8).cuda()

    def net(data):
        with torch.no_grad():
            return {'prob': ops.detector_head(conv(data['image']).float())}

    for it in range(max(2, ITERS // 3)):
        Hc, Wc = int(rng.randint(2, 12)), int(rng.randint(2, 14))
        H, W = 8 * Hc, 8 * Wc
        B, num = int(rng.randint(1, 3)), int(rng.randint(2, 9))
        agg = str(rng.choice(['none', 'prod', 'sum']))
        cfg = dict(num=num, min_count=int(rng.choice([0, 2])), erosion_radius=int(rng.choice([0, 3])), mask_border=bool(rng.rand() < 0.5),
                   filter_size=0, sample_chunk=int(rng.choice([0, 1, 3])))
        np.random.seed(8000 + it)
        full = utils._check_ha_config(dict(cfg))
        Hs, masks = utils.sample_adaptation_homographies((H, W), full, with_masks=True)
        Hm = torch.from_numpy(Hs.astype(np.float32))
        A_w = utils.normalized_warp_matrix(Hm, (H, W), (H, W)).numpy()
        A_u = utils.normalized_warp_matrix(torch.inverse(Hm), (H, W), (H, W)).numpy()
        img_o, img_t = rng.rand(B, 1, H, W).astype(np.float32), rng.rand(B, 1, H, W).astype(np.float32)

        def net_prob(imgs, spectrum):
            return net({'image': cu(imgs.reshape(-1, 1, H, W))})['prob'][:, 0].cpu().numpy()

        tag = str((H, W, B, num, agg, cfg))
        if agg == 'none':
            got = utils.homographic_adaptation({'image': cu(img_o)}, net, dict(cfg), homographies=Hs, masks=None, normalized_matrices=(A_w, A_u))
            want = oracle.homographic_adaptation(img_o[:, 0], net_prob, Hs, masks.astype(np.float32), cfg['min_count'], A_warp=A_w, A_unwarp=A_u)
        else:
            data = {'optical': {'image': cu(img_o), 'is_optical': torch.ones(B, 1, dtype=torch.bool, device="cuda")},
                    'thermal': {'image': cu(img_t), 'is_optical': torch.zeros(B, 1, dtype=torch.bool, device="cuda")}}
            got = utils.homographic_adaptation_multispectral(data, net, dict(cfg, aggregation=agg), homographies=Hs, masks=None,
                                                             normalized_matrices=(A_w, A_u))
            want = oracle.homographic_adaptation(img_o[:, 0], net_prob, Hs, masks.astype(np.float32), cfg['min_count'], images_b=img_t[:, 0],
                                                 aggregation=agg, A_warp=A_w, A_unwarp=A_u)
        np.testing.assert_allclose(got[:, 0].cpu().numpy(), want, rtol=1e-5, atol=1e-7, err_msg=tag)


def test_larger_than_benchmark_shapes(utils, ops, oracle):
    """Images larger than 512 x 640 (the sparse top-k kernel's whole-image bitmap no longer fits shared memory: the tile
    kernels take over), more keypoints than the benchmark's 2048, and a matching problem with ragged tile edges."""
    for seed, H, W in ((9001, 1030, 1284), (9002, 777, 1001)):
        hm = syn.heatmap(seed, 2, H, W)
        for topk in (0, 3000):
            want = oracle.box_nms(hm, 4, 0.015, keep_top_k=topk)
            dense, kp, sc, cnt = ops.box_nms(cu(hm[:, 0]), 4, 0.015, keep_top_k=topk, want_keypoints=True, kp_cap=(topk or 200000))
            np.testing.assert_array_equal(dense.cpu().numpy(), want[:, 0])
            for b in range(2):
                ref = oracle.extract_keypoints(want[b, 0], 0.0)
                assert int(cnt[b]) == len(ref)
                np.testing.assert_array_equal(kp[b, :len(ref)].cpu().numpy(), ref)
    rng = np.random.RandomState(9003)
    a, b = _sets(rng, 3001, 2777, 256, ties=True)
    got = ops.nearest(cu(a), cu(b), metric='l2', algo='tensor', want_scores=False)
    want = oracle.nearest(a, b, mode='bf', f64=True)
    np.testing.assert_array_equal(got['idx12'][0].cpu().numpy(), want['idx12'])
    np.testing.assert_array_equal(got['idx21'][0].cpu().numpy(), want['idx21'])


def test_fuzz_glue_and_magicleap(ops, oracle):
    """The fused backbone glue (any C, H, W incl. the scalar path for W % 4 != 0) against the torch module sequence, and
    the MagicLeap heatmap arithmetic against the oracle, on random shapes."""
    rng = np.random.RandomState(1007)
    torch.manual_seed(7)
    for it in range(ITERS):
        B, C = int(rng.randint(1, 4)), int(rng.randint(1, 40))
        H, W = int(rng.randint(2, 50)), int(rng.randint(2, 70))
        pool = bool(rng.rand() < 0.5)
        if pool:
            H, W = 2 * H, 2 * W
        x = torch.randn(B, C, H, W, device="cuda") * 2
        bn = torch.nn.BatchNorm2d(C).cuda().eval()
        with torch.no_grad():
            bn.weight.uniform_(-1.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.2, 3.0)
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            shift = bn.bias - bn.running_mean * scale
            bn_first, pad, reflect = bool(rng.rand() < 0.5), int(rng.randint(0, 2)), bool(rng.rand() < 0.5)
            cbias = torch.randn(C, device="cuda") if rng.rand() < 0.5 else None
            xb = x if cbias is None else x + cbias[None, :, None, None]
            ref = torch.relu(bn(xb)) if bn_first else bn(torch.relu(xb))
            if pool:
                ref = torch.nn.functional.max_pool2d(ref, 2, 2)
            if pad:
                ref = (torch.nn.ReflectionPad2d(1) if reflect else torch.nn.ZeroPad2d(1))(ref)
            got = ops.relu_bn_pad(x, scale, shift, bn_first=bn_first, pool=pool, pad=pad, reflect=reflect, conv_bias=cbias)
            torch.testing.assert_close(got, ref, rtol=2e-6, atol=2e-6, msg=str((B, C, H, W, pool, bn_first, pad, reflect)))
            # first encoder layer
            img = torch.rand(B, 1, H, W, device="cuda")
            conv = torch.nn.Conv2d(1, C, 3).cuda()
            in_reflect = bool(rng.rand() < 0.5)
            y = conv((torch.nn.ReflectionPad2d(1) if in_reflect else torch.nn.ZeroPad2d(1))(img))
            ref = torch.relu(bn(y)) if bn_first else bn(torch.relu(y))
            if pad:
                ref = (torch.nn.ReflectionPad2d(1) if reflect else torch.nn.ZeroPad2d(1))(ref)
            got = ops.conv1_relu_bn_pad(img, conv.weight, conv.bias, scale, shift, bn_first=bn_first, in_reflect=in_reflect,
                                        pad=pad, out_reflect=reflect)
            torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5, msg=str((B, C, H, W, bn_first, in_reflect, pad, reflect)))
        lg = syn.logits(9100 + it, B, int(rng.randint(1, 30)), int(rng.randint(1, 30)), sigma=float(rng.choice([0.5, 3.0])), bias=float(rng.choice([0.0, 6.0])))
        np.testing.assert_allclose(ops.heatmap_magicleap(cu(lg)).cpu().numpy(), oracle.heatmap_magicleap(lg), rtol=1e-5, atol=1e-12)


def test_fuzz_evaluation_points(ops, oracle):
    rng = np.random.RandomState(1008)
    for it in range(ITERS):
        H, W = int(rng.randint(8, 300)), int(rng.randint(8, 400))
        nq, nt = int(rng.randint(1, 400)), int(rng.randint(1, 400))
        q = rng.randint(-10, max(H, W) + 10, (1, nq, 2)).astype(np.int64)
        t = np.stack([rng.randint(0, H, (1, nt)), rng.randint(0, W, (1, nt))], axis=2).astype(np.int64)
        got = ops.points_min_dist2(cu(q), cu(t), H, W)[0].cpu().numpy()
        np.testing.assert_array_equal(got, oracle.points_min_dist2(q[0], t[0], H, W), err_msg=str((H, W, nq, nt)))
        qw = q[0].astype(np.float64) + rng.rand(nq, 2)
        thr = float(rng.choice([1.0, 4.0, 6.0]))
        ra, _ = ops.points_correct(cu(qw[None]), cu(t), thr)
        np.testing.assert_array_equal(ra[0].cpu().numpy(), oracle.points_correct(qw, t[0], thr)[0], err_msg=str((H, W, nq, nt, thr)))
