"""Pin the CPU oracle (oracle/mp_oracle.c) against the fixtures frozen from the reference's own
functions by oracle/gen_golden.py.  CPU only.  Integer / index results are bit-exact; floating
point within the tolerance the north star states (1e-5 relative), written per assert."""
import numpy as np
import pytest

from conftest import dense_from_sparse, load_golden
from multipoint_b200 import synthetic as syn

RTOL = 1e-5


def assert_equal_up_to_topk_ties(got, want):
    """4-D box_nms with keep_top_k and EXACT score ties straddling k: the reference's batched path
    (torchvision _batched_nms_vanilla) ends in a non-stable sort, so which of the tied survivors
    make the cut is undefined there.  Defined and checked: per image, the
    multiset of kept scores and every keypoint strictly above the cut score."""
    assert got.shape == want.shape
    for b in range(got.shape[0]):
        gb, wb = got[b], want[b]
        np.testing.assert_array_equal(np.sort(gb[gb > 0]), np.sort(wb[wb > 0]))
        if (wb > 0).any():
            cut = wb[wb > 0].min()
            np.testing.assert_array_equal(gb > cut, wb > cut)


def test_detector_head_matches_reference(oracle):
    g = load_golden("heads")
    prob = oracle.detector_head(g["logits"])
    np.testing.assert_allclose(prob, g["prob"], rtol=RTOL, atol=1e-9)
    # full size, regenerated from the seed
    B, _, Hc, Wc = (int(v) for v in g["full_seed"][[1, 0, 2, 3]])
    lg = syn.logits(int(g["full_seed"][0]), B, Hc, Wc, *g["full_sigma_bias"])
    assert syn.checksum(lg) == str(g["full_checksum"])
    full = oracle.detector_head(lg)
    np.testing.assert_allclose(full[0, 0, ::37], g["full_rows"], rtol=RTOL, atol=1e-9)
    assert abs(full.astype(np.float64).sum() - float(g["full_sum"])) < 1e-6 * float(g["full_sum"])
    # depth_to_space(x, 8) is the same index map as PixelShuffle(8) (utils.py:64-69)
    x = g["logits"][:, :64]
    d2s = x.reshape(2, 8, 8, 1, 8, 10).transpose(0, 3, 4, 1, 5, 2).reshape(2, 1, 64, 80)
    np.testing.assert_array_equal(d2s, g["depth_to_space"])


def test_heatmap_magicleap_matches_reference(oracle):
    """SuperPointMagicLeap.generate_heatmap (numpy exp, +1e-5, no max subtraction)."""
    g = load_golden("magicleap")
    np.testing.assert_allclose(oracle.heatmap_magicleap(g["semi"]), g["prob"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(oracle.heatmap_magicleap(g["model_logits"]), g["model_prob"], rtol=RTOL, atol=1e-12)
    assert np.isfinite(g["prob"]).all() and g["prob"][0, 0, 0, 5] > 0.99  # the exp(40) cell


def test_normalize_descriptors_matches_reference(oracle):
    g = load_golden("heads")
    np.testing.assert_allclose(oracle.normalize_descriptors(g["desc_in"]), g["desc"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(oracle.normalize_descriptors(g["desc256_in"]), g["desc256"], rtol=RTOL, atol=1e-9)
    assert np.all(oracle.normalize_descriptors(g["desc_in"])[0, :, 0, 0] == 0.0)  # zero-norm cell


@pytest.mark.parametrize("tag,size,topk", [("chain", 4, 0), ("tie4", 4, 0), ("strict", 4, 0), ("top2", 4, 2)])
def test_box_nms_hand_cases(oracle, tag, size, topk):
    g = load_golden("box_nms")
    p = g[tag + "_in"]
    want = dense_from_sparse(g[tag + "_idx"], g[tag + "_val"], p.shape)
    for literal in (True, False):
        got = oracle.box_nms(p, size, 0.015, keep_top_k=topk, literal=literal)
        np.testing.assert_array_equal(got, want)
    if tag == "chain":
        assert list(g["chain_idx"]) == [10 * 32 + 10, 10 * 32 + 16]  # A and C survive, B does not
    if tag == "tie4":
        assert list(g["tie4_idx"]) == [9 * 32 + 11]  # lowest row-major index wins the tie


@pytest.mark.parametrize("size,iou", [(3, 0.1), (4, 0.1), (8, 0.1), (4, 0.3), (5, 0.1), (2.5, 0.1)])
def test_box_nms_footprint(oracle, size, iou):
    g = load_golden("box_nms")
    want = g["footprint_s%s_i%s" % (size, iou)]
    np.testing.assert_array_equal(oracle.nms_footprint(size, iou), want)
    if (size, iou) == (4, 0.1):
        assert want.sum() == 36  # |dy|,|dx|<=3 minus (3,3),(3,2),(2,3) corners, minus the centre


def test_box_nms_small_random(oracle):
    g = load_golden("box_nms")
    for seed, size, topk, quant, B in g["small_cases"]:
        seed, topk, B = int(seed), int(topk), int(B)
        hm = syn.heatmap(seed, B, 64, 80, quant=(quant or None))
        tag = "small%d" % seed
        want4 = dense_from_sparse(g[tag + "_4d_idx"], g[tag + "_4d_val"], hm.shape)
        want2 = dense_from_sparse(g[tag + "_2d_idx"], g[tag + "_2d_val"], hm.shape[-2:])
        for literal in (True, False):
            got4 = oracle.box_nms(hm, size, 0.015, keep_top_k=topk, literal=literal)
            if quant and topk:
                assert_equal_up_to_topk_ties(got4, want4)
            else:
                np.testing.assert_array_equal(got4, want4)
            # the 2-D path (torchvision nms, stable sort) is defined even with ties at k
            np.testing.assert_array_equal(oracle.box_nms(hm[0, 0], size, 0.015, keep_top_k=topk, literal=literal), want2)


def test_box_nms_full_size(oracle):
    g = load_golden("box_nms")
    for seed, topk, quant, B in g["full_cases"]:
        seed, topk, B = int(seed), int(topk), int(B)
        hm = syn.heatmap(seed, B, 512, 640, quant=(quant or None))
        assert syn.checksum(hm) == str(g["full%d_checksum" % seed])
        want = dense_from_sparse(g["full%d_4d_idx" % seed], g["full%d_4d_val" % seed], hm.shape)
        got = oracle.box_nms(hm, 4, 0.015, keep_top_k=topk)
        if quant and topk:
            assert_equal_up_to_topk_ties(got, want)
        else:
            np.testing.assert_array_equal(got, want)
    # heatmap from the reference's own softmax, stored in the fixture
    prob = g["softmax34_prob"]
    want = dense_from_sparse(g["softmax34_idx"], g["softmax34_val"], prob.shape)
    got = oracle.box_nms(prob, 4, 0.015, keep_top_k=2048)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(oracle.extract_keypoints(got[0, 0], 0.015), g["softmax34_kp"])


def test_interpolate_descriptors(oracle):
    g = load_golden("interpolate")
    for D in (64, 256):
        dm = syn.descriptor_map(int(g["d%d_seed" % D][0]), 1, D, 64, 80)[0]
        assert syn.checksum(dm) == str(g["d%d_checksum" % D])
        got = oracle.interpolate_descriptors(g["d%d_kp" % D], dm, 512, 640)
        np.testing.assert_allclose(got, g["d%d_out" % D], rtol=RTOL, atol=2e-7)
    got = oracle.interpolate_descriptors(g["small_kp"], g["small_in"], 40, 56)
    np.testing.assert_allclose(got, g["small_out"], rtol=RTOL, atol=2e-7)
    assert oracle.interpolate_descriptors(np.zeros((0, 2), np.int64), g["small_in"], 40, 56).shape == tuple(g["empty_shape"])


def _inputs(g, case):
    seed, N1, N2, D, noise, dup = case
    tag = "m%d" % int(seed)
    if tag + "_a" in g.files:
        a, b = g[tag + "_a"], g[tag + "_b"]
    else:
        a, b = syn.descriptor_sets(int(seed), int(N1), int(N2), int(D), float(noise), int(dup))
    assert syn.checksum(a) + syn.checksum(b) == str(g[tag + "_checksum"])
    return tag, a, b


def test_matching_against_reference(oracle):
    """bfmatcher(crossCheck), nnmatcher and the knn ratio test: indices identical, distances 1e-5.
    Every oracle precision (fp32 restatement and fp64 truth) must agree with the reference here:
    the data has no near-ties (planted exact duplicates resolve to the lowest index everywhere)."""
    g = load_golden("matching")
    for case in g["cases"]:
        tag, a, b = _inputs(g, case)
        for f64 in (False, True):
            q, t, d = oracle.match_mutual(a, b, 'bf', f64, cross_check=True)
            np.testing.assert_array_equal(q, g[tag + "_bf_q"])
            np.testing.assert_array_equal(t, g[tag + "_bf_t"])
            np.testing.assert_allclose(d, g[tag + "_bf_d"], rtol=RTOL, atol=1e-6)
            q, t, d = oracle.match_mutual(a, b, 'bf', f64, cross_check=False)
            np.testing.assert_array_equal(q, g[tag + "_bfnc_q"])
            np.testing.assert_array_equal(t, g[tag + "_bfnc_t"])
            for thr, key in [(0.7, "_nn"), (1.1, "_nn11")]:
                q, t, d = oracle.match_mutual(a, b, 'nn', f64, cross_check=True, threshold=thr)
                np.testing.assert_array_equal(q, g[tag + key + "_q"])
                np.testing.assert_array_equal(t, g[tag + key + "_t"])
                # sqrt(2-2s) amplifies the fp32 rounding of s near s=1: absolute 2e-4 there
                np.testing.assert_allclose(d, g[tag + key + "_d"], rtol=RTOL, atol=2e-4)
            if tag + "_knn_q" in g.files:
                q, t, d = oracle.match_ratio(a, b, 'bf', f64, 0.9)
                np.testing.assert_array_equal(q, g[tag + "_knn_q"])
                np.testing.assert_array_equal(t, g[tag + "_knn_t"])
            if tag + "_thr_q" in g.files:
                q, t, d = oracle.match_threshold(a, b, 0.9, f64)
                np.testing.assert_array_equal(q, g[tag + "_thr_q"])
                np.testing.assert_array_equal(t, g[tag + "_thr_t"])
    a, b = syn.descriptor_sets(55, 1024, 1024, 256, 0.05, 0)
    nn = oracle.nearest(a, b, 'nn', f64=True)
    np.testing.assert_array_equal(nn["idx12"], g["m55_f64_row"])
    np.testing.assert_array_equal(nn["idx21"], g["m55_f64_col"])
    assert oracle.match_mutual(np.zeros((0, 64), np.float32), b[:, :64].copy(), 'nn', threshold=0.7)[0].size == int(g["empty_nn"])


def test_warp_matches_torch_restatement(oracle):
    """Rows 9-10 (PARITY UNPINNED: kornia absent).  The C warp against the torch grid_sample
    restatement frozen in the fixture.  Tolerance: 1e-5 of the image range plus 1e-5 relative --
    the two sides round the 3x3 algebra differently (LAPACK vs torch), which moves the sample
    point by ~1e-5 px."""
    g = load_golden("adaptation")
    H, W = g["img_o"].shape[-2:]
    M = g["H"][0].astype(np.float32)
    A = oracle.warp_matrix(M, H, W)
    Ainv = oracle.warp_matrix(np.linalg.inv(M).astype(np.float32), H, W)
    got = oracle.warp(g["img_o"], A, 'bilinear', 'reflection')
    np.testing.assert_allclose(got, g["warp_bilinear_reflection"], rtol=1e-5, atol=2e-5)
    got = oracle.warp(g["img_o"], Ainv, 'bilinear', 'zeros')
    np.testing.assert_allclose(got, g["warp_bilinear_zeros"], rtol=1e-5, atol=2e-5)
    got = oracle.warp(np.repeat(g["masks"][None, :1], 2, 0), Ainv, 'nearest', 'zeros')
    assert (got != g["warp_nearest_zeros"]).mean() < 2e-3  # nearest flips only on half-pixel ties


def assert_close_but_mask_ties(got, want, rtol=1e-4, atol=2e-6, outliers=1e-3):
    """Aggregated heatmaps agree to rtol/atol except where the NEAREST-sampled valid mask flips on
    a half-pixel tie between the two 3x3 roundings (a handful of pixels on the mask border)."""
    bad = np.abs(got - want) > atol + rtol * np.abs(want)
    assert bad.mean() <= outliers, bad.mean()
    assert np.abs(got - want).max() < 0.05 * max(1e-6, np.abs(want).max())


def _stub_net(g):
    w, b = g["stub_w"].reshape(65, 64).astype(np.float32), g["stub_b"].astype(np.float32)

    def net_prob(images, spectrum, _oracle=[None]):
        B, H, W = images.shape
        cells = images.reshape(B, H // 8, 8, W // 8, 8).transpose(0, 1, 3, 2, 4).reshape(B, H // 8, W // 8, 64)
        lg = (cells @ w.T + b).transpose(0, 3, 1, 2)
        return lg
    return net_prob


def test_homographic_adaptation_matches_restatement(oracle):
    g = load_golden("adaptation")
    to_logits = _stub_net(g)

    def net_prob(images, spectrum):
        return oracle.detector_head(to_logits(images, spectrum))[:, 0]

    img_o, img_t = g["img_o"][:, 0], g["img_t"][:, 0]
    # Heatmap values are ~1e-2..0.7.  The fp32 sampling coordinate is only defined to ~1e-5 px
    # (FMA / matmul order in torch vs the plain C order), so on these noise images the agreement
    # is 1e-4 relative + 2e-6 absolute with <=0.3 % outliers; once with the oracle's own 3x3
    # algebra (numpy LAPACK) and once with the matrices exactly as torch rounded them.
    for mats in (dict(), dict(A_warp=g["A_warp"], A_unwarp=g["A_unwarp"])):
        out = oracle.homographic_adaptation(img_o, net_prob, g["H"], g["masks"], min_count=2, **mats)
        assert_close_but_mask_ties(out, g["single"][:, 0], outliers=3e-3)
        for agg in ("prod", "sum"):
            out = oracle.homographic_adaptation(img_o, net_prob, g["H"], g["masks"], min_count=2,
                                                images_b=img_t, aggregation=agg, **mats)
            assert_close_but_mask_ties(out, g["multi_" + agg][:, 0], outliers=3e-3)
    # the 3x3 algebra itself: numpy LAPACK vs torch.inverse agree to a few ulp
    H, W = img_o.shape[-2:]
    for i in range(len(g["H"])):
        M = g["H"][i].astype(np.float32)
        np.testing.assert_allclose(oracle.warp_matrix(M, H, W), g["A_warp"][i], rtol=0, atol=1e-6)
        np.testing.assert_allclose(oracle.warp_matrix(np.linalg.inv(M).astype(np.float32), H, W),
                                   g["A_unwarp"][i], rtol=0, atol=1e-6)


def test_reference_port_matches_reference_fixtures():
    """oracle/reference_port.py (the library-level port bench.py times as the CPU baseline) gives
    exactly what the real reference produced for the fixtures."""
    import sys
    import torch
    from conftest import ROOT
    sys.path.insert(0, ROOT + "/oracle")
    import reference_port as rp
    g = load_golden("heads")
    np.testing.assert_array_equal(rp.detector_head(torch.from_numpy(g["logits"])).numpy(), g["prob"])
    np.testing.assert_array_equal(rp.descriptor_head(torch.from_numpy(g["desc_in"])).numpy(), g["desc"])
    g = load_golden("box_nms")
    for seed, size, topk, quant, B in g["small_cases"]:
        seed, topk, B = int(seed), int(topk), int(B)
        hm = syn.heatmap(seed, B, 64, 80, quant=(quant or None))
        tag = "small%d" % seed
        want2 = dense_from_sparse(g[tag + "_2d_idx"], g[tag + "_2d_val"], hm.shape[-2:])
        np.testing.assert_array_equal(rp.box_nms(torch.from_numpy(hm[0, 0]), size, 0.015, keep_top_k=topk).numpy(), want2)
        if not (quant and topk):
            want4 = dense_from_sparse(g[tag + "_4d_idx"], g[tag + "_4d_val"], hm.shape)
            np.testing.assert_array_equal(rp.box_nms(torch.from_numpy(hm), size, 0.015, keep_top_k=topk).numpy(), want4)
    g = load_golden("interpolate")
    got = rp.interpolate_descriptors(torch.from_numpy(g["small_kp"]), torch.from_numpy(g["small_in"]), 40, 56)
    np.testing.assert_array_equal(got.numpy(), g["small_out"])
    g = load_golden("matching")
    m = rp.get_matches_bf_crosscheck(g["m51_a"], g["m51_b"])
    assert [x.queryIdx for x in m] == list(g["m51_bf_q"]) and [x.trainIdx for x in m] == list(g["m51_bf_t"])


def test_valid_mask_matches_reference(oracle):
    """the restated OpenCV raster (inverse, block origin, reciprocal, half-to-even) and erosion
    against the masks the reference's compute_valid_mask produced (tests/golden/homographies.npz)."""
    g = load_golden("homographies")
    for tag in ("default", "export", "noart", "small"):
        shape = tuple(int(v) for v in g[tag + "_shape"])
        erosion = int(g[tag + "_seed"][1])
        want = np.unpackbits(g[tag + "_mask"], axis=-1)[..., :shape[1]]
        for i, Hm in enumerate(g[tag + "_H"]):
            np.testing.assert_array_equal(oracle.valid_mask(shape, Hm, erosion, True), want[i])
    Hm = g["small_H"][0]
    np.testing.assert_array_equal(oracle.valid_mask((64, 80), Hm, 0, False), np.unpackbits(g["small_mask_e0"], axis=-1)[..., :80])
    np.testing.assert_array_equal(oracle.valid_mask((64, 80), Hm, 2, False), np.unpackbits(g["small_mask_e2nb"], axis=-1)[..., :80])
    # live against OpenCV (same process, no fixture): odd sizes, every erosion mode
    import cv2
    rng = np.random.default_rng(9)
    for shape in [(37, 53), (12, 20), (70, 131)]:
        for _ in range(20):
            th = rng.uniform(-0.5, 0.5)
            Hm = np.array([[np.cos(th) * rng.uniform(0.8, 1.2), -np.sin(th), rng.uniform(-8, 8)],
                           [np.sin(th), np.cos(th) * rng.uniform(0.8, 1.2), rng.uniform(-8, 8)],
                           [rng.uniform(-2e-3, 2e-3), rng.uniform(-2e-3, 2e-3), 1.0]])
            np.testing.assert_array_equal(oracle.invert3x3(Hm), cv2.invert(Hm)[1])
            for r, border in [(0, False), (1, True), (4, False)]:
                ref = cv2.warpPerspective(np.ones(shape), Hm, shape[::-1], flags=cv2.INTER_NEAREST)
                if r > 0:
                    if border:
                        ref = np.pad(ref, 1)
                    ref = cv2.erode(ref, np.ones((2 * r + 1, 2 * r + 1), np.float32), iterations=1)
                    if border:
                        ref = ref[1:-1, 1:-1]
                np.testing.assert_array_equal(oracle.valid_mask(shape, Hm, r, border), ref.astype(np.uint8))


def test_evaluation_point_geometry_matches_reference(oracle):
    """warp_keypoints / filter_points + nearest distance / correct-match flags, against the
    reference's own expressions frozen in tests/golden/evaluation.npz."""
    g = load_golden("evaluation")
    for i in range(6):
        Hm, a, b = g["pt%d_H" % i], g["pt%d_a" % i], g["pt%d_b" % i]
        wi = oracle.warp_keypoints(a, Hm)
        np.testing.assert_array_equal(wi, g["pt%d_warp_int" % i])
        wf = oracle.warp_keypoints(a, Hm, as_int=False)
        np.testing.assert_array_equal(wf, g["pt%d_warp_f64" % i])                   # bit-exact doubles
        d2 = oracle.points_min_dist2(wi, b, 128, 160)
        np.testing.assert_array_equal(wi[d2 >= 0], g["pt%d_filtered" % i])          # filter_points
        np.testing.assert_array_equal(np.sqrt(d2[d2 >= 0].astype(np.float64)), g["pt%d_min_dist" % i])
        row_any, tp = oracle.points_correct(wf, b, 4.0, np.arange(150), np.arange(150))
        np.testing.assert_array_equal(np.flatnonzero(row_any), g["pt%d_correct_rows" % i])
        np.testing.assert_array_equal(tp.astype(bool), g["pt%d_correct_diag" % i])
    assert (oracle.points_min_dist2(np.array([[3, 3]]), np.zeros((0, 2)), 10, 10) == np.iinfo(np.int64).max).all()


def test_backbone_glue_matches_torch_modules(oracle):
    """Row 3 glue (MultiPoint.py:61-90): the oracle's ReLU / BatchNorm(eval) / MaxPool / pad chain and the
    one-input-channel first layer against the torch modules the reference builds them from."""
    import torch
    torch.manual_seed(4)
    x = torch.randn(2, 5, 12, 14) * 2
    bn = torch.nn.BatchNorm2d(5).eval()
    conv = torch.nn.Conv2d(1, 5, 3)
    img = torch.rand(2, 1, 11, 9)
    with torch.no_grad():
        bn.weight.uniform_(-1.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.2, 3.0)
        args = (bn.running_mean.numpy(), bn.running_var.numpy(), bn.eps, bn.weight.numpy(), bn.bias.numpy())
        cb = torch.randn(5)
        for bn_first in (False, True):
            for pool in (False, True):
                for pad, reflect in [(0, True), (1, True), (1, False)]:
                    xb = x + cb[None, :, None, None]
                    ref = torch.relu(bn(xb)) if bn_first else bn(torch.relu(xb))
                    if pool:
                        ref = torch.nn.functional.max_pool2d(ref, 2, 2)
                    if pad:
                        ref = (torch.nn.ReflectionPad2d(1) if reflect else torch.nn.ZeroPad2d(1))(ref)
                    got = oracle.relu_bn_pad(x.numpy(), *args, conv_bias=cb.numpy(), bn_first=bn_first, pool=pool, pad=pad, reflect=reflect)
                    np.testing.assert_allclose(got, ref.numpy(), rtol=2e-6, atol=2e-6)
            for in_reflect in (True, False):
                for pad, out_reflect in [(0, True), (1, True), (1, False)]:
                    y = conv((torch.nn.ReflectionPad2d(1) if in_reflect else torch.nn.ZeroPad2d(1))(img))
                    ref = torch.relu(bn(y)) if bn_first else bn(torch.relu(y))
                    if pad:
                        ref = (torch.nn.ReflectionPad2d(1) if out_reflect else torch.nn.ZeroPad2d(1))(ref)
                    got = oracle.conv1_relu_bn_pad(img.numpy(), conv.weight.numpy(), conv.bias.numpy(), *args, bn_first=bn_first,
                                                   in_reflect=in_reflect, pad=pad, out_reflect=out_reflect)
                    np.testing.assert_allclose(got, ref.numpy(), rtol=1e-5, atol=1e-5)
