"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs
and against the golden fixtures frozen from the reference.  Bit-exact for keypoints, NMS maps and
match indices; floating point within the tolerance written at each assert (north star: 1e-5
relative for heatmaps and descriptors)."""
import numpy as np
import pytest
import torch

from conftest import dense_from_sparse, load_golden
from multipoint_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
RTOL = 1e-5


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


# ------------------------------------------------------------------ row 1: detector head
def test_detector_head_golden_and_oracle(ops, oracle):
    g = load_golden("heads")
    got = ops.detector_head(cu(g["logits"])).cpu().numpy()
    np.testing.assert_allclose(got, g["prob"], rtol=RTOL, atol=1e-9)            # vs the reference
    np.testing.assert_allclose(got, oracle.detector_head(g["logits"]), rtol=RTOL, atol=1e-9)
    for seed, B, Hc, Wc in [(101, 3, 64, 80), (102, 1, 5, 7), (103, 2, 33, 17)]:  # ragged: warps straddle rows/images
        lg = syn.logits(seed, B, Hc, Wc)
        got = ops.detector_head(cu(lg)).cpu().numpy()
        np.testing.assert_allclose(got, oracle.detector_head(lg), rtol=RTOL, atol=1e-9)
    # size-independent property at the bench size: each 8x8 cell sums to 1 - dustbin probability <= 1
    lg = cu(syn.logits(104, 8, 64, 80))
    prob = ops.detector_head(lg)
    cell = prob.reshape(8, 64, 8, 80, 8).sum(dim=(2, 4))
    dust = torch.softmax(lg, 1)[:, 64]
    torch.testing.assert_close(cell, 1.0 - dust, rtol=1e-4, atol=1e-6)


def test_heatmap_magicleap(ops, oracle):
    """SuperPointMagicLeap.generate_heatmap on the device, vs the reference fixture and the oracle."""
    g = load_golden("magicleap")
    got = ops.heatmap_magicleap(cu(g["semi"])).cpu().numpy()
    np.testing.assert_allclose(got, g["prob"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(got, oracle.heatmap_magicleap(g["semi"]), rtol=RTOL, atol=1e-12)
    lg = syn.logits(106, 3, 33, 17, sigma=3.0, bias=6.0)
    np.testing.assert_allclose(ops.heatmap_magicleap(cu(lg)).cpu().numpy(), oracle.heatmap_magicleap(lg), rtol=RTOL, atol=1e-12)
    # the mirrored model: same seeded weights as the reference -> same logits / descriptors / heatmap
    from multipoint_b200.models import SuperPointMagicLeap
    torch.manual_seed(int(g["model_seed"]))
    net = SuperPointMagicLeap().eval().cuda()
    with torch.no_grad():
        o = net({'image': cu(g["model_image"])})
    np.testing.assert_allclose(o['logits'].cpu().numpy(), g["model_logits"], rtol=1e-4, atol=1e-5)  # cuDNN vs CPU conv
    np.testing.assert_allclose(o['desc'].cpu().numpy(), g["model_desc"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(o['prob'].cpu().numpy(), oracle.heatmap_magicleap(o['logits'].cpu().numpy()), rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(o['prob'].cpu().numpy(), g["model_prob"], rtol=1e-3, atol=1e-6)


def test_detector_head_valid_mask(ops, oracle):
    lg = syn.logits(105, 2, 16, 20)
    mask = (np.random.default_rng(5).random((2, 1, 128, 160)) > 0.3)
    got = ops.detector_head(cu(lg), cu(mask)).cpu().numpy()
    np.testing.assert_allclose(got, oracle.detector_head(lg) * mask, rtol=RTOL, atol=1e-9)


def test_depth_to_space(utils):
    g = load_golden("heads")
    x = cu(g["logits"][:, :64].copy())
    np.testing.assert_array_equal(utils.depth_to_space(x, 8).cpu().numpy(), g["depth_to_space"])
    y = torch.randn(2, 12, 5, 7, device="cuda")
    ref = y.view(2, 2, 2, 3, 5, 7).permute(0, 3, 4, 1, 5, 2).reshape(2, 3, 10, 14)
    torch.testing.assert_close(utils.depth_to_space(y, 2), ref, rtol=0, atol=0)


# ------------------------------------------------------------------ row 2: descriptor normalise
def test_normalize_descriptors(ops, oracle):
    g = load_golden("heads")
    for key_in, key_out in [("desc_in", "desc"), ("desc256_in", "desc256")]:
        x = g[key_in]
        nchw, nhwc = ops.normalize_descriptors(cu(x), nchw=True, nhwc=True)
        np.testing.assert_allclose(nchw.cpu().numpy(), g[key_out], rtol=RTOL, atol=1e-9)
        np.testing.assert_allclose(nhwc.permute(0, 3, 1, 2).cpu().numpy(), g[key_out], rtol=RTOL, atol=1e-9)
    for D, HW in [(64, (64, 80)), (256, (64, 80)), (128, (7, 9)), (48, (5, 5)), (300, (4, 6))]:
        x = syn.descriptor_map(7, 2, D, *HW)
        nchw, nhwc = ops.normalize_descriptors(cu(x), nchw=True, nhwc=True)
        want = oracle.normalize_descriptors(x)
        np.testing.assert_allclose(nchw.cpu().numpy(), want, rtol=RTOL, atol=1e-9)
        np.testing.assert_allclose(nhwc.permute(0, 3, 1, 2).cpu().numpy(), want, rtol=RTOL, atol=1e-9)


# ------------------------------------------------------------------ row 4: box_nms
@pytest.mark.parametrize("tag,size,topk", [("chain", 4, 0), ("tie4", 4, 0), ("strict", 4, 0), ("top2", 4, 2)])
def test_box_nms_hand_cases(utils, tag, size, topk):
    g = load_golden("box_nms")
    p = g[tag + "_in"]
    want = dense_from_sparse(g[tag + "_idx"], g[tag + "_val"], p.shape)
    got = utils.box_nms(cu(p), size, 0.015, keep_top_k=topk)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    got = utils.box_nms(torch.from_numpy(p), size, 0.015, keep_top_k=topk)  # host tensor in -> host tensor out
    assert got.device.type == "cpu"
    np.testing.assert_array_equal(got.numpy(), want)


def test_box_nms_sparse_topk_path(utils, oracle):
    """keep_top_k > 0 takes the sparse path (only the candidates that can reach the top k are settled; images it
    cannot handle are redone by the dense kernels).  Every regime against the oracle, bit for bit."""
    rng = np.random.default_rng(21)
    H, W = 96, 160

    def check(hm, size, k, thr=0.015):
        got = utils.box_nms(cu(hm), size, thr, keep_top_k=k).cpu().numpy()
        np.testing.assert_array_equal(got, oracle.box_nms(hm, size, thr, keep_top_k=k), err_msg="size %s k %d" % (size, k))

    noise = syn.heatmap(901, 3, H, W)                                   # i.i.d. peaks
    for k in (1, 7, 100, 400, 5000):                                    # 5000 > all survivors; > SP_LIST_CAP/2: dense path
        check(noise, 4, k)
    check(noise, 3, 50)
    check(noise, 2.5, 50)
    check(noise, 8, 50)                                                 # reach > 3: dense path
    # blobs: every peak is surrounded by slightly lower candidates, so few of the admitted candidates survive and
    # the threshold has to be lowered (retry loop)
    yy, xx = np.mgrid[0:H, 0:W]
    blobs = np.zeros((2, 1, H, W), np.float32)
    for b in range(2):
        for _ in range(60):
            cy, cx, a = rng.integers(0, H), rng.integers(0, W), rng.uniform(0.1, 0.9)
            blobs[b, 0] = np.maximum(blobs[b, 0], (a * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 18.0)).astype(np.float32))
    for k in (5, 40, 300):
        check(blobs, 4, k)
    # long dependency chains: a ramp (every pixel beats its right neighbour) -> round limit -> redone densely
    ramp = (0.02 + 0.9 * (np.arange(H * W, dtype=np.float32)[::-1] / (H * W))).reshape(1, 1, H, W)
    check(ramp, 4, 30)
    # constant map: every pixel ties; more candidates than the list holds at full size -> redone densely
    check(np.full((1, 1, H, W), 0.5, np.float32), 4, 10)
    flat = np.full((1, 1, 512, 640), 0.25, np.float32)
    check(flat, 4, 64)
    # empty map, a single candidate, fewer candidates than k
    check(np.zeros((2, 1, H, W), np.float32), 4, 10)
    one = np.zeros((1, 1, H, W), np.float32)
    one[0, 0, 50, 70] = 0.3
    check(one, 4, 10)
    # a mixed batch: image 0 goes sparse, image 1 (constant) is flagged and redone
    mixed = np.concatenate([syn.heatmap(902, 1, 512, 640), flat])
    check(mixed, 4, 2048)
    # heavy ties at the cut (scores quantised to 1/64): the stable order decides, like the oracle
    check(syn.heatmap(903, 2, H, W, quant=64), 4, 25)
    # not a multiple of 4 wide (scalar staging path)
    check(syn.heatmap(904, 2, 61, 75), 4, 20)


def assert_equal_up_to_topk_ties(got, want):
    for b in range(got.shape[0]):
        gb, wb = got[b], want[b]
        np.testing.assert_array_equal(np.sort(gb[gb > 0]), np.sort(wb[wb > 0]))
        if (wb > 0).any():
            cut = wb[wb > 0].min()
            np.testing.assert_array_equal(gb > cut, wb > cut)


def test_box_nms_small_random_vs_reference_and_oracle(utils, oracle):
    g = load_golden("box_nms")
    for seed, size, topk, quant, B in g["small_cases"]:
        seed, topk, B = int(seed), int(topk), int(B)
        hm = syn.heatmap(seed, B, 64, 80, quant=(quant or None))
        tag = "small%d" % seed
        got4 = utils.box_nms(cu(hm), size, 0.015, keep_top_k=topk).cpu().numpy()
        got2 = utils.box_nms(cu(hm[0, 0]), size, 0.015, keep_top_k=topk).cpu().numpy()
        want4 = dense_from_sparse(g[tag + "_4d_idx"], g[tag + "_4d_val"], hm.shape)
        want2 = dense_from_sparse(g[tag + "_2d_idx"], g[tag + "_2d_val"], hm.shape[-2:])
        np.testing.assert_array_equal(got2, want2)
        if quant and topk:
            assert_equal_up_to_topk_ties(got4, want4)  # undefined in the reference
        else:
            np.testing.assert_array_equal(got4, want4)
        # against the oracle the stable tie rule is defined: bit-exact always
        np.testing.assert_array_equal(got4, oracle.box_nms(hm, size, 0.015, keep_top_k=topk))


def test_box_nms_full_size_vs_reference(utils, ops, oracle):
    g = load_golden("box_nms")
    for seed, topk, quant, B in g["full_cases"]:
        seed, topk, B = int(seed), int(topk), int(B)
        hm = syn.heatmap(seed, B, 512, 640, quant=(quant or None))
        assert syn.checksum(hm) == str(g["full%d_checksum" % seed])
        want = dense_from_sparse(g["full%d_4d_idx" % seed], g["full%d_4d_val" % seed], hm.shape)
        got = utils.box_nms(cu(hm), 4, 0.015, keep_top_k=topk).cpu().numpy()
        if quant and topk:
            assert_equal_up_to_topk_ties(got, want)
            np.testing.assert_array_equal(got, oracle.box_nms(hm, 4, 0.015, keep_top_k=topk))
        else:
            np.testing.assert_array_equal(got, want)
    # the reference's own softmax heatmap + top-k 2048 + the keypoint idiom, in one fused call
    prob = g["softmax34_prob"]
    want = dense_from_sparse(g["softmax34_idx"], g["softmax34_val"], prob.shape)
    dense, kp, sc, cnt = utils.box_nms_keypoints(cu(prob), 4, 0.015, keep_top_k=2048)
    np.testing.assert_array_equal(dense.cpu().numpy(), want)
    assert int(cnt[0]) == len(g["softmax34_kp"])
    np.testing.assert_array_equal(kp[0, :int(cnt[0])].cpu().numpy(), g["softmax34_kp"])
    kpn = g["softmax34_kp"]
    np.testing.assert_array_equal(sc[0, :int(cnt[0])].cpu().numpy(), want[0, 0][kpn[:, 0], kpn[:, 1]])
    np.testing.assert_array_equal(utils.extract_keypoints(dense[0], 0.015).cpu().numpy(), g["softmax34_kp"])


@pytest.mark.parametrize("size,iou", [(3, 0.1), (8, 0.1), (4, 0.3), (5, 0.1), (2.5, 0.1), (12, 0.1)])
def test_box_nms_other_sizes_vs_oracle(utils, oracle, size, iou):
    hm = syn.heatmap(200 + int(size * 2), 2, 96, 136)  # not a multiple of the tile, W % 4 == 0
    got = utils.box_nms(cu(hm), size, 0.015, iou=iou).cpu().numpy()
    np.testing.assert_array_equal(got, oracle.box_nms(hm, size, 0.015, iou=iou))


def test_box_nms_edge_shapes_vs_oracle(utils, oracle):
    for seed, H, W in [(301, 1, 1), (302, 7, 5), (303, 33, 130), (304, 65, 257), (305, 40, 56)]:  # W % 4 != 0: scalar path
        hm = syn.heatmap(seed, 2, H, W)
        for topk in (0, 5):
            got = utils.box_nms(cu(hm), 4, 0.015, keep_top_k=topk).cpu().numpy()
            np.testing.assert_array_equal(got, oracle.box_nms(hm, 4, 0.015, keep_top_k=topk))
    empty = np.zeros((2, 1, 64, 80), np.float32)
    assert float(utils.box_nms(cu(empty), 4, 0.015).abs().sum()) == 0.0
    with pytest.raises(ValueError):
        utils.box_nms(torch.zeros(3, 64, 80, device="cuda"), 4, 0.015)
    with pytest.raises(NotImplementedError):
        utils.box_nms(torch.zeros(64, 80, device="cuda"), 4, -0.5)


def test_box_nms_pathological_maps_vs_oracle(utils, oracle):
    """Flat / constant maps: every pixel is a candidate and (constant map) every score ties, so
    the dependency chain runs across the whole image: exercises the worklist fix-up kernel."""
    flat = np.full((1, 1, 64, 80), 0.02, np.float32)
    np.testing.assert_array_equal(utils.box_nms(cu(flat), 4, 0.015).cpu().numpy(), oracle.box_nms(flat, 4, 0.015))
    ramp = (0.02 + 1e-4 * np.arange(96 * 136, dtype=np.float32).reshape(1, 1, 96, 136)).astype(np.float32)
    np.testing.assert_array_equal(utils.box_nms(cu(ramp), 4, 0.015).cpu().numpy(), oracle.box_nms(ramp, 4, 0.015))
    lg = syn.logits(306, 1, 64, 80, sigma=0.01, bias=0.0)  # random-init-like: ~1/65 everywhere
    prob = oracle.detector_head(lg)
    np.testing.assert_array_equal(utils.box_nms(cu(prob), 4, 0.015).cpu().numpy(), oracle.box_nms(prob, 4, 0.015))


def test_box_nms_bench_size_properties(ops):
    """BASELINE config 2 size (128 images, top-k 2048): size-independent properties."""
    B = 128
    lg = cu(syn.logits(307, B))
    prob = ops.detector_head(lg).reshape(B, 512, 640)
    dense, kp, sc, cnt = ops.box_nms(prob, 4, 0.015, keep_top_k=2048, want_keypoints=True, kp_cap=2048)
    assert int(cnt.min()) == 2048 and int(cnt.max()) == 2048
    assert int((dense > 0).sum()) == 2048 * B
    assert bool(((dense == 0) | (dense == prob)).all())                       # survivors keep their score
    flat = kp[..., 0] * 640 + kp[..., 1]
    assert bool((flat[:, 1:] > flat[:, :-1]).all())                            # row-major, strictly increasing
    assert bool((dense.reshape(B, -1).gather(1, flat) == sc).all())
    # idempotence: survivors are mutually non-suppressing, so NMS of the NMS map is the identity
    again = ops.box_nms(dense, 4, 0.015)
    assert bool((again == dense).all())
    # no two survivors inside each other's footprint: dilating by the footprint never hits another survivor
    m = (dense > 0).float()[:, None]
    k = torch.ones(1, 1, 7, 7, device="cuda")
    k[0, 0, [0, 0, 0, 0, 6, 6, 6, 6, 1, 1, 5, 5], [0, 1, 5, 6, 0, 1, 5, 6, 0, 6, 0, 6]] = 0
    neigh = torch.nn.functional.conv2d(m, k, padding=3)
    assert float((neigh * m).max()) == 1.0                                     # only itself


# ------------------------------------------------------------------ row 5: interpolate_descriptors
def test_interpolate_descriptors(utils, ops, oracle):
    g = load_golden("interpolate")
    for D in (64, 256):
        dm = syn.descriptor_map(int(g["d%d_seed" % D][0]), 1, D, 64, 80)[0]
        kp = g["d%d_kp" % D]
        got = utils.interpolate_descriptors(cu(kp), cu(dm), 512, 640).cpu().numpy()
        np.testing.assert_allclose(got, g["d%d_out" % D], rtol=RTOL, atol=2e-7)                      # vs reference
        np.testing.assert_allclose(got, oracle.interpolate_descriptors(kp, dm, 512, 640), rtol=RTOL, atol=2e-7)
        # channels-last layout gives the same rows
        nhwc = cu(dm).permute(1, 2, 0).contiguous()[None]
        got2 = ops.sample_descriptors(cu(kp)[None], nhwc, 512, 640, channels_last=True)[0].cpu().numpy()
        np.testing.assert_allclose(got2, got, rtol=1e-6, atol=1e-7)
    got = utils.interpolate_descriptors(cu(g["small_kp"]), cu(g["small_in"]), 40, 56).cpu().numpy()
    np.testing.assert_allclose(got, g["small_out"], rtol=RTOL, atol=2e-7)
    kp_before = cu(g["small_kp"])
    kp_copy = kp_before.clone()
    utils.interpolate_descriptors(kp_before, cu(g["small_in"]), 40, 56)
    assert torch.equal(kp_before, kp_copy)                                     # never mutates its input
    assert tuple(utils.interpolate_descriptors(torch.zeros((0, 2), dtype=torch.int64, device="cuda"),
                                               cu(g["small_in"]), 40, 56).shape) == tuple(g["empty_shape"])
    # per-image counts: rows beyond the count are zero
    kp = cu(syn.keypoints(9, 50, 512, 640))[None].repeat(2, 1, 1)
    dm = cu(syn.descriptor_map(9, 2, 64))
    out = ops.sample_descriptors(kp, dm, 512, 640, counts=torch.tensor([50, 20], dtype=torch.int32, device="cuda"))
    assert float(out[1, 20:].abs().sum()) == 0.0 and float(out[1, :20].abs().sum()) > 0
    torch.testing.assert_close(out.norm(dim=2)[0], torch.ones(50, device="cuda"), rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------ rows 6-8: matching
def _golden_inputs(g, case):
    seed, N1, N2, D, noise, dup = case
    tag = "m%d" % int(seed)
    if tag + "_a" in g.files:
        return tag, g[tag + "_a"], g[tag + "_b"]
    return (tag,) + syn.descriptor_sets(int(seed), int(N1), int(N2), int(D), float(noise), int(dup))


@pytest.mark.parametrize("algo", ["simt", "tensor"])
def test_matching_vs_reference_goldens(ops, algo):
    """bfmatcher crossCheck / plain, nnmatcher (two thresholds) and the knn ratio test on the
    fixtures produced by the reference's get_matches: match indices bit-exact, distances 1e-5
    (2e-4 absolute for sqrt(2-2s), which amplifies the fp32 rounding of s near 1)."""
    g = load_golden("matching")
    for case in g["cases"]:
        tag, a, b = _golden_inputs(g, case)
        A, Bm = cu(a), cu(b)

        def run(**kw):
            q, t, d, c = ops.match(A, Bm, algo=algo, **kw)
            n = int(c[0])
            return q[0, :n].cpu().numpy(), t[0, :n].cpu().numpy(), d[0, :n].cpu().numpy()

        q, t, d = run(metric='l2', kind='mutual', cross_check=True)
        np.testing.assert_array_equal(q, g[tag + "_bf_q"]); np.testing.assert_array_equal(t, g[tag + "_bf_t"])
        np.testing.assert_allclose(d, g[tag + "_bf_d"], rtol=RTOL, atol=1e-6)
        q, t, d = run(metric='l2', kind='mutual', cross_check=False)
        np.testing.assert_array_equal(q, g[tag + "_bfnc_q"]); np.testing.assert_array_equal(t, g[tag + "_bfnc_t"])
        for thr, key in [(0.7, "_nn"), (1.1, "_nn11")]:
            q, t, d = run(metric='nn', kind='mutual', cross_check=True, threshold=thr)
            np.testing.assert_array_equal(q, g[tag + key + "_q"]); np.testing.assert_array_equal(t, g[tag + key + "_t"])
            np.testing.assert_allclose(d, g[tag + key + "_d"], rtol=RTOL, atol=2e-4)
        if tag + "_knn_q" in g.files:
            q, t, d = run(metric='l2', kind='ratio', ratio=0.9)
            np.testing.assert_array_equal(q, g[tag + "_knn_q"]); np.testing.assert_array_equal(t, g[tag + "_knn_t"])


@pytest.mark.parametrize("algo", ["simt", "tensor"])
@pytest.mark.parametrize("metric", ["nn", "l2"])
def test_nearest_equals_fp64_truth(ops, oracle, algo, metric):
    """The index contract: argmin equals the fp64 argmin of the oracle on every row, both
    directions, including planted exact duplicates (lowest index) and near-ties."""
    mode = {'nn': 'nn', 'l2': 'bf'}[metric]
    for seed, N1, N2, D, noise, dup in [(401, 500, 700, 64, 0.05, 25), (402, 1000, 900, 256, 0.6, 0),
                                        (403, 130, 257, 128, 0.2, 10), (404, 3, 1, 64, 0.1, 0)]:
        a, b = syn.descriptor_sets(seed, N1, N2, D, noise, dup)
        if seed == 402:  # near-ties: rows that differ from a neighbour by ~1e-7
            b[1::2] = b[0:-1:2] + np.float32(3e-8) * np.sign(b[0:-1:2])
        res = ops.nearest(cu(a), cu(b), metric=metric, algo=algo)
        want = oracle.nearest(a, b, mode, f64=True)
        np.testing.assert_array_equal(res["idx12"][0].cpu().numpy(), want["idx12"])
        np.testing.assert_array_equal(res["idx21"][0].cpu().numpy(), want["idx21"])
        # approximate similarities stay inside the documented error bound (4e-5 |a||b|)
        sim = a.astype(np.float64) @ b.astype(np.float64).T
        best = np.take_along_axis(sim, res["idx12"][0].cpu().numpy().astype(np.int64)[:, None], 1)[:, 0]
        assert np.abs(res["best12"][0].cpu().numpy() - best).max() < 4e-5


@pytest.mark.parametrize("algo", ["simt", "tensor"])
def test_nearest_with_heavy_exact_ties(ops, oracle, algo):
    """Every descriptor of set 2 appears four times (exact 4-way ties for every query, the lowest
    index must win) and a few only twice: exercises the full-rescan kernels (many rows per group,
    and few rows per group) and the two-candidate exact recheck."""
    a, b0 = syn.descriptor_sets(411, 200, 60, 64, 0.05)
    b = np.concatenate([b0, b0, b0, b0])                       # 240 rows: 4 copies each
    a2, c0 = syn.descriptor_sets(412, 150, 150, 64, 0.05)
    c = c0.copy(); c[100:110] = c0[20:30]                      # ten 2-way ties
    d = c0.copy(); d[100:103] = c0[5:8]; d[120:123] = c0[5:8]  # three 3-way ties: few full-list rows
    for x, y in ((a, b), (a2, c), (a2, d)):
        for metric, mode in (("nn", "nn"), ("l2", "bf")):
            res = ops.nearest(cu(x), cu(y), metric=metric, algo=algo)
            want = oracle.nearest(x, y, mode, f64=True)
            np.testing.assert_array_equal(res["idx12"][0].cpu().numpy(), want["idx12"])
            np.testing.assert_array_equal(res["idx21"][0].cpu().numpy(), want["idx21"])


def test_matching_batched_counts_and_empty(ops, oracle):
    """P pairs in one call with per-pair valid counts (the pipeline's layout), vs per-pair oracle."""
    P, N, D = 3, 384, 64
    sets = [syn.descriptor_sets(500 + p, N, N, D, 0.3) for p in range(P)]
    A = cu(np.stack([s[0] for s in sets]))
    Bm = cu(np.stack([s[1] for s in sets]))
    n1 = torch.tensor([384, 100, 0], dtype=torch.int32, device="cuda")
    n2 = torch.tensor([384, 257, 50], dtype=torch.int32, device="cuda")
    for algo in ("simt", "tensor"):
        q, t, d, c = ops.match(A, Bm, metric='l2', algo=algo, kind='mutual', cross_check=True, n1=n1, n2=n2)
        for p in range(P):
            a, b = sets[p][0][:int(n1[p])], sets[p][1][:int(n2[p])]
            wq, wt, wd = oracle.match_mutual(a, b, 'bf', f64=True, cross_check=True) if len(a) else (np.zeros(0),) * 3
            n = int(c[p])
            assert n == len(wq)
            np.testing.assert_array_equal(q[p, :n].cpu().numpy(), wq)
            np.testing.assert_array_equal(t[p, :n].cpu().numpy(), wt)
            np.testing.assert_allclose(d[p, :n].cpu().numpy(), wd, rtol=RTOL, atol=1e-6)


def test_matching_bench_sizes_properties(ops):
    """Config 3 sizes (up to 16k x 16k x 256): planted permutation recovered; mutual property."""
    for N in (2048, 16384):
        a, b = syn.descriptor_sets(600 + N, N, N, 256, 0.05)
        A, Bm = cu(a), cu(b)
        res = ops.nearest(A, Bm, metric='nn', algo='tensor', want_scores=False)
        i12, i21 = res["idx12"][0].long(), res["idx21"][0].long()
        assert bool((i21[i12] == torch.arange(N, device="cuda")).all())      # planted matches are mutual
        sim = (A[:512] @ Bm.T)                                               # fp32 check on a slice
        assert bool((sim.argmax(1) == i12[:512]).all())
        q, t, d, c = ops.match(A, Bm, metric='l2', algo='tensor', kind='mutual', cross_check=True)
        assert int(c[0]) == N and bool((q[0] == torch.arange(N, device="cuda")).all())


def test_column_side_equals_row_side_of_the_swapped_problem(ops):
    """The fused matcher takes idx21 (argmin over axis 0, matching.py:58-59) from the SAME GEMM as idx12, through a
    different code path (warp-collective reductions per 32-row chunk + the chunk-merge kernel) than the in-thread row
    side.  Swapping the operands exchanges the two paths, so at BASELINE sizes, ragged counts, both metrics and both
    descriptor sizes: nearest(a, b).idx21 == nearest(b, a).idx12 and vice versa, index for index (both are the exact
    fp64 argmin with ties to the lowest index); the approximate best similarities agree within the error bound."""
    g = torch.Generator(device="cuda").manual_seed(77)
    for P, N1, N2, D in ((8, 2048, 1900, 256), (8, 2048, 2048, 64), (1, 16384, 16000, 256), (2, 333, 4097, 128)):
        a = torch.nn.functional.normalize(torch.randn((P, N1, D), generator=g, device="cuda"), dim=2)
        b = torch.nn.functional.normalize(torch.randn((P, N2, D), generator=g, device="cuda"), dim=2)
        m = min(N1, N2)
        b[:, :m] = torch.nn.functional.normalize(a[:, torch.randperm(N1, generator=g, device="cuda")[:m]]
                                                 + 0.3 * torch.randn((P, m, D), generator=g, device="cuda") / D ** 0.5, dim=2)
        b[:, 5] = b[:, 4]                                     # exact duplicates: ties to the lowest index on both sides
        n1 = torch.full((P,), N1, dtype=torch.int32, device="cuda"); n1[-1] = N1 - 37
        n2 = torch.full((P,), N2, dtype=torch.int32, device="cuda"); n2[0] = N2 - 129
        for metric in ("l2", "nn"):
            ab = ops.nearest(a, b, metric=metric, algo="tensor", n1=n1, n2=n2)
            ba = ops.nearest(b, a, metric=metric, algo="tensor", n1=n2, n2=n1)
            for p in range(P):
                k1, k2 = int(n1[p]), int(n2[p])
                assert torch.equal(ab["idx21"][p, :k2], ba["idx12"][p, :k2]), (P, N1, N2, D, metric, p)
                assert torch.equal(ab["idx12"][p, :k1], ba["idx21"][p, :k1]), (P, N1, N2, D, metric, p)
                assert float((ab["best21"][p, :k2] - ba["best12"][p, :k2]).abs().max()) < 1e-4


def test_sampler_split_operands_equal_the_matchers_own_prep(ops):
    """mp_sample_descriptors_split_f32 also writes the rows as the tensor-core matcher consumes them; matching from
    those operands (mp_match_split_f32, what KeypointPipeline does) gives the same match list as mp_match_f32, which
    splits the fp32 rows itself."""
    g = torch.Generator(device="cuda").manual_seed(5)
    for D in (64, 256):
        B, K, Hc, Wc, H, W = 6, 700, 32, 40, 256, 320
        desc = torch.nn.functional.normalize(torch.randn((B, Hc, Wc, D), generator=g, device="cuda"), dim=3)
        kp = torch.stack([torch.randint(0, H, (B, K), generator=g, device="cuda"), torch.randint(0, W, (B, K), generator=g, device="cuda")], dim=2)
        cnt = torch.tensor([700, 650, 0, 700, 33, 512], dtype=torch.int32, device="cuda")
        plain = ops.sample_descriptors(kp, desc, H, W, counts=cnt, channels_last=True)
        out, sp = ops.sample_descriptors(kp, desc, H, W, counts=cnt, channels_last=True, split=True)
        assert torch.equal(out, plain)
        hi = out.to(torch.bfloat16)
        assert torch.equal(sp['hi'], hi) and torch.equal(sp['mid'], (out - hi.float()).to(torch.bfloat16))
        torch.testing.assert_close(sp['sq_norms'], (out * out).sum(2), rtol=1e-6, atol=1e-7)
        P = B // 2
        half = lambda d, lo: {k: v[lo:lo + P] for k, v in d.items()}   # noqa: E731
        for metric, thr in (("l2", -1.0), ("nn", 0.9)):
            want = ops.match(out[:P], out[P:], metric=metric, kind='mutual', cross_check=True, threshold=thr, n1=cnt[:P], n2=cnt[P:])
            got = ops.match(out[:P], out[P:], metric=metric, kind='mutual', cross_check=True, threshold=thr, n1=cnt[:P], n2=cnt[P:],
                            split1=half(sp, 0), split2=half(sp, P))
            assert torch.equal(got[3], want[3])
            for p in range(P):
                n = int(want[3][p])
                assert torch.equal(got[0][p, :n], want[0][p, :n]) and torch.equal(got[1][p, :n], want[1][p, :n])
                assert torch.equal(got[2][p, :n], want[2][p, :n])


def test_get_matches_dropin(utils):
    import cv2
    g = load_golden("matching")
    a, b = g["m51_a"], g["m51_b"]
    m = utils.get_matches(a, b, 'bfmatcher', False, crossCheck=True)
    assert isinstance(m[0], cv2.DMatch)
    assert [x.queryIdx for x in m] == list(g["m51_bf_q"]) and [x.trainIdx for x in m] == list(g["m51_bf_t"])
    np.testing.assert_allclose([x.distance for x in m], g["m51_bf_d"], rtol=RTOL, atol=1e-6)
    m = utils.get_matches(a, b, 'nnmatcher', False)
    assert [x.trainIdx for x in m] == list(g["m51_nn_t"])
    m = utils.get_matches(a, b, 'bfmatcher', True)
    assert [x.queryIdx for x in m] == list(g["m51_knn_q"])
    m = utils.get_matches(a, b, 'thresholdmatcher', False, threshold=0.9)
    assert [x.queryIdx for x in m] == list(g["m51_thr_q"]) and [x.trainIdx for x in m] == list(g["m51_thr_t"])
    np.testing.assert_allclose([x.distance for x in m], g["m51_thr_d"], rtol=RTOL, atol=2e-4)
    a4, b4 = g["m54_a"], g["m54_b"]   # planted exact duplicates
    m = utils.get_matches(a4, b4, 'thresholdmatcher', False, threshold=0.9)
    assert [x.trainIdx for x in m] == list(g["m54_thr_t"])
    with pytest.raises(ValueError, match=str(g["err_unknown"])):
        utils.get_matches(a, b, 'nope')
    with pytest.raises(ValueError, match="non-negative"):
        utils.NNMatcher(threshold=-1.0)
    with pytest.raises(AttributeError):
        utils.get_matches(a, b, 'nnmatcher', True)
    with pytest.raises(ValueError, match="not enough values to unpack"):   # knnMatch(k=2) against one descriptor, matching.py:24
        utils.get_matches(a, b[:1], 'bfmatcher', True)
    assert utils.get_matches(np.zeros((0, 64), np.float32), b, 'nnmatcher') == []
    assert utils.get_matches(np.zeros((0, 64), np.float32), b, 'bfmatcher', crossCheck=True) == []


# ------------------------------------------------------------------ rows 9-10: warp / adaptation
def test_warp_vs_oracle_and_restatement(ops, utils, oracle):
    g = load_golden("adaptation")
    H, W = g["img_o"].shape[-2:]
    img = cu(g["img_o"][:, 0])
    for i in range(len(g["H"])):
        for A_np, mode, pad in [(g["A_warp"][i], 'bilinear', 'reflection'), (g["A_unwarp"][i], 'bilinear', 'zeros'),
                                (g["A_unwarp"][i], 'nearest', 'zeros'), (g["A_warp"][i], 'nearest', 'reflection')]:
            got = ops.warp(img, cu(A_np)[None], mode, pad)[0].cpu().numpy()
            want = oracle.warp(g["img_o"][:, 0], A_np, mode, pad)
            # same matrices, same un-fused fp32 op order: the kernel reproduces the oracle to 1 ulp
            np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7)
    # against the torch restatement of the kornia path (different op order: 1e-5 px sampling noise)
    M = torch.from_numpy(g["H"][0].astype(np.float32))[None].repeat(2, 1, 1)
    got = utils.warp_perspective_tensor(cu(g["img_o"]), M.cuda(), (H, W), 'bilinear', 'reflection').cpu().numpy()
    np.testing.assert_allclose(got, g["warp_bilinear_reflection"], rtol=1e-5, atol=2e-5)
    got = utils.warp_perspective_tensor(cu(g["img_o"]), torch.inverse(M).cuda(), (H, W), 'bilinear', 'zeros').cpu().numpy()
    np.testing.assert_allclose(got, g["warp_bilinear_zeros"], rtol=1e-5, atol=2e-5)
    # linspace tables are bit-identical to the oracle's restatement
    xs, ys = ops.linspace_tables(H, W, "cuda")
    np.testing.assert_array_equal(xs.cpu().numpy(), oracle.linspace_table(W))
    np.testing.assert_array_equal(ys.cpu().numpy(), oracle.linspace_table(H))


def test_warp_gather_array_path_is_bit_identical(ops, oracle):
    """Four or more matrices per call stage the planes in a CUDA array and fetch interior footprints with tex2Dgather
    (homographic.cu: warp_kernel<true>); one matrix per call takes the direct gathers.  Same bits, and the oracle's."""
    g = load_golden("adaptation")
    img = cu(g["img_o"][:, 0])
    for A_all, pad in [(g["A_warp"], 'reflection'), (g["A_unwarp"], 'zeros')]:
        A = cu(np.ascontiguousarray(A_all))
        assert A.shape[0] >= 4
        got = ops.warp(img, A, 'bilinear', pad)
        for i in range(A.shape[0]):
            one = ops.warp(img, A[i:i + 1], 'bilinear', pad)[0]
            assert torch.equal(got[i], one), (pad, i)
        np.testing.assert_allclose(got[0].cpu().numpy(), oracle.warp(g["img_o"][:, 0], A_all[0], 'bilinear', pad), rtol=1e-6, atol=1e-7)
    # full-size planes, the reference's homography distribution (rotations up to 180 degrees), NaN-free and bit-equal
    from multipoint_b200 import utils as U
    np.random.seed(11)
    H, W = 512, 640
    cfg = U._check_ha_config({'num': 9})
    Hs, _ = U.sample_adaptation_homographies((H, W), cfg, with_masks=False)
    A_w = U.normalized_warp_matrix(torch.from_numpy(Hs), (H, W), (H, W)).cuda()
    src = torch.rand((3, H, W), device="cuda")
    got = ops.warp(src, A_w, 'bilinear', 'reflection')
    for i in range(A_w.shape[0]):
        assert torch.equal(got[i], ops.warp(src, A_w[i:i + 1], 'bilinear', 'reflection')[0]), i
    # both spectra of a pair in one launch: groups of planes with an output block each
    src4 = torch.rand((4, H, W), device="cuda")
    for mode, pad in (('bilinear', 'reflection'), ('nearest', 'zeros')):
        both = ops.warp(src4, A_w, mode, pad, groups=2)
        assert both.shape == (2, A_w.shape[0], 2, H, W)
        assert torch.equal(both[0], ops.warp(src4[:2], A_w, mode, pad)) and torch.equal(both[1], ops.warp(src4[2:], A_w, mode, pad))
    with pytest.raises(ValueError):
        ops.warp(src, A_w, groups=2)
    # the cached gather array is reused by later calls of the same width: new contents must be seen (no stale texels)
    for seed in range(4):
        fresh = torch.rand((3, H, W), device="cuda", generator=torch.Generator(device="cuda").manual_seed(100 + seed))
        got = ops.warp(fresh, A_w, 'bilinear', 'zeros')
        assert torch.equal(got[2], ops.warp(fresh, A_w[2:3], 'bilinear', 'zeros')[0]), seed


def _stub(g):
    conv = torch.nn.Conv2d(1, 65, 8, stride=8).cuda()
    conv.weight.data = cu(g["stub_w"])
    conv.bias.data = cu(g["stub_b"])

    def net(data):
        from multipoint_b200 import ops as _ops
        with torch.no_grad():
            return {'prob': _ops.detector_head(conv(data['image']).float())}
    return net


def assert_close_and_flip_free(got, want, tag, rtol=1e-4, atol=2e-6, outliers=3e-3, peak=5e-3):
    """Against the torch restatement of the (absent, unpinned) kornia warper.  Its grid is a fused fp32 matmul, the
    kernel's the un-fused oracle order: sampling positions differ by ~1e-5 px, which the stub network's 8x8 convolution
    amplifies to <= ~1.5e-3 relative on a few per cent of the pixels (measured: profiles/r2_adaptation_parity.txt).
    What must NOT happen is a flipped NEAREST mask sample: that changes count by one, i.e. the pixel by >= 1/num = 17 %.
    So: no pixel may be off by more than `peak` (flip-free, counted explicitly), and all but `outliers` are within rtol."""
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-3 * np.abs(want).max())
    flips = int((rel > 0.05).sum())
    assert flips == 0, "%s: %d pixels look like nearest-mask tie flips" % (tag, flips)
    assert rel.max() < peak, (tag, float(rel.max()))
    bad = np.abs(got - want) > atol + rtol * np.abs(want)
    assert bad.mean() <= outliers, (tag, float(bad.mean()))


def test_homographic_adaptation_vs_restatement_and_oracle(utils, ops, oracle):
    g = load_golden("adaptation")
    net = _stub(g)
    cfg = dict(num=6, min_count=2, erosion_radius=3, filter_size=0)
    masks = (g["masks"] != 0).astype(np.uint8)
    img_o, img_t = cu(g["img_o"]), cu(g["img_t"])
    data = {'optical': {'image': img_o, 'is_optical': torch.ones(2, 1, dtype=torch.bool, device="cuda")},
            'thermal': {'image': img_t, 'is_optical': torch.zeros(2, 1, dtype=torch.bool, device="cuda")}}
    fixture_mats = (g["A_warp"], g["A_unwarp"])
    for fs, suffix in ((0, ""), (5, "_f5")):       # filter_size=5: the Gaussian branch, homographies.py:54-58,100-102
        c = dict(cfg, filter_size=fs)
        for mats in (None, fixture_mats):          # own 3x3 algebra, and the fixture's own normalised matrices
            out = utils.homographic_adaptation({'image': img_o}, net, dict(c), homographies=g["H"], masks=masks, normalized_matrices=mats)
            assert_close_and_flip_free(out.cpu().numpy(), g["single" + suffix], "single" + suffix)
            for agg in ("prod", "sum") if fs == 0 else ("prod",):
                out = utils.homographic_adaptation_multispectral(data, net, dict(c, aggregation=agg), homographies=g["H"], masks=masks,
                                                                 normalized_matrices=mats)
                assert_close_and_flip_free(out.cpu().numpy(), g["multi_" + agg + suffix], "multi_" + agg + suffix)
    # end to end against the C oracle's chain on the fixture's matrices (same un-fused arithmetic, same stub network on the
    # GPU): warp -> net -> unwarp/aggregate/finish must agree to fp32 rounding, which pins the sample order, the masks, the
    # spectra combination and the finish of the whole entry point, not just the kernels one by one
    n, B, H, W = len(g["H"]), 2, 64, 80

    def heat(imgs_np):   # (n*B,H,W) -> (n,B,H,W) heatmaps of the stub network
        return net({'image': cu(imgs_np.reshape(-1, 1, H, W))})['prob'][:, 0].reshape(-1, B, H, W).cpu().numpy()

    warped_o = np.stack([oracle.warp(g["img_o"][:, 0], g["A_warp"][i], 'bilinear', 'reflection') for i in range(n)])
    warped_t = np.stack([oracle.warp(g["img_t"][:, 0], g["A_warp"][i], 'bilinear', 'reflection') for i in range(n)])
    p0_o, p0_t = heat(g["img_o"][:, 0])[0], heat(g["img_t"][:, 0])[0]
    pw_o, pw_t = heat(warped_o), heat(warped_t)
    want, _ = oracle.ha_aggregate(p0_o, pw_o, None, masks.astype(np.float32), g["A_unwarp"], "none", 2)
    got = utils.homographic_adaptation({'image': img_o}, net, dict(cfg), homographies=g["H"], masks=masks, normalized_matrices=fixture_mats)
    np.testing.assert_allclose(got[:, 0].cpu().numpy(), want, rtol=1e-5, atol=1e-7)
    for agg in ("prod", "sum"):
        p0 = p0_o * p0_t if agg == "prod" else p0_o + p0_t
        want, _ = oracle.ha_aggregate(p0, pw_o, pw_t, masks.astype(np.float32), g["A_unwarp"], agg, 2)
        got = utils.homographic_adaptation_multispectral(data, net, dict(cfg, aggregation=agg), homographies=g["H"], masks=masks,
                                                         normalized_matrices=fixture_mats)
        np.testing.assert_allclose(got[:, 0].cpu().numpy(), want, rtol=1e-5, atol=1e-7)
    # chunked samples (O(chunk*B) memory) carry the accumulators in the same order: bit-identical to one chunk
    one = utils.homographic_adaptation_multispectral(data, net, dict(cfg, aggregation="prod", sample_chunk=64), homographies=g["H"], masks=masks)
    two = utils.homographic_adaptation_multispectral(data, net, dict(cfg, aggregation="prod", sample_chunk=2), homographies=g["H"], masks=masks)
    assert torch.equal(one, two)
    # the Gaussian filter itself (utils.py:124-157) on the device
    f5 = utils.get_gaussian_filter(5).cuda()
    np.testing.assert_array_equal(f5.weight.detach().cpu().numpy(), g["gauss_w_5"])
    blurred = f5(torch.nn.ReflectionPad2d(2)(cu(g["gauss_in"]))).detach().cpu().numpy()
    np.testing.assert_allclose(blurred, g["gauss_out_5"], rtol=1e-5, atol=1e-8)
    # the aggregate kernel alone against the C oracle on identical inputs: tight
    rng = np.random.default_rng(3)
    n, B, H, W = 5, 2, 64, 80
    pa = rng.random((n, B, H, W), dtype=np.float32) * 0.3
    pb = rng.random((n, B, H, W), dtype=np.float32) * 0.3
    p0 = rng.random((B, H, W), dtype=np.float32) * 0.1
    for agg, second in [("none", None), ("prod", pb), ("sum", pb)]:
        want, wcount = oracle.ha_aggregate(p0, pa, second, masks.astype(np.float32), g["A_unwarp"], agg, 2)
        got = ops.ha_aggregate(cu(p0), cu(pa), None if second is None else cu(second), cu(masks), cu(g["A_unwarp"]), agg, 2)
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-6, atol=1e-7)
        # the TMA-staged variant (source windows in shared memory) takes the same taps in the same order: same bits
        got_st = ops.ha_aggregate(cu(p0), cu(pa), None if second is None else cu(second), cu(masks), cu(g["A_unwarp"]), agg, 2, staged=True)
        assert torch.equal(got_st, got)
        # split into two partial accumulations (the multi-GPU path) and finish: same result to fp32 rounding
        acc = ops.ha_aggregate(cu(p0), cu(pa[:3]), None if second is None else cu(second[:3]), cu(masks[:3]),
                               cu(g["A_unwarp"][:3]), agg, 2, init=True, finish=False)
        got2 = ops.ha_aggregate(None, cu(pa[3:]), None if second is None else cu(second[3:]), cu(masks[3:]),
                                cu(g["A_unwarp"][3:]), agg, 2, init=False, finish=True, prob_acc=acc[0], count_acc=acc[1])
        np.testing.assert_allclose(got2.cpu().numpy(), want, rtol=1e-6, atol=1e-7)
    # error behaviour (homographies.py:42-46,68,123)
    with pytest.raises(ValueError, match="num must be larger than 0"):
        utils.homographic_adaptation({'image': img_o}, net, dict(cfg, num=0))
    with pytest.raises(ValueError, match="filter_size must be uneven"):
        utils.homographic_adaptation({'image': img_o}, net, dict(cfg, filter_size=4))
    with pytest.raises(ValueError, match="Unknown aggregation"):
        utils.homographic_adaptation_multispectral(data, net, dict(cfg, num=2, aggregation='max'))


def test_valid_masks_device(utils, ops, oracle):
    """batched device-side compute_valid_mask, bit-exact against the reference's masks
    (fixture), the oracle, and the host cv2 path."""
    g = load_golden("homographies")
    for tag in ("default", "export", "noart", "small"):
        shape = tuple(int(v) for v in g[tag + "_shape"])
        erosion = int(g[tag + "_seed"][1])
        want = np.unpackbits(g[tag + "_mask"], axis=-1)[..., :shape[1]]
        got = utils.compute_valid_masks(shape, g[tag + "_H"], erosion, True).cpu().numpy()
        np.testing.assert_array_equal(got, want)
    Hm = g["small_H"][:1]
    np.testing.assert_array_equal(utils.compute_valid_masks((64, 80), Hm, 0, False).cpu().numpy()[0],
                                  np.unpackbits(g["small_mask_e0"], axis=-1)[..., :80])
    np.testing.assert_array_equal(utils.compute_valid_masks((64, 80), Hm, 2, False).cpu().numpy()[0],
                                  np.unpackbits(g["small_mask_e2nb"], axis=-1)[..., :80])
    # ragged sizes (W not a multiple of 4 / 32, H not a multiple of the row tile), every erosion mode, vs the oracle
    rng = np.random.default_rng(11)
    for shape in [(37, 53), (12, 20), (70, 131), (100, 64)]:
        Hs = []
        for _ in range(7):
            th = rng.uniform(-0.5, 0.5)
            Hs.append(np.array([[np.cos(th) * rng.uniform(0.8, 1.2), -np.sin(th), rng.uniform(-8, 8)],
                                [np.sin(th), np.cos(th) * rng.uniform(0.8, 1.2), rng.uniform(-8, 8)],
                                [rng.uniform(-2e-3, 2e-3), rng.uniform(-2e-3, 2e-3), 1.0]]))
        Hs.append(np.eye(3))
        Hs = np.stack(Hs)
        for r, border in [(0, False), (1, True), (4, False), (31, True)]:
            got = utils.compute_valid_masks(shape, Hs, r, border).cpu().numpy()
            for i, Hm in enumerate(Hs):
                np.testing.assert_array_equal(got[i], oracle.valid_mask(shape, Hm, r, border), err_msg=str((shape, r, border, i)))
                if r <= 4:
                    np.testing.assert_array_equal(got[i], utils.compute_valid_mask(shape, Hm, r, border).astype(np.uint8))
    assert utils.compute_valid_masks((16, 16), np.zeros((0, 3, 3))).shape == (0, 16, 16)
    with pytest.raises(NotImplementedError):
        utils.compute_valid_masks((64, 80), np.eye(3)[None], 40, True)
    # bench size: 99 homographies of 512x640 in one launch == the host loop on a sample of them
    np.random.seed(4)
    cfg = utils._check_ha_config({})
    Hs, _ = utils.sample_adaptation_homographies((512, 640), cfg, with_masks=False)
    got = utils.compute_valid_masks((512, 640), Hs, cfg['erosion_radius'], cfg['mask_border'])
    assert got.shape == (99, 512, 640)
    for i in (0, 17, 98):
        np.testing.assert_array_equal(got[i].cpu().numpy(), utils.compute_valid_mask((512, 640), Hs[i], 5, True).astype(np.uint8))


def test_homographic_adaptation_host_sampling_matches_reference_stream(utils):
    """With np.random.seed(5) the product draws the same homographies and masks as the reference run
    that produced the fixture, so the end-to-end call (no injected samples) matches it too."""
    g = load_golden("adaptation")
    net = _stub(g)
    cfg = dict(num=6, aggregation='prod', erosion_radius=3, mask_border=True, min_count=2, filter_size=0,
               homographies=dict(translation=True, rotation=True, scaling=True, perspective=True, scaling_amplitude=0.2,
                                 perspective_amplitude_x=0.2, perspective_amplitude_y=0.2, patch_ratio=0.85,
                                 max_angle=1.57, allow_artifacts=True))
    np.random.seed(int(g["seed"][0]))
    out = utils.homographic_adaptation({'image': cu(g["img_o"])}, net, cfg)
    assert_close_and_flip_free(out.cpu().numpy(), g["single"], "single (host sampling)")


def test_relu_bn_pad_kernel_vs_torch(ops):
    """The fused glue between the backbone convolutions against the module sequence it replaces
    (ReLU -> BatchNorm2d(eval) [-> MaxPool2d] [-> ReflectionPad2d | ZeroPad2d], or BatchNorm first)."""
    torch.manual_seed(5)
    for (B, C, H, W) in [(3, 16, 32, 40), (2, 5, 17, 23), (1, 64, 128, 160)]:
        x = torch.randn(B, C, H, W, device="cuda") * 2
        bn = torch.nn.BatchNorm2d(C).cuda().eval()
        with torch.no_grad():
            bn.weight.uniform_(-1.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.2, 3.0)
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            shift = bn.bias - bn.running_mean * scale
            for bn_first in (False, True):
                for pool in (False, True):
                    if pool and (H % 2 or W % 2):
                        continue
                    for pad, reflect in [(0, True), (1, True), (1, False)]:
                        cbias = torch.randn(C, device="cuda") if (pad and pool) or bn_first else None
                        xb = x if cbias is None else x + cbias[None, :, None, None]
                        ref = torch.relu(bn(xb)) if bn_first else bn(torch.relu(xb))
                        if pool:
                            ref = torch.nn.functional.max_pool2d(ref, 2, 2)
                        if pad:
                            ref = (torch.nn.ReflectionPad2d(1) if reflect else torch.nn.ZeroPad2d(1))(ref)
                        got = ops.relu_bn_pad(x, scale, shift, bn_first=bn_first, pool=pool, pad=pad, reflect=reflect, conv_bias=cbias)
                        torch.testing.assert_close(got, ref, rtol=2e-6, atol=2e-6)
    with pytest.raises(ValueError):
        ops.relu_bn_pad(torch.zeros(1, 2, 5, 6, device="cuda"), torch.ones(2, device="cuda"), torch.zeros(2, device="cuda"), pool=True)


def test_conv1_relu_bn_pad_kernel_vs_torch(ops):
    """The encoders' first layer as one kernel against pad -> cuDNN conv -> ReLU/BatchNorm -> pad."""
    torch.manual_seed(6)
    for (B, C, H, W) in [(2, 64, 64, 80), (3, 7, 17, 23), (1, 64, 128, 640)]:
        img = torch.rand(B, 1, H, W, device="cuda")
        conv = torch.nn.Conv2d(1, C, 3).cuda()
        bn = torch.nn.BatchNorm2d(C).cuda().eval()
        with torch.no_grad():
            bn.weight.uniform_(-1.5, 1.5); bn.bias.normal_(); bn.running_mean.normal_(std=0.1); bn.running_var.uniform_(0.2, 3.0)
            scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            shift = bn.bias - bn.running_mean * scale
            for bn_first in (False, True):
                for in_reflect in (True, False):
                    for pad, out_reflect in [(0, True), (1, True), (1, False)]:
                        y = conv((torch.nn.ReflectionPad2d(1) if in_reflect else torch.nn.ZeroPad2d(1))(img))
                        ref = torch.relu(bn(y)) if bn_first else bn(torch.relu(y))
                        if pad:
                            ref = (torch.nn.ReflectionPad2d(1) if out_reflect else torch.nn.ZeroPad2d(1))(ref)
                        got = ops.conv1_relu_bn_pad(img, conv.weight, conv.bias, scale, shift, bn_first=bn_first, in_reflect=in_reflect,
                                                    pad=pad, out_reflect=out_reflect)
                        torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)


def test_multipoint_fused_inference_path_equals_module_path():
    """MultiPoint in eval / no_grad (fused glue kernels between the cuDNN convolutions) against the same network run
    module by module (what it does whenever autograd is on): same logits / descriptors to fp32 rounding."""
    from multipoint_b200.models import MultiPoint
    from multipoint_b200.pipeline import calibrate_random_init
    for cfg in ({'multispectral': True, 'descriptor_size': 64},
                {'multispectral': False, 'descriptor_size': 32, 'bn_first': True, 'reflection_pad': False},
                {'multispectral': False, 'descriptor_size': 32, 'double_convolution': False}):
        torch.manual_seed(9)
        net = MultiPoint(dict(cfg)).cuda().eval()
        img = torch.rand(4, 1, 64, 80, device="cuda")
        opt = torch.tensor([[1], [0], [1], [0]], dtype=torch.bool, device="cuda")
        calibrate_random_init(net, img, is_optical=opt)
        data = {'image': img, 'is_optical': opt}
        with torch.no_grad():
            xf = net.encode(data)
            fl, fr = net.backbone_outputs(data)
            fused = net(data)
        with torch.enable_grad():
            xm = net.encode(data).detach()
            ml, mr = net.backbone_outputs(data)
        # encoder output: fp32 rounding apart
        torch.testing.assert_close(xf, xm, rtol=2e-5, atol=2e-5 * float(xm.abs().max()))
        # head outputs: the calibrated random-init network is ill-conditioned (BatchNorm subtracts a mean much larger
        # than the spread), so judge both fp32 paths against the same network in float64
        import copy
        net64 = copy.deepcopy(net).double()
        with torch.enable_grad():
            tl, tr = net64.backbone_outputs({'image': img.double(), 'is_optical': opt})
        for f, m, t in ((fl, ml.detach(), tl.detach()), (fr, mr.detach(), tr.detach())):
            err_f = float((f.double() - t).abs().max())
            err_m = float((m.double() - t).abs().max())
            assert err_f <= 3.0 * err_m + 1e-6 * float(t.abs().max()), (err_f, err_m)
        assert fused['prob'].shape == (4, 1, 64, 80) and fused['logits'] is None


# ------------------------------------------------------------------ row 3: model contract
def test_multipoint_forward_vs_reference(oracle):
    from multipoint_b200.models import MultiPoint
    g = load_golden("model")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for tag, cfg in [("shipped", {'multispectral': False, 'descriptor_size': 64, 'bn_first': False,
                                  'descriptor_head': True, 'final_batchnorm': True, 'reflection_pad': True,
                                  'normalize_descriptors': True}),
                     ("multi", {'multispectral': True, 'descriptor_size': 256})]:
        torch.manual_seed(0)
        net = MultiPoint(cfg)
        assert list(net.state_dict().keys()) == list(g[tag + "_keys"])
        assert [str(tuple(v.shape)) for v in net.state_dict().values()] == list(g[tag + "_shapes"])
        net = net.cuda().eval()
        img = syn.images(81, 2, 64, 80)
        data = {'image': cu(img), 'is_optical': torch.tensor([[True], [False]], device="cuda")}
        with torch.no_grad():
            out = net(data)
        assert out['logits'] is None
        # same seeded weights as the reference; cuDNN fp32 vs the reference's CPU convolutions
        np.testing.assert_allclose(out['prob'].cpu().numpy(), g[tag + "_prob"], rtol=2e-3, atol=1e-6)
        np.testing.assert_allclose(out['desc'].cpu().numpy(), g[tag + "_desc"], rtol=2e-3, atol=2e-5)
        net.set_force_return_logits(True)
        with torch.no_grad():
            o2 = net(data)
        assert o2['prob'] is None and o2['logits'].shape == (2, 65, 8, 10)
        # given identical backbone outputs the tails match the oracle to 1e-5
        np.testing.assert_allclose(out['prob'].cpu().numpy(), oracle.detector_head(o2['logits'].cpu().numpy()), rtol=RTOL, atol=1e-9)
        with pytest.raises(ValueError):
            net.set_force_return_logits(1)


def test_pipeline_matches_stagewise_reference_chain(utils, oracle):
    """The fused sync-free pipeline equals the stage-by-stage drop-in calls (what
    predict_align_image_pair.py does) and the oracle chain on the same backbone outputs."""
    from multipoint_b200.pipeline import KeypointPipeline
    B, H, W, D = 2, 128, 160, 64
    lg = syn.logits(701, 2 * B, H // 8, W // 8)
    raw = syn.descriptor_map(702, 2 * B, D, H // 8, W // 8)
    pipe = KeypointPipeline(None, nms=4, detection_threshold=0.015, topk=300)
    ext = pipe.extract_from_backbone(cu(lg), cu(raw), H, W)
    ea = {k: v[:B] for k, v in ext.items()}
    eb = {k: v[B:] for k, v in ext.items()}
    m = pipe.match(ea, eb)
    prob = oracle.detector_head(lg)
    desc = oracle.normalize_descriptors(raw)
    kps, descs = [], []
    for b in range(2 * B):
        nms = oracle.box_nms(ext['prob'][b, 0].cpu().numpy(), 4, 0.015, keep_top_k=300)  # same heatmap bits as the GPU
        kp = oracle.extract_keypoints(nms, 0.015)
        n = int(ext['counts'][b])
        np.testing.assert_array_equal(ext['keypoints'][b, :n].cpu().numpy(), kp)
        np.testing.assert_array_equal(ext['prob_nms'][b, 0].cpu().numpy(), nms)
        d = oracle.interpolate_descriptors(kp, desc[b], H, W)
        np.testing.assert_allclose(ext['desc'][b, :n].cpu().numpy(), d, rtol=RTOL, atol=2e-7)
        kps.append(kp); descs.append(ext['desc'][b, :n].cpu().numpy())
    np.testing.assert_allclose(ext['prob'].cpu().numpy(), prob, rtol=RTOL, atol=1e-9)
    for p in range(B):
        wq, wt, wd = oracle.match_mutual(descs[p], descs[B + p], 'bf', f64=True, cross_check=True)
        n = int(m['counts'][p])
        np.testing.assert_array_equal(m['query'][p, :n].cpu().numpy(), wq)
        np.testing.assert_array_equal(m['train'][p, :n].cpu().numpy(), wt)
        np.testing.assert_allclose(m['distance'][p, :n].cpu().numpy(), wd, rtol=RTOL, atol=1e-6)


def test_pipeline_stream_equals_call():
    """KeypointPipeline.stream (pinned host batches, upload of batch i+1 overlapped with batch i) yields exactly
    what KeypointPipeline.__call__ returns for each batch on its own."""
    from multipoint_b200.models import MultiPoint
    from multipoint_b200.pipeline import KeypointPipeline, calibrate_random_init
    torch.manual_seed(3)
    net = MultiPoint({'multispectral': True, 'descriptor_size': 64}).cuda().eval()
    calibrate_random_init(net, torch.rand(4, 1, 64, 80, device="cuda"), is_optical=torch.tensor([[1], [1], [0], [0]], dtype=torch.bool, device="cuda"))
    pipe = KeypointPipeline(net, nms=4, detection_threshold=0.015, topk=64)
    batches = []
    for seed in (11, 12, 13):
        b = syn.image_pair_batch(seed, 2, 64, 80)
        batches.append({s: {k: torch.from_numpy(v).pin_memory() for k, v in b[s].items() if k != 'valid_mask'} for s in ('optical', 'thermal')})
    streamed = []
    for r in pipe.stream(iter(batches), "cuda"):
        streamed.append({'kp': r['optical']['keypoints'].cpu(), 'cnt': r['thermal']['counts'].cpu(), 'q': r['matches']['query'].cpu(),
                         't': r['matches']['train'].cpu(), 'n': r['matches']['counts'].cpu()})
    assert len(streamed) == 3
    # routing by the dict keys (no host sync) == routing every row by its is_optical flag (the reference's way)
    by_flag = KeypointPipeline(net, nms=4, detection_threshold=0.015, topk=64, trust_spectrum_keys=False)
    dev_b = {s: {k: v.cuda() for k, v in d.items()} for s, d in batches[0].items()}
    ra, rb = pipe(dev_b), by_flag(dev_b)
    assert torch.equal(ra['optical']['keypoints'], rb['optical']['keypoints']) and torch.equal(ra['thermal']['counts'], rb['thermal']['counts'])
    assert torch.equal(ra['matches']['counts'], rb['matches']['counts'])
    for b, got in zip(batches, streamed):
        r = pipe({s: {k: v.cuda() for k, v in d.items()} for s, d in b.items()})
        assert torch.equal(got['kp'], r['optical']['keypoints'].cpu()) and torch.equal(got['cnt'], r['thermal']['counts'].cpu())
        assert torch.equal(got['n'], r['matches']['counts'].cpu())
        for p in range(2):   # entries beyond the count are not defined
            n = int(got['n'][p])
            assert torch.equal(got['q'][p, :n], r['matches']['query'][p, :n].cpu()) and torch.equal(got['t'][p, :n], r['matches']['train'][p, :n].cpu())


# ------------------------------------------------------------------ multi-GPU path, simulated in one process
def test_sharded_adaptation_equals_fused(utils, ops):
    """Two ranks' partial accumulators summed (what the NCCL all-reduce does) then finished equal
    the single-process fused result to fp32 rounding."""
    g = load_golden("adaptation")
    net = _stub(g)
    masks = (g["masks"] != 0).astype(np.uint8)
    img_o, img_t = cu(g["img_o"]), cu(g["img_t"])
    cfg = utils._check_ha_config(dict(num=6, min_count=2, aggregation='prod'))
    opt = torch.ones(2, 1, dtype=torch.bool, device="cuda")
    second = (img_t, ~opt)
    fused = utils._adaptation(img_o, opt, net, cfg, second, g["H"], masks)
    parts = [utils._adaptation_core(img_o, opt, net, cfg, second, g["H"], masks, r, 2, False) for r in range(2)]
    prob_sum, count_sum = parts[0][0] + parts[1][0], parts[0][1] + parts[1][1]
    out = utils.adaptation_finish(prob_sum, count_sum, 'prod', 2)[:, None]
    # not bit-exact: the summation order changes and cuDNN picks batch-size dependent algorithms for
    # the per-rank forward passes; sqrt of the product aggregation amplifies both near zero
    torch.testing.assert_close(out, fused, rtol=2e-4, atol=1e-5)
    # world larger than the number of samples: some ranks only contribute zeros
    parts = [utils._adaptation_core(img_o, opt, net, cfg, second, g["H"], masks, r, 8, False) for r in range(8)]
    out8 = utils.adaptation_finish(sum(p[0] for p in parts), sum(p[1] for p in parts), 'prod', 2)[:, None]
    torch.testing.assert_close(out8, fused, rtol=2e-4, atol=1e-5)


# ------------------------------------------------------------------ entry points
def test_entry_points_run_and_agree(tmp_path, utils):
    import yaml
    from multipoint_b200.scripts import export_keypoints, predict_align_image_pair, predict_keypoints
    model_dir = tmp_path / "model"
    model_dir.mkdir()
    (model_dir / "params.yaml").write_text(yaml.dump({'model': {'type': 'MultiPoint', 'multispectral': False, 'descriptor_size': 64,
                                                                 'descriptor_head': True, 'final_batchnorm': True}}))
    cfg = {'prediction': {'allow_gpu': True, 'batchsize': 2, 'detection_threshold': 0.015, 'nms': 4, 'cpu_nms': True, 'topk': 100,
                          'reprojection_threshold': 3,
                          'matching': {'method': 'bfmatcher', 'method_kwargs': {'crossCheck': True}, 'knn_matches': False},
                          'homographic_adaptation': {'num': 4, 'aggregation': 'sum', 'erosion_radius': 3, 'min_count': 1}}}
    (tmp_path / "cfg.yaml").write_text(yaml.dump(cfg))
    imgs = syn.image_pair_batch(11, 2, 64, 80)
    np.savez(tmp_path / "in.npz", optical=imgs['optical']['image'], thermal=imgs['thermal']['image'], names=np.array(['a', 'b']))
    common = ['-y', str(tmp_path / "cfg.yaml"), '-m', str(model_dir), '-v', 'none', '--input', str(tmp_path / "in.npz")]
    r1 = predict_align_image_pair.main(common + ['-i', '1', '-o', str(tmp_path / "pair.npz")])
    assert r1['keypoints_optical'].shape[1] == 2 and len(r1['query']) == len(r1['train']) == len(r1['distance'])
    assert (tmp_path / "pair.npz").exists()
    r2 = predict_keypoints.main(common + ['-b', '-o', str(tmp_path / "kp.npz")])
    dense = torch.from_numpy(r2['prob_optical'][1, 0])
    np.testing.assert_array_equal(torch.nonzero((dense > 0.015).float()).numpy(), r2['keypoints_optical_1'])
    assert len(r2['keypoints_thermal_0']) == 100                               # top-k 100 of a dense candidate map
    # same seed-0 random-init weights and the same batch size of 1 in both scripts: same keypoints
    torch.manual_seed(0)   # predict_align_image_pair seeds torch itself (-s 0); predict_keypoints does not (like the reference)
    r2s = predict_keypoints.main(common + ['-i', '1'])
    np.testing.assert_array_equal(r2s['keypoints_optical_0'], r1['keypoints_optical'])
    out = tmp_path / "labels.npz"
    r3 = export_keypoints.main(common + ['-o', str(out)])
    z = np.load(out)
    assert sorted(z.files) == ['a/keypoints', 'b/keypoints'] and z['a/keypoints'].dtype == np.int64
    assert export_keypoints.main(common + ['-o', str(out), '-skip']) == {}     # resume: nothing left to do


# ------------------------------------------------------------------ evaluation loops
def test_evaluation_point_kernels(ops, oracle):
    g = load_golden("evaluation")
    # batched: the six fixture problems in one call each, ragged counts
    P, capa, capb = 6, 190, 160
    a = np.zeros((P, capa, 2), np.int64)
    b = np.zeros((P, capb, 2), np.int64)
    Hs = np.zeros((P, 3, 3), np.float64)
    na, nb = [], []
    for i in range(P):
        ai, bi = g["pt%d_a" % i], g["pt%d_b" % i]
        a[i, :len(ai)] = ai
        b[i, :len(bi)] = bi
        Hs[i] = g["pt%d_H" % i]
        na.append(len(ai))
        nb.append(len(bi))
    ca, cb = cu(np.array(na, np.int32)), cu(np.array(nb, np.int32))
    wi = ops.warp_keypoints(cu(a), cu(Hs), ca)
    wf = ops.warp_keypoints(cu(a), cu(Hs), ca, as_int=False)
    d2 = ops.points_min_dist2(wi, cu(b), 128, 160, ca, cb)
    mq = np.tile(np.arange(150, dtype=np.int32), (P, 1))
    row_any, tp = ops.points_correct(wf, cu(b), 4.0, ca, cb, cu(mq), cu(mq), cu(np.full(P, 150, np.int32)))
    for i in range(P):
        n = na[i]
        np.testing.assert_array_equal(wi[i, :n].cpu().numpy(), g["pt%d_warp_int" % i])
        np.testing.assert_array_equal(wf[i, :n].cpu().numpy(), g["pt%d_warp_f64" % i])       # bit-exact doubles vs cv2
        d2i = d2[i, :n].cpu().numpy()
        np.testing.assert_array_equal(d2i, oracle.points_min_dist2(g["pt%d_warp_int" % i], g["pt%d_b" % i], 128, 160))
        np.testing.assert_array_equal(np.sqrt(d2i[d2i >= 0].astype(np.float64)), g["pt%d_min_dist" % i])
        np.testing.assert_array_equal(np.flatnonzero(row_any[i, :n].cpu().numpy()), g["pt%d_correct_rows" % i])
        np.testing.assert_array_equal(tp[i].cpu().numpy().astype(bool), g["pt%d_correct_diag" % i])
        assert int(d2[i, n:].max()) == -1 and int(row_any[i, n:].max()) == 0                 # padding untouched
    # no targets / no queries
    none = ops.points_min_dist2(wi[:1], cu(np.zeros((1, 1, 2), np.int64)), 128, 160, ca[:1], cu(np.zeros(1, np.int32)))
    inside = none[0, :na[0]].cpu().numpy()
    assert set(np.unique(inside)) <= {-1, np.iinfo(np.int64).max}
    # a larger random problem against the oracle
    rng = np.random.default_rng(5)
    q = rng.integers(-20, 660, (1, 3000, 2))
    t = rng.integers(0, 640, (1, 2500, 2))
    got = ops.points_min_dist2(cu(q), cu(t), 512, 640)[0].cpu().numpy()
    np.testing.assert_array_equal(got, oracle.points_min_dist2(q[0], t[0], 512, 640))
    qw = q[0].astype(np.float64) + rng.random((3000, 2))
    ra, _ = ops.points_correct(cu(qw[None]), cu(t), 6.0)
    np.testing.assert_array_equal(ra[0].cpu().numpy(), oracle.points_correct(qw, t[0], 6.0)[0])


def _evaluation_loader(g):
    batches = []
    for bi in range(2):
        b = {}
        for s in ('optical', 'thermal'):
            d = {'image': torch.zeros((2, 1, 128, 160))}
            for k in ('valid_mask', 'stub_prob', 'stub_desc', 'homography'):
                key = "in%d_%s_%s" % (bi, s, k)
                if key in g.files:
                    d[k] = torch.from_numpy(g[key].copy())
            b[s] = d
        batches.append(b)
    return batches


def test_evaluation_loops_match_reference():
    """compute_repeatability_multispectral / compute_descriptor_metrics on canned network outputs: every returned
    value equals what the reference's loops returned on the CPU (tests/golden/evaluation.npz)."""
    from multipoint_b200 import evaluation
    g = load_golden("evaluation")

    def net(d):
        return {'prob': d['stub_prob'].clone(), 'desc': d['stub_desc'].clone()}

    for tag, topk in (("rep_top0", 0), ("rep_top150", 150)):
        cfg = {'prediction': {'detection_threshold': 0.015, 'nms': 4, 'topk': topk, 'cpu_nms': True}}
        mean, rep, nko, nkt = evaluation.compute_repeatability_multispectral(net, _evaluation_loader(g), 'cuda', cfg, distance_thresh=3)
        np.testing.assert_array_equal(np.array([mean] + list(rep)), g[tag])
        np.testing.assert_array_equal(np.array([nko, nkt]), g[tag + "_nkp"])
    for tag, method, kwargs, topk in (("desc_bf", "bfmatcher", {'crossCheck': True}, 0), ("desc_nn", "nnmatcher", {'threshold': 0.9}, 200)):
        cfg = {'detection_threshold': 0.015, 'nms': 4, 'topk': topk, 'cpu_nms': True, 'reprojection_threshold': 3,
               'matching': {'method': method, 'knn_matches': False, 'method_kwargs': kwargs}}
        res = evaluation.compute_descriptor_metrics(net, _evaluation_loader(g), 'cuda', cfg, threshold_keypoints=4, threshold_warp=4)
        assert sorted(res.keys()) == sorted(k[len(tag) + 1:] for k in g.files if k.startswith(tag + "_"))
        for k, v in res.items():
            want = g[tag + "_" + k]
            if k.startswith(('tp_', 'fp_')):
                np.testing.assert_array_equal(np.asarray(v), want, err_msg=k)
            elif k in ('pts_dist', 'average_h_error'):
                np.testing.assert_allclose(np.asarray(v), want, rtol=1e-6, err_msg=k)     # RANSAC + LM refinement on the host
            else:
                np.testing.assert_allclose(np.asarray(v), want, rtol=1e-6, atol=1e-7, err_msg=k)
