"""CPU-only checks: the C-ABI library loads and exports every symbol include/multipoint_b200.h
declares, the host-side mirror of the reference interface (homography sampling, masks, model
layout, error behaviour) matches the fixtures frozen from the reference, and the product path
fails loudly without a CUDA device instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from multipoint_b200 import _lib
from multipoint_b200 import synthetic as syn


def test_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "multipoint_b200.h")).read()
    declared = set(re.findall(r"MP_API\s+[\w\s\*]+?\b(mp_\w+)\s*\(", header))
    assert len(declared) >= 17
    lib = _lib.load()                                  # loads without a GPU
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.mp_version() == 100
    # size queries are pure host functions
    assert lib.mp_box_nms_workspace_bytes(1, 512, 640) >= 512 * 640 * 12
    assert lib.mp_match_workspace_bytes(1, 2048, 2048, 256) >= 2 * 2 * 2048 * 256 * 2
    assert lib.mp_extract_keypoints_workspace_bytes(2, 512, 640) >= 2 * 512 * 640 // 8


def test_abi_argument_errors_without_gpu():
    lib = _lib.load()
    rc = lib.mp_detector_head_f32(None, 1, 8, 8, None, None, None)
    assert rc == _lib.MP_ERR_INVALID and "null" in _lib.last_error()
    rc = lib.mp_box_nms_f32(None, 1, 8, 8, 4.0, -1.0, 0.1, 0, None, None, None, None, 0, None, 0, None)
    assert rc == _lib.MP_ERR_UNSUPPORTED
    with pytest.raises(NotImplementedError):
        _lib.check(rc, "mp_box_nms_f32")
    rc = lib.mp_nearest_f32(None, None, 4, None, None, 4, 1, 100, 0, 0, None, None, None, None, None, None, None, 0, None)
    assert rc == _lib.MP_ERR_UNSUPPORTED and "D % 64 == 0" in _lib.last_error()
    rc = lib.mp_warp_f32(None, 1, 1, 8, 8, None, None, None, 7, 0, None, None)
    assert rc == _lib.MP_ERR_INVALID
    with pytest.raises(ValueError):
        _lib.check(rc, "mp_warp_f32")


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from multipoint_b200 import ops, utils
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.detector_head(torch.zeros(1, 65, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        utils.box_nms(torch.zeros(16, 16), 4, 0.015)
    with pytest.raises(RuntimeError, match="CUDA"):
        utils.get_matches(np.zeros((4, 64), np.float32), np.zeros((4, 64), np.float32), 'bfmatcher', crossCheck=True)
    # reference error behaviour that does not need a device
    with pytest.raises(ValueError, match="either 2D"):
        utils.box_nms(torch.zeros(3, 16, 16), 4, 0.015)
    with pytest.raises(ValueError, match="unknown matching method"):
        utils.get_matches(np.zeros((4, 64), np.float32), np.zeros((4, 64), np.float32), 'nope')
    with pytest.raises(ValueError, match="non-negative"):
        utils.ThresholdMatcher(threshold=-0.1)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "multipoint_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "mp_oracle" not in src and "import oracle" not in src and "ref_shim" not in src, f


def test_sample_homography_and_valid_mask_match_reference():
    from multipoint_b200 import utils
    g = load_golden("homographies")
    cfg_default = dict(utils.homography_adaptation_default_config['homographies'])
    cfg_export = dict(translation=True, rotation=True, scaling=True, perspective=True, scaling_amplitude=0.2,
                      perspective_amplitude_x=0.2, perspective_amplitude_y=0.2, patch_ratio=0.85, max_angle=1.57,
                      allow_artifacts=True)
    cfg_noart = dict(cfg_export, allow_artifacts=False, translation_overflow=0.05)
    for tag, cfg in [("default", cfg_default), ("export", cfg_export), ("noart", cfg_noart), ("small", cfg_export)]:
        shape = tuple(int(v) for v in g[tag + "_shape"])
        seed, erosion = (int(v) for v in g[tag + "_seed"])
        np.random.seed(seed)
        want_masks = np.unpackbits(g[tag + "_mask"], axis=-1)[..., :shape[1]]
        for i in range(len(g[tag + "_H"])):
            Hm = utils.sample_homography(np.array(shape), **cfg)
            np.testing.assert_array_equal(Hm, g[tag + "_H"][i])           # same RNG stream, same solver: bit-exact
            mask = utils.compute_valid_mask(shape, Hm, erosion, True)
            np.testing.assert_array_equal(mask != 0, want_masks[i] != 0)
    Hm = g["small_H"][0]
    np.testing.assert_array_equal(utils.compute_valid_mask((64, 80), Hm, 0, False) != 0,
                                  np.unpackbits(g["small_mask_e0"], axis=-1)[..., :80] != 0)
    np.testing.assert_array_equal(utils.compute_valid_mask((64, 80), Hm, 2, False) != 0,
                                  np.unpackbits(g["small_mask_e2nb"], axis=-1)[..., :80] != 0)
    np.testing.assert_array_equal(utils.warp_keypoints(g["wk_kp"], Hm), g["wk_out"])
    np.testing.assert_array_equal(utils.filter_points(g["wk_out"], (64, 80)), g["wk_filtered"])
    # pre-sampling num-1 homographies consumes the stream exactly like the reference's loop
    cfg = utils._check_ha_config(dict(num=9, erosion_radius=3, homographies=cfg_export))
    np.random.seed(3)
    Hs, masks = utils.sample_adaptation_homographies((64, 80), cfg)
    np.testing.assert_array_equal(Hs, g["small_H"])
    np.testing.assert_array_equal(masks != 0, np.unpackbits(g["small_mask"], axis=-1)[..., :80] != 0)


def test_invert_homographies_is_cv_invert():
    """The batched closed-form inverse fed to mp_valid_mask_u8 is bit-identical to cv2.invert (what
    cv2.warpPerspective applies internally), including the singular case."""
    import cv2
    from multipoint_b200 import utils
    g = load_golden("homographies")
    Hs = np.concatenate([g["default_H"], g["export_H"], g["small_H"], np.zeros((1, 3, 3)), np.eye(3)[None]])
    got = utils.invert_homographies(Hs)
    for Hm, inv in zip(Hs, got):
        np.testing.assert_array_equal(inv, cv2.invert(Hm)[1])
    Hs2, none = utils.sample_adaptation_homographies((64, 80), utils._check_ha_config(dict(num=4)), with_masks=False)
    assert none is None and Hs2.shape == (3, 3, 3)


def test_normalized_warp_matrix_matches_restatement():
    from multipoint_b200 import utils
    g = load_golden("adaptation")
    M = torch.from_numpy(g["H"].astype(np.float32))
    np.testing.assert_allclose(utils.normalized_warp_matrix(M, (64, 80), (64, 80)).numpy(), g["A_warp"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(utils.normalized_warp_matrix(torch.inverse(M), (64, 80), (64, 80)).numpy(), g["A_unwarp"], rtol=0, atol=1e-6)


def test_config_handling_matches_reference():
    from multipoint_b200 import utils
    d = utils.dict_update({'a': {'b': 1, 'c': 2}, 'd': 3}, {'a': {'b': 5}, 'e': 6})
    assert d == {'a': {'b': 5, 'c': 2}, 'd': 3, 'e': 6}
    with pytest.raises(ValueError, match="num must be larger than 0"):
        utils._check_ha_config({'num': 0})
    with pytest.raises(ValueError, match="filter_size must be uneven"):
        utils._check_ha_config({'filter_size': 2})
    utils._check_ha_config({'num': 2})
    assert utils.homography_adaptation_default_config['num'] == 100          # defaults are not mutated
    w = utils.fix_model_weigth_keys({'module__encoder.0.weight': 1, 'x__y__head.bias': 2, 'plain': 3})
    assert list(w.keys()) == ['encoder.0.weight', 'head.bias', 'plain']
    f = utils.get_gaussian_filter(5)
    assert f.weight.shape == (1, 1, 5, 5) and abs(float(f.weight.sum()) - 1.0) < 1e-6
    x = torch.arange(2 * 3 * 4 * 6, dtype=torch.float32).reshape(2, 3, 4, 6)
    s2d = utils.space_to_depth(x, 2)
    assert s2d.shape == (2, 12, 2, 3)


def test_model_layout_matches_reference_state_dict():
    from multipoint_b200.models import MultiPoint
    g = load_golden("model")
    for tag, cfg in [("shipped", {'multispectral': False, 'descriptor_size': 64}), ("multi", {'multispectral': True, 'descriptor_size': 256})]:
        torch.manual_seed(0)
        net = MultiPoint(cfg)
        assert list(net.state_dict().keys()) == list(g[tag + "_keys"])
        assert [str(tuple(v.shape)) for v in net.state_dict().values()] == list(g[tag + "_shapes"])
        assert sum(p.numel() for p in net.parameters()) == int(g[tag + "_nparams"])
    net = MultiPoint({'multispectral': False, 'descriptor_size': 64})
    net.train()
    out = net({'image': torch.from_numpy(syn.images(1, 1, 32, 32))})          # training mode: logits only, pure torch
    assert out['prob'] is None and out['logits'].shape == (1, 65, 4, 4) and out['desc'].shape == (1, 64, 4, 4)
    with pytest.raises(ValueError):
        net.set_force_return_logits("yes")


def test_synthetic_inputs_are_reproducible():
    a = syn.heatmap(31, 1, 64, 80)
    assert syn.checksum(a) == syn.checksum(syn.heatmap(31, 1, 64, 80))
    frac = float((syn.heatmap(5, 1, 512, 640) > 0.015).mean())
    assert 0.1 < frac < 0.25
    d1, d2 = syn.descriptor_sets(1, 50, 70, 64, 0.05, 3)
    assert d1.shape == (50, 64) and d2.shape == (70, 64)
    np.testing.assert_allclose(np.linalg.norm(d2, axis=1), 1.0, atol=1e-6)
    batch = syn.image_pair_batch(2, 3, 32, 40)
    assert batch['optical']['image'].shape == (3, 1, 32, 40) and batch['thermal']['is_optical'].sum() == 0


def test_fused_glue_walker_on_cpu_with_oracle_kernels(monkeypatch):
    """models.MultiPoint._run replaces ReLU / BatchNorm / MaxPool / pad (and the first layer) by fused kernels in
    inference.  The CUDA kernels are checked on the GPU; here the *walker* -- which modules it fuses, which pad mode
    and pooling it passes on, that nothing is applied twice or skipped -- runs on the CPU for every block layout,
    with the two kernels replaced by the oracle's restatements, against the plain module-by-module forward."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import oracle
    from multipoint_b200 import models, ops

    def unfold(scale, shift):   # the walker passes the folded affine; the oracle takes BatchNorm parameters
        C = scale.numel()
        return dict(mean=np.zeros(C, np.float32), var=np.ones(C, np.float32), eps=0.0, weight=scale.numpy(), bias=shift.numpy())

    def fake_glue(x, scale, shift, bn_first=False, pool=False, pad=1, reflect=True, conv_bias=None):
        return torch.from_numpy(oracle.relu_bn_pad(x.numpy(), conv_bias=None if conv_bias is None else conv_bias.numpy(), bn_first=bn_first,
                                                   pool=pool, pad=pad, reflect=reflect, **unfold(scale, shift)))

    def fake_conv1(image, weight, conv_bias, scale, shift, bn_first=False, in_reflect=True, pad=1, out_reflect=True):
        return torch.from_numpy(oracle.conv1_relu_bn_pad(image.numpy(), weight.numpy(), None if conv_bias is None else conv_bias.numpy(),
                                                         bn_first=bn_first, in_reflect=in_reflect, pad=pad, out_reflect=out_reflect,
                                                         **unfold(scale, shift)))

    monkeypatch.setattr(ops, "relu_bn_pad", fake_glue)
    monkeypatch.setattr(ops, "conv1_relu_bn_pad", fake_conv1)
    monkeypatch.setattr(models.MultiPoint, "_glue_on_any_device", True)
    for cfg in ({'multispectral': True, 'descriptor_size': 32},
                {'multispectral': False, 'descriptor_size': 16, 'bn_first': True},
                {'multispectral': False, 'descriptor_size': 16, 'reflection_pad': False},
                {'multispectral': False, 'descriptor_size': 16, 'double_convolution': False},
                {'multispectral': False, 'descriptor_size': 16, 'final_batchnorm': False, 'channel_version': 1}):
        torch.manual_seed(2)
        net = models.MultiPoint(dict(cfg)).eval()
        with torch.no_grad():
            for m in net.modules():          # non-trivial BatchNorm statistics
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.running_mean.normal_(std=0.2); m.running_var.uniform_(0.5, 2.0); m.weight.uniform_(0.5, 1.5); m.bias.normal_(std=0.2)
            img = torch.rand(2, 1, 32, 40)
            data = {'image': img, 'is_optical': torch.tensor([[1], [0]], dtype=torch.bool)}
            x_fused = net.encode(data)
            fl, fr = net.backbone_outputs(data)
        with torch.enable_grad():            # autograd on -> the module path
            x_mod = net.encode(data).detach()
            ml, mr = net.backbone_outputs(data)
        np.testing.assert_allclose(x_fused.numpy(), x_mod.numpy(), rtol=1e-4, atol=1e-5, err_msg=str(cfg))
        np.testing.assert_allclose(fl.numpy(), ml.detach().numpy(), rtol=1e-3, atol=1e-4, err_msg=str(cfg))
        np.testing.assert_allclose(fr.numpy(), mr.detach().numpy(), rtol=1e-3, atol=1e-4, err_msg=str(cfg))


def test_folded_batchnorm_cache_follows_training_mode_updates():
    """a training-mode forward rewrites running_mean / running_var without bumping the tensors' version
    counters, so an affine folded during an earlier eval forward must not survive it."""
    import torch
    from multipoint_b200 import models
    from multipoint_b200.pipeline import calibrate_random_init

    torch.manual_seed(0)
    net = models.MultiPoint({'multispectral': False, 'descriptor_size': 16}).eval()
    bns = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    before = [tuple(t.clone() for t in models.MultiPoint._folded_bn(m)) for m in bns]   # what an eval forward caches
    assert all('_mp_folded' in m.__dict__ for m in bns)
    calibrate_random_init(net, torch.rand(2, 1, 32, 40))       # train-mode forward with momentum 1
    assert not net.training
    changed = 0
    for m, (s0, b0) in zip(bns, before):
        scale, shift = models.MultiPoint._folded_bn(m)
        want = m.weight.detach() / torch.sqrt(m.running_var + m.eps)
        torch.testing.assert_close(scale, want, rtol=1e-6, atol=0)
        torch.testing.assert_close(shift, m.bias.detach() - m.running_mean * want, rtol=1e-5, atol=1e-7)
        changed += int(not torch.equal(scale, s0))
    assert changed == len(bns)
    # a BatchNorm put into training mode on its own (not through the parent) is covered as well
    bn = bns[0]
    models.MultiPoint._folded_bn(bn)
    bn.train()
    bn(torch.randn(2, bn.num_features, 4, 4) * 3 + 1)
    bn.eval()
    scale, _ = models.MultiPoint._folded_bn(bn)
    torch.testing.assert_close(scale, bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps), rtol=1e-6, atol=0)


def test_gaussian_filter_matches_reference_fixture():
    """Row 12 (utils.get_gaussian_filter, multipoint/utils/utils.py:124-157): weights for three sizes (default and
    explicit sigma) and the filter's action with the adaptation's reflection padding, against the reference's own
    outputs frozen by oracle/gen_golden.py."""
    import torch
    from multipoint_b200 import utils
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "adaptation.npz"))
    for k, sg in ((3, None), (5, None), (7, 1.5)):
        f = utils.get_gaussian_filter(k) if sg is None else utils.get_gaussian_filter(k, sg)
        assert isinstance(f, torch.nn.Conv2d) and f.bias is None and not f.weight.requires_grad
        np.testing.assert_array_equal(f.weight.detach().numpy(), g["gauss_w_%d" % k])
    out = utils.get_gaussian_filter(5)(torch.nn.ReflectionPad2d(2)(torch.from_numpy(g["gauss_in"]))).detach().numpy()
    np.testing.assert_allclose(out, g["gauss_out_5"], rtol=1e-6, atol=1e-9)
