"""Host-side mirror of ``multipoint.models.MultiPoint`` (reference: multipoint/models/MultiPoint.py).

Same constructor config, same module / state-dict names (``encoder[_optical|_thermal].N.*``,
``detector_head_convolutions.N.*``, ``descriptor_head_convolutions.N.*``), same parameter creation
order (so a given ``torch.manual_seed`` gives the reference's initial weights) and the same
``forward(data) -> {'prob', 'logits', 'desc'}`` contract (MultiPoint.py:99-135).

The convolutional backbone and the two head convolutions stay in PyTorch / cuDNN (north star).
What follows them is the hot path and runs in this repo's CUDA kernels through the C ABI:
  detector head tail   softmax(65) + dustbin drop + PixelShuffle(8)   MultiPoint.py:150-158
  descriptor head tail F.normalize(dim=1)                             MultiPoint.py:160-166
There is no eager fallback for those tails: inference needs a CUDA device.
"""
import copy

import torch
import torch.nn as nn

from . import ops
from .utils import dict_update

_CHANNELS = {0: ([1, 64, 64, 128, 128], None), 1: ([1, 32, 64, 96, 128], 'desc'), 2: ([1, 8, 16, 32, 64], 'desc')}


def _drop_folded_when_training(bn, _inputs):
    if bn.training:
        bn.__dict__.pop('_mp_folded', None)


class MultiPoint(nn.Module):
    # tests set this to walk the fused path on the CPU with ops.relu_bn_pad / ops.conv1_relu_bn_pad replaced by the oracle
    _glue_on_any_device = False
    default_config = {
        'multispectral': True,
        'descriptor_head': True,
        'intepolation_mode': 'bilinear',
        'descriptor_size': 256,
        'normalize_descriptors': True,
        'final_batchnorm': True,
        'reflection_pad': True,
        'bn_first': False,
        'double_convolution': True,
        'channel_version': 0,
        'verbose': False,
        'mixed_precision': False,
        'force_return_logits': False,
    }

    def __init__(self, config=None):
        super().__init__()
        # MultiPoint.py:28-31: a given config is merged into a copy of the defaults
        self.config = dict_update(copy.deepcopy(self.default_config), config) if config else self.default_config
        cfg = self.config
        self.pad_method = nn.ReflectionPad2d if cfg['reflection_pad'] else nn.ZeroPad2d

        version = cfg['channel_version']
        if version not in _CHANNELS:
            print('Unknown channel_version: ', version)
            version = 0
        self.n_channels, head = _CHANNELS[version]
        self.head_channels = cfg['descriptor_size'] if head == 'desc' else 256

        # creation order matters for seeded initialisation: thermal before optical (MultiPoint.py:54-58)
        if cfg['multispectral']:
            self.encoder_thermal = self.generate_encoder()
            self.encoder_optical = self.generate_encoder()
        else:
            self.encoder = self.generate_encoder()

        self.detector_head_convolutions = self._head(65)
        # kept for state-dict / attribute compatibility; the tails run in CUDA kernels
        self.softmax = nn.Softmax2d()
        self.shuffle = nn.PixelShuffle(8)
        if cfg['descriptor_head']:
            self.descriptor_head_convolutions = self._head(cfg['descriptor_size'])

        if cfg['verbose']:
            print('MultiPoint number of trainable parameter: ' + str(sum(p.numel() for p in self.parameters())))

    # ------------------------------------------------------------------ construction helpers
    @staticmethod
    def _batchnorm(N):
        """BatchNorm2d whose cached eval-mode affine (``_folded_bn``) is dropped by every training-mode forward: that
        forward rewrites running_mean / running_var without bumping the tensors' version counters."""
        bn = nn.BatchNorm2d(N)
        bn.register_forward_pre_hook(_drop_folded_when_training)
        return bn

    def getNonlinearity(self, N):
        if self.config['bn_first']:
            return self._batchnorm(N), nn.ReLU(True)
        return nn.ReLU(True), self._batchnorm(N)

    def getConvolutionBlock(self, N_in, N_out):
        block = [self.pad_method(1), nn.Conv2d(N_in, N_out, 3), *self.getNonlinearity(N_out)]
        if self.config['double_convolution']:
            block += [self.pad_method(1), nn.Conv2d(N_out, N_out, 3), *self.getNonlinearity(N_out)]
        return tuple(block)

    def generate_encoder(self):
        c = self.n_channels
        layers = []
        for stage in range(4):
            layers += list(self.getConvolutionBlock(c[stage], c[stage + 1]))
            if stage < 3:
                layers.append(nn.MaxPool2d(2, 2))
        return nn.Sequential(*layers)

    def _head(self, out_channels):
        layers = [self.pad_method(1), nn.Conv2d(self.n_channels[4], self.head_channels, 3),
                  *self.getNonlinearity(self.head_channels), nn.Conv2d(self.head_channels, out_channels, 1)]
        if self.config['final_batchnorm']:
            layers.append(self._batchnorm(out_channels))
        return nn.Sequential(*layers)

    # ------------------------------------------------------------------ fused inference path
    @staticmethod
    def _folded_bn(bn):
        """Eval-mode BatchNorm as a per-channel affine (scale, shift); cached on the module until a parameter or
        buffer changes: in-place updates such as load_state_dict bump the tensors' version counters, and a
        training-mode forward (which updates the running statistics without bumping them) drops the cache through
        the module's forward-pre hook (``_batchnorm``)."""
        ver = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
               bn.weight.data_ptr(), bn.running_var.data_ptr())
        hit = getattr(bn, '_mp_folded', None)
        if hit is None or hit[0] != ver:
            with torch.no_grad():
                scale = (bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)).contiguous()
                shift = (bn.bias.float() - bn.running_mean.float() * scale).contiguous()
            hit = (ver, scale, shift)
            bn._mp_folded = hit
        return hit[1], hit[2]

    def _run(self, seq, x):
        """``seq(x)`` for the encoder / head Sequentials.  In inference (eval mode, no autograd, fp32 CUDA input) the
        ReLU -> BatchNorm [-> MaxPool] [-> pad] chain after every 3x3 convolution runs as one pass
        (ops.relu_bn_pad) instead of three or four full-tensor elementwise kernels; the convolutions are the same
        cuDNN calls.  Anything that does not match that pattern, and every other mode, goes through the modules."""
        if (self.training or torch.is_grad_enabled() or not (x.is_cuda or self._glue_on_any_device) or x.dtype != torch.float32
                or self.config['mixed_precision'] or torch.is_autocast_enabled()):
            return seq(x)
        pads = (nn.ReflectionPad2d, nn.ZeroPad2d)
        mods = list(seq)
        i, padded = 0, False
        # first encoder layer: pad -> Conv2d(1 -> C, 3x3) -> ReLU/BatchNorm [-> pad] on a one-channel image is one kernel
        if (len(mods) >= 4 and isinstance(mods[0], pads) and tuple(mods[0].padding) == (1, 1, 1, 1) and x.shape[1] == 1
                and isinstance(mods[1], nn.Conv2d) and mods[1].in_channels == 1 and mods[1].kernel_size == (3, 3)
                and mods[1].stride == (1, 1) and mods[1].padding == (0, 0) and mods[1].dilation == (1, 1) and mods[1].groups == 1
                and {type(mods[2]), type(mods[3])} == {nn.ReLU, nn.BatchNorm2d} and not (mods[2].training or mods[3].training)
                and min(x.shape[-2:]) >= 2 and x.shape[-1] + 2 <= 768
                and not (len(mods) > 4 and isinstance(mods[4], nn.MaxPool2d))):
            bn_first = isinstance(mods[2], nn.BatchNorm2d)
            scale, shift = self._folded_bn(mods[2] if bn_first else mods[3])
            pad_next = len(mods) > 4 and isinstance(mods[4], pads) and tuple(mods[4].padding) == (1, 1, 1, 1)
            conv = mods[1]
            x = ops.conv1_relu_bn_pad(x.contiguous(), conv.weight.detach(), None if conv.bias is None else conv.bias.detach(), scale, shift,
                                      bn_first=bn_first, in_reflect=isinstance(mods[0], nn.ReflectionPad2d), pad=1 if pad_next else 0,
                                      out_reflect=isinstance(mods[4], nn.ReflectionPad2d) if pad_next else True)
            padded = pad_next
            i = 4
        while i < len(mods):
            m = mods[i]
            if isinstance(m, pads) and padded:      # the fused pass before already produced the padded tensor
                padded = False
                i += 1
                continue
            pair = mods[i + 1:i + 3]
            fusable = (isinstance(m, nn.Conv2d) and m.padding_mode == 'zeros' and len(pair) == 2
                       and {type(pair[0]), type(pair[1])} == {nn.ReLU, nn.BatchNorm2d}
                       and not (pair[0].training or pair[1].training))
            if not fusable:
                x = m(x)
                i += 1
                continue
            # the convolution without its bias (the fused pass adds it: torch would spend one more elementwise kernel)
            x = nn.functional.conv2d(x, m.weight, None, m.stride, m.padding, m.dilation, m.groups)
            bn_first = isinstance(pair[0], nn.BatchNorm2d)
            bn = pair[0] if bn_first else pair[1]
            j = i + 3
            pool = (j < len(mods) and isinstance(mods[j], nn.MaxPool2d) and mods[j].kernel_size in (2, (2, 2))
                    and mods[j].stride in (2, (2, 2)) and mods[j].padding in (0, (0, 0)) and not mods[j].ceil_mode
                    and x.shape[-2] % 2 == 0 and x.shape[-1] % 2 == 0)
            if pool:
                j += 1
            pad_next = (j < len(mods) and isinstance(mods[j], pads) and tuple(mods[j].padding) == (1, 1, 1, 1)
                        and min(x.shape[-2:]) // (2 if pool else 1) >= 2)
            scale, shift = self._folded_bn(bn)
            x = ops.relu_bn_pad(x.contiguous(), scale, shift, bn_first=bn_first, pool=pool, pad=1 if pad_next else 0,
                                reflect=isinstance(mods[j], nn.ReflectionPad2d) if pad_next else True,
                                conv_bias=None if m.bias is None else m.bias.detach())
            padded = pad_next
            i = j
        return x

    # ------------------------------------------------------------------ reference API
    def set_force_return_logits(self, value):
        if not isinstance(value, bool):
            raise ValueError('set_force_return_logits: The input value needs to be a bool')
        self.config['force_return_logits'] = value

    def forward(self, data):
        if self.config['mixed_precision']:
            with torch.autocast('cuda'):
                return self.forward_impl(data)
        return self.forward_impl(data)

    def encode(self, data):
        """Encoder routing of MultiPoint.py:107-124: each row goes through the optical or the thermal
        encoder according to data['is_optical'][:,0]."""
        image = data['image']
        if not self.config['multispectral']:
            return self._run(self.encoder, image)
        n_hint = data.get('n_optical')
        if n_hint is not None:
            # the caller vouches that rows [0, n) are optical and the rest thermal (KeypointPipeline builds its batch
            # that way): no device->host read of is_optical, no gather / scatter of the rows
            n = int(n_hint)
            if n == image.shape[0]:
                return self._run(self.encoder_optical, image)
            if n == 0:
                return self._run(self.encoder_thermal, image)
            return torch.cat([self._run(self.encoder_optical, image[:n]), self._run(self.encoder_thermal, image[n:])])
        sel = data['is_optical'][:, 0].bool()
        n_opt = int(sel.sum())
        if n_opt == image.shape[0]:
            return self._run(self.encoder_optical, image)
        if n_opt == 0:
            return self._run(self.encoder_thermal, image)
        xo = self._run(self.encoder_optical, image[sel])
        xt = self._run(self.encoder_thermal, image[~sel])
        x = torch.empty((image.shape[0],) + tuple(xo.shape[1:]), dtype=xo.dtype, device=xo.device)
        x[sel] = xo
        x[~sel] = xt
        return x

    def forward_impl(self, data):
        x = self.encode(data)
        prob, logits = self.detector_head(x)
        out = {'prob': prob, 'logits': logits}
        if self.config['descriptor_head']:
            out['desc'] = self.descriptor_head(x)
        return out

    def detector_head(self, x):
        logits = self._run(self.detector_head_convolutions, x).to(torch.float)
        if self.training or self.config['force_return_logits']:
            return None, logits
        return ops.detector_head(logits), None

    def descriptor_head(self, x, channels_last=False):
        x = self._run(self.descriptor_head_convolutions, x).to(torch.float)
        if not self.config['normalize_descriptors']:
            return x
        if torch.is_grad_enabled() and x.requires_grad:
            # training is outside the hot path: keep autograd alive
            return torch.nn.functional.normalize(x, p=2, dim=1)
        nchw, nhwc = ops.normalize_descriptors(x, nchw=not channels_last, nhwc=channels_last)
        return nhwc if channels_last else nchw

    # ------------------------------------------------------------------ fused entry used by the pipeline
    def backbone_outputs(self, data):
        """(logits (B,65,Hc,Wc), raw descriptor map (B,D,Hc,Wc)) -- everything cuDNN computes."""
        x = self.encode(data)
        logits = self._run(self.detector_head_convolutions, x).to(torch.float)
        raw = self._run(self.descriptor_head_convolutions, x).to(torch.float) if self.config['descriptor_head'] else None
        return logits, raw


class SuperPointMagicLeap(nn.Module):
    """Mirror of multipoint/models/SuperPointMagicLeap.py:5-66: the pretrained
    MagicLeap SuperPoint layout (same attribute names, so its state dict loads) with the heatmap
    built on the device by ops.heatmap_magicleap instead of the reference's per-sample
    tensor -> numpy -> tensor loop (generate_heatmap, :68-85)."""

    def __init__(self, config=None):
        super().__init__()
        self.relu = nn.ReLU(inplace=True)
        self.pool = nn.MaxPool2d(kernel_size=2, stride=2)
        # creation order = the reference's, so a seeded random init draws the same weights
        layers = [('1a', 1, 64, 3), ('1b', 64, 64, 3), ('2a', 64, 64, 3), ('2b', 64, 64, 3), ('3a', 64, 128, 3),
                  ('3b', 128, 128, 3), ('4a', 128, 128, 3), ('4b', 128, 128, 3), ('Pa', 128, 256, 3), ('Pb', 256, 65, 1),
                  ('Da', 128, 256, 3), ('Db', 256, 256, 1)]
        for tag, cin, cout, k in layers:
            setattr(self, 'conv' + tag, nn.Conv2d(cin, cout, kernel_size=k, stride=1, padding=k // 2))

    def forward(self, data):
        x = data['image']
        for stage in ('1', '2', '3', '4'):
            x = self.relu(getattr(self, 'conv%sa' % stage)(x))
            x = self.relu(getattr(self, 'conv%sb' % stage)(x))
            if stage != '4':
                x = self.pool(x)
        semi = self.convPb(self.relu(self.convPa(x)))
        desc = self.convDb(self.relu(self.convDa(x)))
        desc = desc.div(torch.unsqueeze(torch.norm(desc, p=2, dim=1), 1))  # no eps clamp, like :59-60
        return {'logits': semi, 'desc': desc, 'prob': self.generate_heatmap(semi, data['image'].shape)}

    def generate_heatmap(self, semi, shape):
        prob = ops.heatmap_magicleap(semi)
        if tuple(prob.shape) != tuple(shape):
            raise ValueError("image shape %s does not match 8x the logits grid %s" % (tuple(shape), tuple(prob.shape)))
        return prob
