#!/usr/bin/env python
"""Experiment: cuDNN fp32 (no TF32) 3x3 convolutions of the MultiPoint encoder, NCHW against channels_last."""
import torch, time
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
B = 64
layers = [(64, 64, 512, 640), (64, 64, 256, 320), (64, 128, 128, 160), (128, 128, 128, 160), (128, 128, 64, 80), (128, 256, 64, 80)]
for cin, cout, H, W in layers:
    conv = torch.nn.Conv2d(cin, cout, 3).to(dev)
    x = torch.randn(B, cin, H + 2, W + 2, device=dev)
    res = {}
    for name, fmt in (("nchw", torch.contiguous_format), ("nhwc", torch.channels_last)):
        c = conv.to(memory_format=fmt)
        xi = x.to(memory_format=fmt)
        with torch.no_grad():
            for _ in range(3):
                y = c(xi)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                y = c(xi)
            e1.record()
            torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 5
    fl = 2.0 * B * cout * cin * 9 * H * W
    print("%3d -> %3d @ %dx%d: nchw %.3f ms (%.1f TF/s)  nhwc %.3f ms (%.1f TF/s)" % (cin, cout, H, W, res["nchw"], fl / res["nchw"] / 1e9, res["nhwc"], fl / res["nhwc"] / 1e9), flush=True)
