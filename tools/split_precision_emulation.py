#!/usr/bin/env python
"""CPU emulation: how many bf16 planes / tensor-core passes does a 3x3 convolution need to stay at fp32
accuracy?    float64 convolution = truth; errors relative to the output rms."""
import torch

torch.manual_seed(0)
conv = torch.nn.Conv2d(64, 64, 3, bias=False)
x = torch.randn(2, 64, 66, 82)
w = conv.weight.detach()


def planes(t):
    hi = t.bfloat16().float()
    mid = (t - hi).bfloat16().float()
    lo = (t - hi - mid).bfloat16().float()
    return hi, mid, lo


f = torch.nn.functional.conv2d
xh, xm, xl = planes(x)
wh, wm, wl = planes(w)
truth = f(x.double(), w.double())
rms = float(truth.pow(2).mean().sqrt())
for name, y in [("fp32 convolution", f(x, w)),
                ("hi+mid, 3 terms (hh, hm, mh)", f(xm, wh) + f(xh, wm) + f(xh, wh)),
                ("hi+mid, 4 terms (+ mm)", f(xm, wm) + f(xm, wh) + f(xh, wm) + f(xh, wh)),
                ("hi+mid+lo, 6 terms (+ mm, hl, lh)", f(xl, wh) + f(xh, wl) + f(xm, wm) + f(xm, wh) + f(xh, wm) + f(xh, wh))]:
    e = (y.double() - truth).abs()
    print("%-36s max %.3e   rms %.3e   rms / output rms %.2e" % (name, float(e.max()), float(e.pow(2).mean().sqrt()),
                                                               float(e.pow(2).mean().sqrt()) / rms))
