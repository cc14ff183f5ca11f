#!/usr/bin/env python
"""Kernel-level breakdown of one backbone forward (128 images of 512x640, fp32) with torch.profiler.  Tuning aid."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from multipoint_b200 import synthetic as syn
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
net = bench.build_net(256, dev)
b = syn.image_pair_batch(1000, 64, 512, 640)
img = torch.cat([torch.from_numpy(b['optical']['image']), torch.from_numpy(b['thermal']['image'])]).to(dev)
opt = torch.cat([torch.from_numpy(b['optical']['is_optical']), torch.from_numpy(b['thermal']['is_optical'])]).to(dev)
data = {'image': img, 'is_optical': opt}
with torch.no_grad():
    for _ in range(3):
        net.backbone_outputs(data)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        net.backbone_outputs(data)
        torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
    if t:
        rows.append((t / 1000.0, e.count, e.key[:110]))
rows.sort(reverse=True)
total = sum(r[0] for r in rows)
print("total device ms %.2f" % total)
for t, n, k in rows[:25]:
    print("%8.3f ms  x%-3d %s" % (t, n, k))
