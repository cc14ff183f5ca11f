#!/usr/bin/env python
"""How fast can cuDNN run the (unchanged) MultiPoint backbone in fp32?  Tuning aid, not a bench line."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from multipoint_b200 import synthetic as syn

def timed(fn, it=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it

dev = torch.device("cuda")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
net = bench.build_net(256, dev)
b = syn.image_pair_batch(1000, 64, 512, 640)
img = torch.cat([torch.from_numpy(b['optical']['image']), torch.from_numpy(b['thermal']['image'])]).to(dev)
opt = torch.cat([torch.from_numpy(b['optical']['is_optical']), torch.from_numpy(b['thermal']['is_optical'])]).to(dev)
res = {}
with torch.no_grad():
    res["nchw_fp32_b128"] = timed(lambda: net.backbone_outputs({'image': img, 'is_optical': opt}))
    for chunk in (16, 32, 64):
        def run():
            for s in range(0, 128, chunk):
                net.backbone_outputs({'image': img[s:s + chunk], 'is_optical': opt[s:s + chunk]})
        res["nchw_fp32_chunks_of_%d" % chunk] = timed(run)
    net_cl = net.to(memory_format=torch.channels_last)
    img_cl = img.contiguous(memory_format=torch.channels_last)
    res["channels_last_fp32_b128"] = timed(lambda: net_cl.backbone_outputs({'image': img_cl, 'is_optical': opt}))
    torch.backends.cudnn.allow_tf32 = True
    res["channels_last_tf32_b128"] = timed(lambda: net_cl.backbone_outputs({'image': img_cl, 'is_optical': opt}))
    net = net.to(memory_format=torch.contiguous_format)
    res["nchw_tf32_b128"] = timed(lambda: net.backbone_outputs({'image': img, 'is_optical': opt}))
    with torch.autocast('cuda', dtype=torch.bfloat16):
        res["nchw_bf16_autocast_b128"] = timed(lambda: net.backbone_outputs({'image': img, 'is_optical': opt}))
print(json.dumps(res))
