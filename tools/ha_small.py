import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from multipoint_b200 import ops
n, B, H, W = 3, 2, 64, 80
g = torch.Generator(device="cuda").manual_seed(0)
pa = torch.rand((n, B, H, W), generator=g, device="cuda"); pb = torch.rand((n, B, H, W), generator=g, device="cuda")
p0 = torch.rand((B, H, W), generator=g, device="cuda")
masks = torch.ones((n, H, W), dtype=torch.uint8, device="cuda")
A = torch.eye(3, device="cuda")[None].repeat(n, 1, 1).contiguous()
out = ops.ha_aggregate(p0, pa, pb, masks, A, 'prod', 2)
torch.cuda.synchronize()
print("ok", float(out.sum()))
