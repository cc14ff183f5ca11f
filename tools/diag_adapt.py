import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from multipoint_b200 import utils, ops
g = dict(np.load("/root/repo/tests/golden/adaptation.npz"))
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
conv = torch.nn.Conv2d(1, 65, 8, stride=8).cuda(); conv.weight.data = cu(g["stub_w"]); conv.bias.data = cu(g["stub_b"])
def net(data):
    with torch.no_grad(): return {'prob': ops.detector_head(conv(data['image']).float())}
torch.backends.cudnn.allow_tf32 = False
masks = (g["masks"] != 0).astype(np.uint8)
H, W = 64, 80
def ties(A, delta):
    xs = np.linspace(-1, 1, W); ys = np.linspace(-1, 1, H)
    X, Y = np.meshgrid(xs, ys)
    t = np.zeros((H, W), bool)
    for M in A.astype(np.float64):
        z = M[2,0]*X + M[2,1]*Y + M[2,2]
        ix = ((M[0,0]*X + M[0,1]*Y + M[0,2]) / z + 1) / 2 * (W - 1)
        iy = ((M[1,0]*X + M[1,1]*Y + M[1,2]) / z + 1) / 2 * (H - 1)
        t |= (np.abs(ix - np.floor(ix) - 0.5) < delta) | (np.abs(iy - np.floor(iy) - 0.5) < delta)
    return t
for fs, key in ((0, "single"), (5, "single_f5")):
    cfg = dict(num=6, min_count=2, erosion_radius=3, filter_size=fs)
    for nm in (None, (g["A_warp"], g["A_unwarp"])):
        out = utils.homographic_adaptation({'image': cu(g["img_o"])}, net, dict(cfg), homographies=g["H"], masks=masks, normalized_matrices=nm).cpu().numpy()
        want = g[key]
        for rtol in (1e-5, 1e-4):
            bad = np.abs(out - want) > 1e-7 + rtol * np.abs(want)
            t = ties(g["A_unwarp"], 1e-3)
            print(key, "fixture matrices" if nm else "own matrices", "rtol", rtol, "bad", int(bad.sum()), "bad outside ties", int((bad & ~t[None, None]).sum()), "tie px", int(t.sum()),
                  "max rel outside ties", float((np.abs(out - want) / np.maximum(np.abs(want), 1e-6))[~np.broadcast_to(t[None,None], out.shape)].max()))
