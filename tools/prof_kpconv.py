"""One KPConv forward + backward on a subsampled synthetic tile (the subject of the ncu captures under profiles/).

    python tools/prof_kpconv.py [C] [reps]
"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from weasal_b200 import ops
from weasal_b200.synthetic import make_als_tile
C_ = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
pts, _, _ = make_als_tile(1, 50.0, 40.0)
P = torch.from_numpy(pts).to(dev)
L = np.array([len(pts)], np.int32)
sp, sl = ops.grid_subsample(P, L, sampleDl=0.4, order="first")
S = sp.contiguous(); Ls = np.array([len(S)], np.int32)
nb = ops.batch_query(S, S, Ls, Ls, 1.0, dtype=torch.int32, cap_hint=64)
n = len(S)
torch.manual_seed(0)
x = torch.randn(n, C_, device=dev, requires_grad=True)
w = (torch.randn(15, C_, C_, device=dev) / C_ ** 0.5).requires_grad_(True)
# kernel points as the reference lays them out: one at the centre, 14 on a shell of 0.66 x the conv radius (1.0 here),
# influence extent 0.4 (KP_extent 1.2 x radius / conv_radius 2.5 = 0.48 in the configs): ~1 kernel point per neighbour
v = torch.randn(15, 3, device=dev); kp = v / v.norm(dim=1, keepdim=True) * 0.66; kp[0] = 0
g = torch.randn(n, C_, device=dev)
for it in range(reps):
    x.grad = w.grad = None
    ops.kpconv(S, S, nb, x, w, kp, 0.4).backward(g)
torch.cuda.synchronize()
print("ok", n, nb.shape)
