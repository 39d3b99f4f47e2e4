"""Error of the KPConv weight gradient against an fp64 evaluation of the reference expression (models/blocks.py:277-374)
on a subsampled synthetic tile, and the kernel's device time. WEASAL_DW_TMA=0 selects the thread-staged dOut path.

    python tools/dw_error.py [C]
"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from weasal_b200 import ops, _lib
from weasal_b200.synthetic import make_als_tile
dev = torch.device("cuda", 0)
Cc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pts, _, _ = make_als_tile(1, 30.0, 40.0)
P = torch.from_numpy(pts).to(dev)
L = np.array([len(pts)], np.int32)
sp, sl = ops.grid_subsample(P, L, sampleDl=0.4, order="first")
S = sp.contiguous(); Ls = np.array([len(S)], np.int32)
nb = ops.batch_query(S, S, Ls, Ls, 1.0, dtype=torch.int64, cap_hint=64)
n = len(S)
torch.manual_seed(0)
x = torch.randn(n, Cc, device=dev, requires_grad=True)
w = (torch.randn(15, Cc, Cc, device=dev) / Cc ** 0.5).requires_grad_(True)
kp = torch.randn(15, 3, device=dev) * 0.4
g = torch.randn(n, Cc, device=dev)
lib = _lib.lib()
for it in range(3):
    if it == 2:
        torch.cuda.synchronize(); lib.kp_profile_enable(1)
    x.grad = w.grad = None
    ops.kpconv(S, S, nb, x, w, kp, 0.4).backward(g)
torch.cuda.synchronize(); lib.kp_profile_enable(0)
buf = C.create_string_buffer(1 << 16); lib.kp_profile_read(buf, len(buf))
times = {l.split()[0]: round(float(l.split()[2]), 3) for l in buf.value.decode().splitlines()}
# fp64 reference of dW: WF[i,k,c] = sum_h w_ikh x[idx_ih, c];  dW[k] = WF[:,k,:]^T g
Sd = torch.cat([S.double(), torch.full((1, 3), 1e6, device=dev, dtype=torch.float64)])
xd = torch.cat([x.detach().double(), torch.zeros(1, Cc, device=dev, dtype=torch.float64)])
dw_ref = torch.zeros(15, Cc, Cc, device=dev, dtype=torch.float64)
for a in range(0, n, 4096):
    b = min(n, a + 4096)
    idx = nb[a:b].long().clamp(max=n)
    idx[nb[a:b] < 0] = n
    rel = Sd[idx] - S[a:b].double()[:, None, :]
    d = (rel[:, :, None, :] - kp.double()[None, None]).norm(dim=3)
    wgt = (1 - d / 0.4).clamp(min=0).transpose(1, 2)          # [m, K, H]
    wf = wgt @ xd[idx]                                        # [m, K, C]
    dw_ref += torch.einsum("mkc,mo->kco", wf, g[a:b].double())
err = (w.grad.double() - dw_ref)
print("DW_TMA", os.environ.get("WEASAL_DW_TMA", "1"), "C", Cc, "n", n, "rel_max", float(err.abs().max() / dw_ref.abs().max()),
      "rel_l2", float(err.norm() / dw_ref.norm()), "mean_signed_ratio", float((w.grad.double() * dw_ref).sum() / (dw_ref * dw_ref).sum()) - 1, times)
