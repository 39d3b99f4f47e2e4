"""Per-kernel device times (the library's own CUDA-event timers, kp_profile_enable) of one KPConv forward + backward on
a subsampled synthetic tile, C = 64 / 128 (KP_TIMES_C=64,256,... for other widths), random and shell-shaped kernel points. WEASAL_B200_LIB selects an experiment
build of the library (e.g. one compiled with -DKP_ASSEMBLE_V=1) for A/B comparisons.

    python tools/kpconv_kernel_times.py
"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from weasal_b200 import ops, _lib
from weasal_b200.synthetic import make_als_tile
dev = torch.device("cuda", 0)
pts, _, _ = make_als_tile(1, 50.0, 40.0)
P = torch.from_numpy(pts).to(dev)
L = np.array([len(pts)], np.int32)
sp, sl = ops.grid_subsample(P, L, sampleDl=0.4, order="first")
S = sp.contiguous(); Ls = np.array([len(S)], np.int32)
nb = ops.batch_query(S, S, Ls, Ls, 1.0, dtype=torch.int32, cap_hint=64)
n = len(S)
lib = _lib.lib()
for Cc in [int(c) for c in os.environ.get("KP_TIMES_C", "64,128").split(",")]:
    for kpmode in os.environ.get("KP_TIMES_KP", "randn,shell").split(","):
        x = torch.randn(n, Cc, device=dev, requires_grad=True)
        w = (torch.randn(15, Cc, Cc, device=dev) / Cc ** 0.5).requires_grad_(True)
        if kpmode == "randn":
            kp = torch.randn(15, 3, device=dev) * 0.4
        else:
            v = torch.randn(15, 3, device=dev); kp = v / v.norm(dim=1, keepdim=True) * 0.66; kp[0] = 0
        g = torch.randn(n, Cc, device=dev)
        for it in range(3):
            if it == 2:
                torch.cuda.synchronize(); lib.kp_profile_enable(1)
            x.grad = w.grad = None
            ops.kpconv(S, S, nb, x, w, kp, 0.4).backward(g)
        torch.cuda.synchronize(); lib.kp_profile_enable(0)
        buf = C.create_string_buffer(1 << 16); lib.kp_profile_read(buf, len(buf))
        print(os.environ.get("WEASAL_B200_LIB", "default")[-12:], "C", Cc, kpmode, "n", n, "H", nb.shape[1], {l.split()[0]: round(float(l.split()[2]), 3) for l in buf.value.decode().splitlines()})
