"""Per-CTA phase timeline of the last KPConv forward-type launch (WEASAL_KP_TRACE=1): where a CTA's time goes."""
import ctypes as C, os, sys
os.environ["WEASAL_KP_TRACE"] = "1"
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from weasal_b200 import ops, _lib
from weasal_b200.synthetic import make_als_tile
C_ = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
pts, _, _ = make_als_tile(1, 50.0, 40.0)
P = torch.from_numpy(pts).to(dev); L = np.array([len(pts)], np.int32)
sp, sl = ops.grid_subsample(P, L, sampleDl=0.4, order="first")
S = sp.contiguous(); Ls = np.array([len(S)], np.int32)
nb = ops.batch_query(S, S, Ls, Ls, 1.0, dtype=torch.int32, cap_hint=64)
n = len(S)
x = torch.randn(n, C_, device=dev)
w = torch.randn(15, C_, C_, device=dev) / C_ ** 0.5
v = torch.randn(15, 3, device=dev); kp = v / v.norm(dim=1, keepdim=True) * 0.66 * 0.4; kp[0] = 0
lib = _lib.lib()
lib.kp_debug_trace_read.argtypes = [C.c_void_p, C.c_int]
for it in range(3):
    y = ops.kpconv(S, S, nb, x, w, kp, 0.4)
torch.cuda.synchronize()
buf = np.zeros((1 << 16, 16), np.int64)
nc = lib.kp_debug_trace_read(buf.ctypes.data, 1 << 16)
t = buf[:nc].astype(np.float64)
t0 = t[:, 0].min()
names = ["start", "prologue", "efull0", "produced0", "fenced0", "arrived0", "efull1", "produced1", "fenced1", "arrived1", "mma_done", "epi_start", "epi_end", "exit", "eload_issue0", "efull0_seen"]
print("CTAs", nc, "kernel span us", (t[:, 13].max() - t0) / 1e3)
rel = t - t[:, :1]
for i, nm in enumerate(names):
    col = rel[:, i][t[:, i] > 0]
    if len(col):
        print(f"{nm:10s} median {np.median(col) / 1e3:8.2f} us   p90 {np.percentile(col, 90) / 1e3:8.2f}   (n={len(col)})")
starts = np.sort(t[:, 0] - t0) / 1e3
print("CTA start times us: p10 %.1f p50 %.1f p90 %.1f max %.1f" % tuple(np.percentile(starts, [10, 50, 90, 100])))
dur = (t[:, 13] - t[:, 0]) / 1e3
print("CTA duration us: median %.2f p90 %.2f" % (np.median(dur), np.percentile(dur, 90)))
