"""Operator sweep (BASELINE.json configs[4]): grid subsampling + batch radius search on 1e5..1e7 raw points and KPConv
(K = 15) at Cin = Cout in {64, 128, 256}, with achieved GB/s against the SURVEY.md section-8d algorithmic bytes.
Writes one JSON object per line to stdout."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from weasal_b200 import ops  # noqa: E402
from weasal_b200.synthetic import make_als_tile  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = float(PEAK.get("hbm_gbs", 6650.0))
dev = torch.device("cuda", 0)


def timeit(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


def precompute_sweep(sizes):
    for n_raw in sizes:
        extent = float(np.sqrt(n_raw / 40.0))  # DALES-like raw density, 40 pts/m^2
        pts, _, _ = make_als_tile(1, extent, 40.0)
        P = torch.from_numpy(pts).to(dev)
        L = np.array([len(pts)], np.int32)
        for order in ("reference", "first"):
            ms, (sp, sl) = timeit(lambda: ops.grid_subsample(P, L, sampleDl=0.4, order=order))
            b = 12 * len(pts) + 12 * len(sp) + 8
            print(json.dumps({"op": "grid_subsample", "order": order, "N": len(pts), "M": int(len(sp)), "ms": ms,
                              "Mpts_per_s": len(pts) / ms / 1e3, "algorithmic_GBs": b / ms / 1e6,
                              "frac_hbm": b / ms / 1e6 / HBM}), flush=True)
        S = sp.contiguous()
        Ls = np.array([len(S)], np.int32)
        ms, nb = timeit(lambda: ops.batch_query(S, S, Ls, Ls, 1.0, dtype=torch.int32, cap_hint=64))
        b = 24 * len(S) + 8 + 4 * len(S) * nb.shape[1]
        print(json.dumps({"op": "batch_query", "N": int(len(S)), "Hmax": int(nb.shape[1]), "ms": ms,
                          "Mqueries_per_s": len(S) / ms / 1e3, "algorithmic_GBs": b / ms / 1e6,
                          "frac_hbm": b / ms / 1e6 / HBM}), flush=True)
        yield S, Ls, nb


def kpconv_sweep(S, Ls, nb):
    n = len(S)
    for C in (64, 128, 256):
        x = torch.randn(n, C, device=dev, requires_grad=True)
        w = (torch.randn(15, C, C, device=dev) / C ** 0.5).requires_grad_(True)
        v = torch.randn(15, 3, device=dev)  # the reference's layout: centre + shell at 0.66 x conv radius (1.0), extent 0.4
        kp = v / v.norm(dim=1, keepdim=True) * 0.66
        kp[0] = 0
        ms_f, y = timeit(lambda: ops.kpconv(S, S, nb, x, w, kp, 0.4), warm=3, reps=5)
        g = torch.randn_like(y)

        def fb():
            yy = ops.kpconv(S, S, nb, x, w, kp, 0.4)
            yy.backward(g)
            return yy
        ms_fb, _ = timeit(fb, warm=3, reps=5)  # (the first calls size allocator pools of a few hundred MB)
        H = nb.shape[1]
        bytes_f = 4 * n * H + 24 * n + 4 * n * C * 2 + 4 * 15 * C * C
        flops = 2 * n * 15 * C * C
        print(json.dumps({"op": "kpconv", "N": n, "H": int(H), "C": C, "fwd_ms": ms_f, "fwd_bwd_ms": ms_fb,
                          "fwd_algorithmic_GBs": bytes_f / ms_f / 1e6, "fwd_frac_hbm": bytes_f / ms_f / 1e6 / HBM,
                          "fwd_contraction_TFLOPs": flops / ms_f / 1e9,
                          "fwd_Mpts_per_s": n / ms_f / 1e3}), flush=True)
        del x, w, y, g


if __name__ == "__main__":
    sizes = [int(float(a)) for a in sys.argv[1:]] or [100_000, 1_000_000, 10_000_000]
    for S, Ls, nb in precompute_sweep(sizes):
        if len(S) <= 600_000:
            kpconv_sweep(S, Ls, nb)
