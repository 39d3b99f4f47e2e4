"""Bring-up probe for the GPU box: runs each component against the oracle and prints diagnostics instead of
stopping at the first mismatch. Not part of the test suite."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from weasal_b200.synthetic import make_als_tile, make_batch  # noqa: E402


def section(name):
    print(f"\n=== {name} ===", flush=True)


def main():
    import torch
    stages = sys.argv[1:] or ["pre", "simt", "tc", "timing"]
    print(torch.cuda.get_device_name(0), torch.version.cuda, stages)
    from weasal_b200 import grid_subsampling as gs, ops, radius_neighbors as rn

    b = make_batch("vaihingen_pl", seed=0, batch_num=2, in_radius=6.0)
    P, L = b["points"], b["lengths"]
    section("radius search")
    try:
        if "pre" not in stages:
            raise KeyboardInterrupt
        want = oracle.batch_neighbors(P, P, L, L, 0.6)
        got = rn.batch_query(P, P, L, L, radius=0.6)
        print("shape", got.shape, want.shape, "equal", np.array_equal(got, want))
        if got.shape == want.shape and not np.array_equal(got, want):
            bad = np.nonzero((got != want).any(1))[0]
            print("bad rows", len(bad), bad[:5], got[bad[0]], want[bad[0]])
    except KeyboardInterrupt:
        pass
    except Exception:
        traceback.print_exc()
    section("grid subsample")
    try:
        if "pre" not in stages:
            raise KeyboardInterrupt
        for order in ("first", "reference"):
            wp, wl = oracle.grid_subsample_batch(P, L, sampleDl=0.48, order=order)
            gp, gl = gs.subsample_batch(P, L, sampleDl=0.48, order=order)
            print(order, gl, wl, gp.shape, wp.shape, "equal", gp.shape == wp.shape and np.array_equal(gp, wp))
            if gp.shape == wp.shape and not np.array_equal(gp, wp):
                same_set = np.array_equal(gp[np.lexsort(gp.T)], wp[np.lexsort(wp.T)])
                print("   same set:", same_set, "first diff row", np.nonzero((gp != wp).any(1))[0][:5])
    except KeyboardInterrupt:
        pass
    except Exception:
        traceback.print_exc()
    section("kpconv")
    g = np.load(os.path.join(ROOT, "tests", "golden", "kpconv_ref.npz"))
    for impl in [s for s in ("simt", "tc") if s in stages]:
        os.environ["WEASAL_KPCONV_IMPL"] = impl
        for name in ["c4_32", "c16_16", "c64_64", "c32_128", "c3_64", "strided16"]:
            try:
                a = {k.split(".", 1)[1]: g[k] for k in g.files if k.startswith(name + ".")}
                q = torch.from_numpy(a["q_pts"]).cuda(); s = torch.from_numpy(a["s_pts"]).cuda()
                idx = torch.from_numpy(a["idx"].astype(np.int64)).cuda()
                x = torch.from_numpy(a["x"]).cuda().requires_grad_(True)
                w = torch.from_numpy(a["weights"]).cuda().requires_grad_(True)
                kp = torch.from_numpy(a["kernel_points"]).cuda()
                out = ops.kpconv(q, s, idx, x, w, kp, float(a["extent"]))
                torch.cuda.synchronize()
                rel = lambda u, v: float(np.abs(u - v).max() / np.abs(v).max())
                e_f = rel(out.detach().cpu().numpy(), a["out"])
                out.backward(torch.from_numpy(a["d_out"]).cuda())
                torch.cuda.synchronize()
                e_x = rel(x.grad.cpu().numpy(), a["dx"]); e_w = rel(w.grad.cpu().numpy(), a["dw"])
                print(f"{impl:5s} {name:10s} fwd {e_f:.2e} dx {e_x:.2e} dw {e_w:.2e}", flush=True)
            except Exception:
                traceback.print_exc()
    os.environ.pop("WEASAL_KPCONV_IMPL", None)
    section("timing (VPL batch)")
    try:
        if "timing" not in stages:
            return
        bb = make_batch("vaihingen_pl", seed=0)
        dP = torch.from_numpy(bb["points"]).cuda(); LL = bb["lengths"]
        for _ in range(3):
            nb = ops.batch_query(dP, dP, LL, LL, 0.6)
        torch.cuda.synchronize(); t = time.time()
        for _ in range(10):
            nb = ops.batch_query(dP, dP, LL, LL, 0.6)
        torch.cuda.synchronize(); print("batch_query N=%d H=%d: %.3f ms" % (len(dP), nb.shape[1], (time.time() - t) * 100))
        for _ in range(3):
            sp, sl = ops.grid_subsample(dP, LL, sampleDl=0.48)
        torch.cuda.synchronize(); t = time.time()
        for _ in range(10):
            sp, sl = ops.grid_subsample(dP, LL, sampleDl=0.48)
        torch.cuda.synchronize(); print("grid_subsample -> %d: %.3f ms" % (len(sp), (time.time() - t) * 100))
        for cin, cout in ((16, 16), (64, 64), (128, 128)):
            x = torch.randn(len(dP), cin, device="cuda", requires_grad=True)
            w = torch.randn(15, cin, cout, device="cuda", requires_grad=True)
            kp = torch.randn(15, 3, device="cuda") * 0.25
            for impl in [s for s in ("simt", "tc") if s in stages]:
                os.environ["WEASAL_KPCONV_IMPL"] = impl
                for _ in range(3):
                    y = ops.kpconv(dP, dP, nb, x, w, kp, 0.24)
                torch.cuda.synchronize(); t = time.time()
                for _ in range(10):
                    y = ops.kpconv(dP, dP, nb, x, w, kp, 0.24)
                torch.cuda.synchronize(); tf = (time.time() - t) * 100
                g_ = torch.randn_like(y)
                t = time.time()
                for _ in range(10):
                    y = ops.kpconv(dP, dP, nb, x, w, kp, 0.24); y.backward(g_)
                torch.cuda.synchronize(); tb = (time.time() - t) * 100
                print(f"kpconv {impl} {cin}->{cout}: fwd {tf:.3f} ms, fwd+bwd {tb:.3f} ms", flush=True)
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    main()
