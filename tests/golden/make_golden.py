"""Generates tests/golden/*.npz from the UNMODIFIED reference, in the build container only.

The reference ships no golden vectors for this path, so they are produced here by running
the reference itself on seeded synthetic inputs:
  * KPConv forward / dX / dW from the reference's own ``models.blocks.KPConv`` (PyTorch CPU, fp32);
  * neighbour matrices and subsampled clouds from the reference C++ cores (oracle/_ref);
  * a whole pyramid from the reference's ``datasets.common.PointCloudDataset.segmentation_inputs``
    with the two extension modules backed by oracle/_ref (random grid rotations captured).
/root/reference does not exist on the GPU box, so only the committed .npz files travel.

    python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from weasal_b200.synthetic import make_batch  # noqa: E402


def install_reference_import_harness():
    """make models.blocks / datasets.common importable without their GUI deps."""
    import torch

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm"):
        sys.modules.setdefault(name, types.ModuleType(name))
    ts = types.ModuleType("torch_scatter")
    sys.modules["torch_scatter"] = ts
    for n in ("datasets", "utils", "models", "kernels"):
        m = types.ModuleType(n)
        m.__path__ = [os.path.join(REF, n)]
        sys.modules[n] = m
    for n in ("cpp_wrappers", "cpp_wrappers.cpp_subsampling", "cpp_wrappers.cpp_neighbors"):
        m = types.ModuleType(n)
        m.__path__ = []
        sys.modules[n] = m
    gs = types.ModuleType("cpp_wrappers.cpp_subsampling.grid_subsampling")

    def subsample(points, features=None, classes=None, sampleDl=0.1, method="barycenters", verbose=0):
        return oracle.ref_subsample(points, features, classes, sampleDl)

    def subsample_batch(points, batches, features=None, classes=None, sampleDl=0.1, method="barycenters",
                        max_p=0, verbose=0):
        return oracle.ref_subsample_batch(points, batches, features, classes, sampleDl, max_p)

    gs.subsample, gs.subsample_batch = subsample, subsample_batch
    sys.modules[gs.__name__] = gs
    sys.modules["cpp_wrappers.cpp_subsampling"].grid_subsampling = gs
    rn = types.ModuleType("cpp_wrappers.cpp_neighbors.radius_neighbors")

    def batch_query(queries, supports, q_batches, s_batches, radius=0.1):
        return oracle.ref_batch_neighbors(queries, supports, q_batches, s_batches, radius)

    rn.batch_query = batch_query
    sys.modules[rn.__name__] = rn
    sys.modules["cpp_wrappers.cpp_neighbors"].radius_neighbors = rn
    sys.path.insert(0, REF)
    os.chdir(REF)  # load_kernels uses the relative path kernels/dispositions (kernel_points.py:410)
    torch.Tensor.cuda = lambda self, *a, **k: self


def golden_kpconv():
    import torch
    from models.blocks import KPConv

    cases = [  # (name, Cin, Cout, layer radius multiple, in_radius)
        ("c4_32", 4, 32, 1, 5.0),
        ("c16_16", 16, 16, 1, 5.0),
        ("c64_64", 64, 64, 2, 6.0),
        ("c32_128", 32, 128, 2, 6.0),
        ("c3_64", 3, 64, 1, 5.0),
    ]
    out = {}
    for ci, (name, cin, cout, mult, in_r) in enumerate(cases):
        np.random.seed(100 + ci)
        torch.manual_seed(100 + ci)
        b = make_batch("vaihingen_pl", seed=10 + ci, batch_num=2, in_radius=in_r)
        pts, lens = b["points"], b["lengths"]
        dl = 0.24 * mult
        if mult > 1:
            pts, lens = oracle.ref_subsample_batch(pts, lens, sampleDl=dl)
        radius = dl * 2.5
        extent = radius * 1.0 / 2.5
        idx = oracle.ref_batch_neighbors(pts, pts, lens, lens, radius).astype(np.int64)
        conv = KPConv(15, 3, cin, cout, extent, radius)
        x = torch.randn(len(pts), cin, dtype=torch.float32, requires_grad=True)
        q = torch.from_numpy(pts)
        y = conv(q, q, torch.from_numpy(idx), x)
        d_out = torch.randn_like(y)
        y.backward(d_out)
        out.update({
            f"{name}.q_pts": pts, f"{name}.s_pts": pts, f"{name}.idx": idx.astype(np.int32),
            f"{name}.x": x.detach().numpy(), f"{name}.weights": conv.weights.detach().numpy(),
            f"{name}.kernel_points": conv.kernel_points.detach().numpy(),
            f"{name}.extent": np.float32(extent), f"{name}.radius": np.float32(radius),
            f"{name}.out": y.detach().numpy(), f"{name}.d_out": d_out.numpy(),
            f"{name}.dx": x.grad.numpy(), f"{name}.dw": conv.weights.grad.numpy(),
        })
        print(name, pts.shape, idx.shape, float(y.abs().max()))
    # strided case: queries = next layer, supports = this layer, idx = pool neighbours
    np.random.seed(7)
    torch.manual_seed(7)
    b = make_batch("vaihingen_pl", seed=3, batch_num=2, in_radius=5.0)
    s_pts, s_len = b["points"], b["lengths"]
    q_pts, q_len = oracle.ref_subsample_batch(s_pts, s_len, sampleDl=0.48)
    idx = oracle.ref_batch_neighbors(q_pts, s_pts, q_len, s_len, 0.6).astype(np.int64)
    conv = KPConv(15, 3, 16, 16, 0.24, 0.6)
    x = torch.randn(len(s_pts), 16, requires_grad=True)
    y = conv(torch.from_numpy(q_pts), torch.from_numpy(s_pts), torch.from_numpy(idx), x)
    d_out = torch.randn_like(y)
    y.backward(d_out)
    name = "strided16"
    out.update({
        f"{name}.q_pts": q_pts, f"{name}.s_pts": s_pts, f"{name}.idx": idx.astype(np.int32),
        f"{name}.x": x.detach().numpy(), f"{name}.weights": conv.weights.detach().numpy(),
        f"{name}.kernel_points": conv.kernel_points.detach().numpy(),
        f"{name}.extent": np.float32(0.24), f"{name}.radius": np.float32(0.6),
        f"{name}.out": y.detach().numpy(), f"{name}.d_out": d_out.numpy(),
        f"{name}.dx": x.grad.numpy(), f"{name}.dw": conv.weights.grad.numpy(),
    })
    print(name, q_pts.shape, s_pts.shape, idx.shape)
    np.savez_compressed(os.path.join(HERE, "kpconv_ref.npz"), **out)


def golden_precompute():
    out = {}
    b = make_batch("vaihingen_pl", seed=21, batch_num=3, in_radius=6.0)
    pts, lens, feats, labels = b["points"], b["lengths"], b["features"], b["labels"].astype(np.int32)
    out["pts"], out["lens"], out["feats"], out["labels"] = pts, lens, feats, labels
    # radius search: conv (q = s), wired-in nanoflann path and the stable-order arbiter
    out["nbr_r0.6_nanoflann"] = oracle.ref_batch_neighbors(pts, pts, lens, lens, 0.6)
    out["nbr_r0.6_ordered"] = oracle.ref_batch_neighbors(pts, pts, lens, lens, 0.6, ordered=True)
    sp, sl = oracle.ref_subsample_batch(pts, lens, sampleDl=0.48)
    out["sub0.48_pts"], out["sub0.48_lens"] = sp, sl
    out["pool_r0.6_nanoflann"] = oracle.ref_batch_neighbors(sp, pts, sl, lens, 0.6)
    out["up_r1.2_nanoflann"] = oracle.ref_batch_neighbors(pts, sp, lens, sl, 1.2)
    # whole-cloud form with features + labels
    p2, f2, c2 = oracle.ref_subsample(pts, features=feats, classes=labels, sampleDl=0.9)
    out["sub0.9_pts"], out["sub0.9_feats"], out["sub0.9_classes"] = p2, f2, c2
    np.savez_compressed(os.path.join(HERE, "precompute_ref.npz"), **out)
    print("precompute", pts.shape, out["nbr_r0.6_nanoflann"].shape, sp.shape, p2.shape)


def golden_pyramid():
    from datasets.common import PointCloudDataset

    class Cfg:
        first_subsampling_dl = 0.24
        conv_radius = 2.5
        deform_radius = 6.0
        architecture = ['simple', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb_strided', 'resnetb',
                        'resnetb_strided', 'resnetb', 'nearest_upsample', 'unary', 'nearest_upsample', 'unary',
                        'nearest_upsample', 'unary']

    b = make_batch("vaihingen_pl", seed=31, batch_num=2, in_radius=7.0)
    ds = PointCloudDataset("x")
    ds.config = Cfg()
    ds.neighborhood_limits = [12, 20, 30, 40]
    np.random.seed(1234)
    li = ds.segmentation_inputs(b["points"], b["features"], b["labels"], b["lengths"])
    L = (len(li) - 2) // 5
    out = {"in_pts": b["points"], "in_lens": b["lengths"], "limits": np.asarray(ds.neighborhood_limits, np.int32),
           "seed": np.int64(1234), "L": np.int64(L)}
    for l in range(L):
        out[f"points{l}"] = li[l]
        out[f"neighbors{l}"] = li[L + l].astype(np.int32)
        out[f"pools{l}"] = li[2 * L + l].astype(np.int32)
        out[f"upsamples{l}"] = li[3 * L + l].astype(np.int32)
        out[f"lengths{l}"] = np.asarray(li[4 * L + l], np.int32)
        print(l, li[l].shape, li[L + l].shape, li[2 * L + l].shape, li[3 * L + l].shape)
    np.savez_compressed(os.path.join(HERE, "pyramid_ref.npz"), **out)


if __name__ == "__main__":
    oracle.build()
    install_reference_import_harness()
    golden_kpconv()
    golden_precompute()
    golden_pyramid()
    os.chdir(ROOT)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
