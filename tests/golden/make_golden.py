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

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from weasal_b200.synthetic import make_als_tile, make_batch  # noqa: E402


def install_reference_import_harness():
    """SURVEY.md Appendix A.2 (shared with the tests: oracle/ref_harness.py), extension modules backed by oracle/_ref."""
    from oracle import ref_harness
    ref_harness.install(REF, backend="oracle_ref")


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


PL_ARCH = ['simple', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb_strided', 'resnetb',
           'resnetb_strided', 'resnetb', 'nearest_upsample', 'unary', 'nearest_upsample', 'unary', 'nearest_upsample',
           'unary', 'nearest_upsample', 'unary']


def golden_pyramid_dales():
    """The DALES PseudoLabel walk (train_DALES_PseudoLabel.py:98-121: dl 0.4, 5 layers, conv_radius 2.5) on spheres cut
    from a tile that was first subsampled at dl like the dataset does (DALES_PseudoLabel.py load_subsampled_clouds),
    with neighbourhood limits that bite on some layers (crop ties are exercised). Index matrices are stored as uint16
    (every layer has < 65535 points)."""
    from datasets.common import PointCloudDataset
    from weasal_b200.synthetic import extract_spheres, make_als_tile, pick_centres

    class Cfg:
        first_subsampling_dl = 0.4
        conv_radius = 2.5
        deform_radius = 6.0
        architecture = PL_ARCH

    tile, _, _ = make_als_tile(41, 60.0, 14.0)
    sub = oracle.ref_subsample(tile, sampleDl=0.4)
    centres = pick_centres(sub, 2, 12.0, 42)
    pts, lens, _ = extract_spheres(sub, centres, 12.0)
    ds = PointCloudDataset("x")
    ds.config = Cfg()
    ds.neighborhood_limits = [26, 40, 48, 50, 40]
    np.random.seed(4321)
    feats = np.ones((len(pts), 1), np.float32)
    li = ds.segmentation_inputs(pts, feats, np.zeros(len(pts), np.int64), lens)
    L = (len(li) - 2) // 5
    out = {"in_pts": pts, "in_lens": lens, "limits": np.asarray(ds.neighborhood_limits, np.int32),
           "seed": np.int64(4321), "L": np.int64(L)}
    for l in range(L):
        assert len(li[l]) < 65535
        out[f"points{l}"] = li[l]
        out[f"neighbors{l}"] = li[L + l].astype(np.uint16)
        out[f"pools{l}"] = li[2 * L + l].astype(np.uint16)
        out[f"upsamples{l}"] = li[3 * L + l].astype(np.uint16)
        out[f"lengths{l}"] = np.asarray(li[4 * L + l], np.int32)
        print("dales", l, li[l].shape, li[L + l].shape, li[2 * L + l].shape, li[3 * L + l].shape)
    np.savez_compressed(os.path.join(HERE, "pyramid_dales_ref.npz"), **out)


WIDE_CASES = [("w128_128", 128, 128), ("w256_256", 256, 256), ("w512_512", 512, 512), ("w512_256", 512, 256),
              ("w256_32", 256, 32), ("w32_256", 32, 256)]


def wide_inputs(seed, n, cin, cout):
    """x, weights, d_out of a wide case from numpy's PCG64 stream (the test regenerates them: only outputs are stored)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, cin), dtype=np.float32)
    w = (rng.standard_normal((15, cin, cout), dtype=np.float32) / np.float32(np.sqrt(cin * 4.0))).astype(np.float32)
    d_out = rng.standard_normal((n, cout), dtype=np.float32)
    return x, w, d_out


def golden_kpconv_wide():
    """The wide layers of the DALES net (128 -> 512 channels, train_DALES_PseudoLabel.py:120) and of the WL attention
    heads (256->32, 32->256, 512->256: models/blocks.py:778-784, 844-849, 909, 981) from the reference's own KPConv on
    one shared geometry of > 5000 points (42 tiles of 128: the multi-wave / split-reduction paths are active). To keep the
    fixture small only the geometry, the kernel points and SAMPLES of the outputs are stored (256 rows of out and dX,
    every 8th / 4th row and column of dW); x, weights and d_out are regenerated from a seeded numpy stream."""
    import torch
    from models.blocks import KPConv

    tile, _, _ = make_als_tile(51, 64.0, 14.0)
    dl = 0.96
    pts = oracle.ref_subsample(tile, sampleDl=dl)
    lens = np.asarray([len(pts)], np.int32)
    radius = dl * 2.5
    extent = radius / 2.5
    idx = oracle.ref_batch_neighbors(pts, pts, lens, lens, radius).astype(np.int64)
    n = len(pts)
    assert 5000 < n < 65535, n
    rows = np.sort(np.random.default_rng(9).choice(n, 256, replace=False))
    out = {"pts": pts, "idx": idx.astype(np.uint16), "rows": rows.astype(np.int32), "extent": np.float32(extent),
           "radius": np.float32(radius)}
    q = torch.from_numpy(pts)
    for ci, (name, cin, cout) in enumerate(WIDE_CASES):
        np.random.seed(200 + ci)
        torch.manual_seed(200 + ci)
        conv = KPConv(15, 3, cin, cout, extent, radius)
        x_np, w_np, do_np = wide_inputs(300 + ci, n, cin, cout)
        with torch.no_grad():
            conv.weights.copy_(torch.from_numpy(w_np))
        x = torch.from_numpy(x_np).requires_grad_(True)
        y = conv(q, q, torch.from_numpy(idx), x)
        y.backward(torch.from_numpy(do_np))
        sc, so = max(cin // 64, 1), max(cout // 64, 1)
        out.update({f"{name}.kernel_points": conv.kernel_points.detach().numpy(), f"{name}.seed": np.int64(300 + ci),
                    f"{name}.out_rows": y.detach().numpy()[rows], f"{name}.dx_rows": x.grad.numpy()[rows],
                    f"{name}.dw_sub": conv.weights.grad.numpy()[:, ::sc, ::so].copy(),
                    f"{name}.dw_stride": np.asarray([sc, so], np.int32),
                    f"{name}.out_absmax": np.float32(y.detach().abs().max()),
                    f"{name}.dx_absmax": np.float32(x.grad.abs().max()),
                    f"{name}.dw_absmax": np.float32(conv.weights.grad.abs().max())})
        print(name, n, idx.shape, float(y.abs().max()))
    np.savez_compressed(os.path.join(HERE, "kpconv_wide_ref.npz"), **out)


if __name__ == "__main__":
    oracle.build()
    install_reference_import_harness()
    only_new = "--new" in sys.argv  # keep the round-1 fixtures byte-identical: generate the added ones only
    if not only_new:
        golden_kpconv()
        golden_precompute()
        golden_pyramid()
    golden_pyramid_dales()
    golden_kpconv_wide()
    os.chdir(ROOT)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
