"""GPU parity tests (run with ``-m gpu`` on the B200 box). Everything goes through the C-ABI library; the CPU oracle
is only the checker. Bars: bit-exact for indices / voxel membership / barycentres; 1e-3 relative
(max-abs error over max-abs reference, and L2 norm-wise) for KPConv features and gradients, TF32 operands with fp32
accumulation against the f64-accumulating oracle and the reference's own fp32 outputs."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, assert_same_up_to_ties, load_case
from weasal_b200.synthetic import make_als_tile, make_batch

pytestmark = pytest.mark.gpu

KP_TOL = 1e-3


def rel_max(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def pre_golden():
    return np.load(os.path.join(GOLDEN, "precompute_ref.npz"))


@pytest.fixture(scope="module")
def kp_golden():
    return np.load(os.path.join(GOLDEN, "kpconv_ref.npz"))


# ------------------------------------------------------------------------------------------------------ radius search
def test_batch_query_golden(pre_golden, torch_cuda):
    from weasal_b200 import radius_neighbors as rn
    g = pre_golden
    pts, lens = g["pts"], g["lens"]
    got = rn.batch_query(pts, pts, lens, lens, radius=0.6)
    assert got.dtype == np.int32
    assert np.array_equal(got, g["nbr_r0.6_ordered"])
    sp, sl = g["sub0.48_pts"], g["sub0.48_lens"]
    pool = rn.batch_query(sp, pts, sl, lens, radius=0.6)
    assert pool.shape == g["pool_r0.6_nanoflann"].shape
    assert np.array_equal(np.sort(pool, 1), np.sort(g["pool_r0.6_nanoflann"], 1))
    assert np.array_equal(pool, oracle.batch_neighbors(sp, pts, sl, lens, 0.6))
    up = rn.batch_query(pts, sp, lens, sl, radius=1.2)
    assert np.array_equal(up, oracle.batch_neighbors(pts, sp, lens, sl, 1.2))


@pytest.mark.parametrize("seed,radius,nb", [(0, 0.6, 1), (1, 0.6, 4), (2, 1.2, 3), (3, 2.4, 2), (4, 0.05, 2)])
def test_batch_query_vs_oracle(seed, radius, nb, torch_cuda):
    from weasal_b200 import radius_neighbors as rn
    b = make_batch("vaihingen_pl", seed=seed, batch_num=nb, in_radius=6.0)
    P, L = b["points"], b["lengths"]
    want = oracle.batch_neighbors(P, P, L, L, radius)
    got = rn.batch_query(P, P, L, L, radius=radius)
    assert np.array_equal(got, want)
    assert (got[:, 0] == np.arange(len(P))).all()  # self is neighbour 0 when q is s (d2 = 0)


def test_batch_query_ragged_and_empty(torch_cuda):
    from weasal_b200 import radius_neighbors as rn
    rng = np.random.default_rng(0)
    q = rng.uniform(0, 3, (50, 3)).astype(np.float32)
    s = rng.uniform(0, 3, (70, 3)).astype(np.float32)
    qb, sb = np.array([20, 0, 30], np.int32), np.array([0, 40, 30], np.int32)  # empty elements on both sides
    got = rn.batch_query(q, s, qb, sb, radius=1.0)
    assert np.array_equal(got, oracle.batch_neighbors(q, s, qb, sb, 1.0))
    assert (got[:20] == 70).all()  # queries whose batch element has no supports: all shadow
    with pytest.raises(RuntimeError, match="^Error$"):
        rn.batch_query(q, s + 100.0, qb, sb, radius=0.01)  # no neighbour at all -> the reference's "Error"
    with pytest.raises(RuntimeError, match="^Error$"):
        rn.batch_query(np.zeros((0, 3), np.float32), s, np.array([0], np.int32), np.array([70], np.int32), radius=1.0)
    with pytest.raises(RuntimeError, match="query.shape"):
        rn.batch_query(q[:, :2], s, qb, sb, radius=1.0)


def test_batch_query_dense_rows_exceed_first_capacity(torch_cuda):
    from weasal_b200 import radius_neighbors as rn
    rng = np.random.default_rng(1)
    s = rng.uniform(0, 1, (900, 3)).astype(np.float32)  # ~250 neighbours per query: beyond the first 64-wide buffer
    L = np.array([900], np.int32)
    got = rn.batch_query(s, s, L, L, radius=0.45)
    assert got.shape[1] > 64
    assert np.array_equal(got, oracle.batch_neighbors(s, s, L, L, 0.45))


def test_batch_query_duplicate_points_tie_break(torch_cuda):
    from weasal_b200 import radius_neighbors as rn
    base = np.random.default_rng(2).uniform(0, 2, (100, 3)).astype(np.float32)
    s = np.concatenate([base, base, base[:30]], 0)  # exact duplicates => exact d2 ties
    L = np.array([len(s)], np.int32)
    got = rn.batch_query(s, s, L, L, radius=0.5)
    assert np.array_equal(got, oracle.batch_neighbors(s, s, L, L, 0.5))


def test_batch_query_device_limit_and_int64(torch_cuda):
    torch = torch_cuda
    from weasal_b200 import ops
    b = make_batch("vaihingen_pl", seed=5, batch_num=2, in_radius=6.0)
    P, L = b["points"], b["lengths"]
    want = oracle.batch_neighbors(P, P, L, L, 0.9)
    dP = torch.from_numpy(P).cuda()
    full = ops.batch_query(dP, dP, L, L, 0.9)
    assert full.dtype == torch.int64 and np.array_equal(full.cpu().numpy(), want)
    lim = 7
    crop = ops.batch_query(dP, dP, L, L, 0.9, limit=lim, dtype=torch.int32)
    assert np.array_equal(crop.cpu().numpy(), want[:, :lim])  # big_neighborhood_filter keeps the closest `limit`


# -------------------------------------------------------------------------------------------------- grid subsampling
def test_subsample_golden(pre_golden, torch_cuda):
    from weasal_b200 import grid_subsampling as gs
    g = pre_golden
    sp, sl = gs.subsample_batch(g["pts"], g["lens"], sampleDl=0.48)
    assert np.array_equal(sl, g["sub0.48_lens"])
    assert np.array_equal(sp, g["sub0.48_pts"])
    p2, f2, c2 = gs.subsample(g["pts"], features=g["feats"], classes=g["labels"], sampleDl=0.9)
    assert np.array_equal(p2, g["sub0.9_pts"])
    assert np.array_equal(f2, g["sub0.9_feats"])
    assert c2.shape == g["sub0.9_classes"].shape and np.array_equal(c2, g["sub0.9_classes"])


@pytest.mark.parametrize("seed,dl,n_side", [(0, 0.24, 25.0), (1, 0.4, 40.0), (2, 1.3, 40.0), (3, 2.4, 60.0),
                                           (4, 0.3, 90.0), (5, 0.8, 130.0)])  # the last two: > 60k points, the device-wide order replay
def test_subsample_vs_oracle(seed, dl, n_side, torch_cuda):
    from weasal_b200 import grid_subsampling as gs
    pts, inten, lab = make_als_tile(seed, n_side, 12.0)
    feats = np.stack([inten, pts[:, 2], pts[:, 0]], 1)
    for order in ("reference", "first"):
        want = oracle.grid_subsample(pts, features=feats, classes=lab, sampleDl=dl, order=order)
        got = gs.subsample(pts, features=feats, classes=lab, sampleDl=dl, order=order)
        for w, g_ in zip(want, got):
            assert w.shape == g_.shape and np.array_equal(w, g_)
    only_pts = gs.subsample(pts, sampleDl=dl)
    assert np.array_equal(only_pts, oracle.grid_subsample(pts, sampleDl=dl))


def test_subsample_batch_max_p_and_rotation(torch_cuda):
    from weasal_b200 import grid_subsampling as gs
    from weasal_b200.pyramid import random_grid_rotations
    b = make_batch("vaihingen_pl", seed=7, batch_num=3, in_radius=5.0)
    P, L = b["points"], b["lengths"]
    want_p, want_l = oracle.grid_subsample_batch(P, L, sampleDl=0.5, max_p=200)
    got_p, got_l = gs.subsample_batch(P, L, sampleDl=0.5, max_p=200)
    assert np.array_equal(got_l, want_l) and np.array_equal(got_p, want_p)
    np.random.seed(11)
    R = random_grid_rotations(len(L))
    rot = P.copy()
    i0 = 0
    for bi, n in enumerate(L):
        rot[i0:i0 + n] = oracle.rotate(P[i0:i0 + n], R[bi])
        i0 += n
    wp, wl = oracle.grid_subsample_batch(rot, L, sampleDl=0.5)
    i0 = 0
    for bi, n in enumerate(wl):
        wp[i0:i0 + n] = oracle.rotate(wp[i0:i0 + n], R[bi], transpose=True)
        i0 += n
    gp, gl = gs.subsample_batch(P, L, sampleDl=0.5, rot=R)
    assert np.array_equal(gl, wl) and np.array_equal(gp, wp)


def test_subsample_labels_many_classes_and_ties(torch_cuda):
    from weasal_b200 import grid_subsampling as gs
    rng = np.random.default_rng(5)
    pts = rng.uniform(0, 4, (4000, 3)).astype(np.float32)
    lab = rng.integers(-3, 40, 4000).astype(np.int32)  # > 13 distinct labels per voxel: histogram rehash path
    wp, wc = oracle.grid_subsample(pts, classes=lab, sampleDl=1.0)
    gp, gc = gs.subsample(pts, classes=lab, sampleDl=1.0)
    assert np.array_equal(gp, wp) and np.array_equal(gc, wc)


def test_subsample_edge_cases(torch_cuda):
    from weasal_b200 import grid_subsampling as gs
    one = np.array([[1.5, -2.0, 3.0]], np.float32)
    assert np.array_equal(gs.subsample(one, sampleDl=0.3), oracle.grid_subsample(one, sampleDl=0.3))
    same = np.repeat(one, 50, 0)
    assert np.array_equal(gs.subsample(same, sampleDl=0.3), oracle.grid_subsample(same, sampleDl=0.3))
    with pytest.raises(RuntimeError, match="^Error$"):
        gs.subsample(np.zeros((0, 3), np.float32), sampleDl=0.3)
    with pytest.raises(RuntimeError, match="points.shape"):
        gs.subsample(np.zeros((5, 2), np.float32), sampleDl=0.3)
    rng = np.random.default_rng(3)
    P = rng.uniform(0, 5, (300, 3)).astype(np.float32)
    L = np.array([100, 0, 200], np.int32)  # an empty batch element in the middle
    wp, wl = oracle.grid_subsample_batch(P, L, sampleDl=0.8)
    gp, gl = gs.subsample_batch(P, L, sampleDl=0.8)
    assert np.array_equal(gl, wl) and np.array_equal(gp, wp)


def test_subsample_large_properties(torch_cuda):
    """Full-size (1M raw points, config 5) checks through size-independent properties."""
    from weasal_b200 import grid_subsampling as gs
    pts, inten, _ = make_als_tile(9, 250.0, 16.0)
    dl = 0.4
    sp, sf = gs.subsample(pts, features=np.ones((len(pts), 1), np.float32) * 2.0, sampleDl=dl)
    # one barycentre per occupied voxel, each inside (or on the edge of) its own voxel, features averaged exactly
    org = np.floor(pts.min(0) * np.float32(1 / np.float32(dl))) * np.float32(dl)
    vox = np.floor((pts - org) / np.float32(dl)).astype(np.int64)
    n_vox = len(np.unique(vox, axis=0))
    assert len(sp) == n_vox
    svox = np.floor((sp - org) / np.float32(dl) + 1e-3).astype(np.int64)
    assert len(np.unique(svox, axis=0)) >= 0.999 * n_vox
    assert np.array_equal(sf, np.full((len(sp), 1), 2.0, np.float32))
    first = gs.subsample(pts, sampleDl=dl, order="first")
    assert np.array_equal(first[np.lexsort(first.T)], sp[np.lexsort(sp.T)])  # same set, different order


# ------------------------------------------------------------------------------------------------------------ KPConv
def _run_kpconv(torch, a, idx_dtype, stride_pad=0):
    from weasal_b200 import ops
    dev = "cuda"
    q = torch.from_numpy(a["q_pts"]).to(dev)
    s = torch.from_numpy(a["s_pts"]).to(dev)
    idx_np = a["idx"].astype(np.int64 if idx_dtype == torch.int64 else np.int32)
    if stride_pad:
        buf = torch.full((idx_np.shape[0], idx_np.shape[1] + stride_pad), len(a["s_pts"]), dtype=idx_dtype, device=dev)
        buf[:, :idx_np.shape[1]] = torch.from_numpy(idx_np).to(dev)
        idx = buf[:, :idx_np.shape[1]]  # strided view, like the radius-search output
    else:
        idx = torch.from_numpy(idx_np).to(dev)
    x = torch.from_numpy(a["x"]).to(dev).requires_grad_(True)
    w = torch.from_numpy(a["weights"]).to(dev).requires_grad_(True)
    kp = torch.from_numpy(a["kernel_points"]).to(dev)
    out = ops.kpconv(q, s, idx, x, w, kp, float(a["extent"]))
    out.backward(torch.from_numpy(a["d_out"]).to(dev))
    torch.cuda.synchronize()
    return out.detach().cpu().numpy(), x.grad.cpu().numpy(), w.grad.cpu().numpy()


@pytest.mark.parametrize("name", ["c4_32", "c16_16", "c64_64", "c32_128", "c3_64", "strided16"])
@pytest.mark.parametrize("impl", ["tc", "simt"])
def test_kpconv_golden(kp_golden, name, impl, torch_cuda, monkeypatch):
    torch = torch_cuda
    monkeypatch.setenv("WEASAL_KPCONV_IMPL", impl)
    a = load_case(kp_golden, name)
    out, dx, dw = _run_kpconv(torch, a, torch.int64)
    o_out = oracle.kpconv_forward(a["q_pts"], a["s_pts"], a["idx"], a["x"], a["weights"], a["kernel_points"], float(a["extent"]))
    o_dx, o_dw = oracle.kpconv_backward(a["q_pts"], a["s_pts"], a["idx"], a["x"], a["weights"], a["kernel_points"],
                                        float(a["extent"]), a["d_out"])
    for got, ref, orc in ((out, a["out"], o_out), (dx, a["dx"], o_dx), (dw, a["dw"], o_dw)):
        assert got.shape == ref.shape
        assert rel_max(got, ref) < KP_TOL and rel_l2(got, ref) < KP_TOL      # vs the reference's own fp32 KPConv
        assert rel_max(got, orc) < KP_TOL and rel_l2(got, orc) < KP_TOL      # vs the f64-accumulating oracle


WIDE_CASES = [("w128_128", 128, 128), ("w256_256", 256, 256), ("w512_512", 512, 512), ("w512_256", 512, 256),
              ("w256_32", 256, 32), ("w32_256", 32, 256)]


def wide_inputs(seed, n, cin, cout):
    """tests/golden/make_golden.py wide_inputs: the same seeded numpy stream the fixture was generated with"""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, cin), dtype=np.float32)
    w = (rng.standard_normal((15, cin, cout), dtype=np.float32) / np.float32(np.sqrt(cin * 4.0))).astype(np.float32)
    d_out = rng.standard_normal((n, cout), dtype=np.float32)
    return x, w, d_out


@pytest.mark.parametrize("name,cin,cout", WIDE_CASES)
def test_kpconv_wide_golden(name, cin, cout, torch_cuda):
    """The wide layers (DALES 128..512 channels; the WL attention heads' 256->32, 32->256, 512->256) on 17k points = 134
    tiles of 128 (several waves, reduction split active) against the reference's own fp32 KPConv (sampled outputs in
    tests/golden/kpconv_wide_ref.npz): forward, dX and dW within 1e-3 of the reference's largest magnitude, and norm-wise
    over the sample."""
    torch = torch_cuda
    g = np.load(os.path.join(GOLDEN, "kpconv_wide_ref.npz"))
    pts, idx, rows = g["pts"], g["idx"].astype(np.int32), g["rows"].astype(np.int64)
    x, w, d_out = wide_inputs(int(g[f"{name}.seed"]), len(pts), cin, cout)
    a = dict(q_pts=pts, s_pts=pts, idx=idx, x=x, weights=w, kernel_points=g[f"{name}.kernel_points"],
             extent=g["extent"], d_out=d_out)
    out, dx, dw = _run_kpconv(torch, a, torch.int64)
    sc, so = (int(v) for v in g[f"{name}.dw_stride"])
    for got, ref, amax, what in ((out[rows], g[f"{name}.out_rows"], g[f"{name}.out_absmax"], "out"),
                                 (dx[rows], g[f"{name}.dx_rows"], g[f"{name}.dx_absmax"], "dx"),
                                 (dw[:, ::sc, ::so], g[f"{name}.dw_sub"], g[f"{name}.dw_absmax"], "dw")):
        assert got.shape == ref.shape
        err = float(np.abs(got - ref).max() / float(amax))
        assert err < KP_TOL, f"{name} {what}: {err}"
        assert rel_l2(got, ref) < KP_TOL, f"{name} {what}: l2 {rel_l2(got, ref)}"
    assert np.isfinite(out).all() and np.isfinite(dx).all() and np.isfinite(dw).all()


def test_kpconv_int32_indices_and_strided_rows(kp_golden, torch_cuda):
    torch = torch_cuda
    a = load_case(kp_golden, "c16_16")
    ref = _run_kpconv(torch, a, torch.int64)
    alt = _run_kpconv(torch, a, torch.int32, stride_pad=5)
    for r, t in zip(ref, alt):
        assert rel_max(t, r) < 1e-6


@pytest.mark.parametrize("cin,cout,H", [(8, 24, 9), (64, 256, 20), (256, 64, 33), (128, 512, 17), (20, 10, 40)])
def test_kpconv_random_shapes_vs_oracle(cin, cout, H, torch_cuda):
    torch = torch_cuda
    rng = np.random.default_rng(cin * 1000 + cout)
    nq, ns = 300, 450
    s = rng.uniform(0, 3, (ns, 3)).astype(np.float32)
    q = s[rng.choice(ns, nq, replace=False)] + rng.normal(0, 0.05, (nq, 3)).astype(np.float32)
    L = np.array([nq], np.int32), np.array([ns], np.int32)
    idx = oracle.batch_neighbors(q, s, L[0], L[1], 0.9)[:, :H]
    a = dict(q_pts=q, s_pts=s, idx=idx, x=rng.normal(size=(ns, cin)).astype(np.float32),
             weights=(rng.normal(size=(15, cin, cout)) / np.sqrt(cin)).astype(np.float32),
             kernel_points=(rng.normal(size=(15, 3)) * 0.4).astype(np.float32), extent=np.float32(0.36),
             d_out=rng.normal(size=(nq, cout)).astype(np.float32))
    out, dx, dw = _run_kpconv(torch, a, torch.int64)
    o_out = oracle.kpconv_forward(q, s, idx, a["x"], a["weights"], a["kernel_points"], 0.36)
    o_dx, o_dw = oracle.kpconv_backward(q, s, idx, a["x"], a["weights"], a["kernel_points"], 0.36, a["d_out"])
    assert rel_max(out, o_out) < KP_TOL and rel_l2(out, o_out) < KP_TOL
    assert rel_max(dx, o_dx) < KP_TOL and rel_l2(dx, o_dx) < KP_TOL
    assert rel_max(dw, o_dw) < KP_TOL and rel_l2(dw, o_dw) < KP_TOL


def test_kpconv_prefetched_lists_and_packed_weights_equal_the_plain_operator(torch_cuda):
    """The split form of the operator (lists built ahead by kp_kpconv_prepare_dev for forward / dX incl. the transposed
    table of a strided conv, weights packed by one kp_pack_weights_dev launch, dW and dX as separate calls) gives the
    plain operator's results (up to the order of the float atomics that join split reductions)."""
    torch = torch_cuda
    import ctypes as C
    from weasal_b200 import _lib, ops
    from weasal_b200.kpconv import KPConv
    from weasal_b200.plan import ConvPlans, ConvSpec, WeightPacker
    b = make_batch("vaihingen_pl", seed=12, batch_num=2, in_radius=8.0)
    dev = torch.device("cuda")
    P0 = torch.from_numpy(b["points"]).to(dev)
    lens = b["lengths"]
    P1, lens1 = ops.grid_subsample(P0, lens, sampleDl=0.48)
    P1 = P1.contiguous()
    conv_idx = ops.batch_query(P0, P0, lens, lens, 0.6)
    pool_idx = ops.batch_query(P1, P0, lens1, lens, 0.6)
    np.random.seed(3)
    torch.manual_seed(3)
    for strided, (q, idx) in ((False, (P0, conv_idx)), (True, (P1, pool_idx))):
        idx = idx.contiguous()
        conv = KPConv(15, 3, 16, 32, 0.24, 0.6).to(dev)
        x = torch.randn(P0.shape[0], 16, device=dev)
        g = torch.randn(q.shape[0], 32, device=dev)
        # plain operator
        x1 = x.clone().requires_grad_(True)
        y1 = conv(q, P0, idx, x1)
        y1.backward(g)
        dw1, dx1 = conv.weights.grad.clone(), x1.grad.clone()
        conv.weights.grad = None
        # split form
        sp = ConvSpec(0, strided, conv)
        n_cap = [P0.shape[0], P1.shape[0]]
        H = idx.shape[1]
        used = 15 * idx.shape[0] * H
        plans = ConvPlans([sp], n_cap, [H, H], [H, H], [used])
        buf = torch.zeros(plans.nbytes, dtype=torch.uint8, device=dev)
        jobs = plans.jobs([P0, P1], [idx, None], [idx, None], idx.dtype == torch.int64, buf)
        assert len(jobs) == (3 if strided else 2)
        plans.run(jobs, buf, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert int(buf[:4].view(torch.int32)[0]) == 0
        idx2 = idx.clone()
        plans.attach(buf, [idx2, None], [idx2, None])
        packer = WeightPacker(conv)
        packer.pack()
        x2 = x.clone().requires_grad_(True)
        y2 = conv(q, P0, idx2, x2)
        y2.backward(g)
        packer.release()
        # (same kernels on the same lists; reductions split across CTAs meet through float atomics, so last bits may differ)
        assert rel_max(y2.detach().cpu().numpy(), y1.detach().cpu().numpy()) < 1e-5, strided
        assert rel_max(x2.grad.cpu().numpy(), dx1.cpu().numpy()) < 1e-5, strided
        assert rel_max(conv.weights.grad.cpu().numpy(), dw1.cpu().numpy()) < 1e-5, strided
        # a list buffer that is too small raises the overflow flag instead of writing out of bounds
        small = ConvPlans([sp], n_cap, [H, H], [H, H], [64])
        buf_s = torch.zeros(small.nbytes, dtype=torch.uint8, device=dev)
        small.run(small.jobs([P0, P1], [idx, None], [idx, None], idx.dtype == torch.int64, buf_s), buf_s,
                  C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert int(buf_s[:4].view(torch.int32)[0]) == 1


def test_kpconv_all_shadow_rows_and_linearity(torch_cuda):
    torch = torch_cuda
    from weasal_b200 import ops
    rng = np.random.default_rng(0)
    ns, nq, cin, cout, H = 200, 130, 16, 32, 6
    s = torch.from_numpy(rng.uniform(0, 2, (ns, 3)).astype(np.float32)).cuda()
    q = s[:nq].clone()
    idx = torch.full((nq, H), ns, dtype=torch.int64, device="cuda")  # every neighbour is the shadow point
    x = torch.randn(ns, cin, device="cuda")
    w = torch.randn(15, cin, cout, device="cuda")
    kp = torch.randn(15, 3, device="cuda") * 0.3
    assert float(ops.kpconv(q, s, idx, x, w, kp, 0.3).abs().max()) == 0.0
    idx2 = torch.from_numpy(oracle.batch_neighbors(q.cpu().numpy(), s.cpu().numpy(), [nq], [ns], 0.6).astype(np.int64)).cuda()
    y1 = ops.kpconv(q, s, idx2, x, w, kp, 0.3)
    y2 = ops.kpconv(q, s, idx2, 2.0 * x, w, kp, 0.3)
    assert float((y2 - 2.0 * y1).abs().max()) <= 2e-3 * float(y1.abs().max())  # linear in x up to TF32 rounding


# ----------------------------------------------------------------------------------------------------------- pyramid
PL_ARCH5 = ['simple', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb_strided', 'resnetb',
            'resnetb_strided', 'resnetb', 'nearest_upsample', 'unary', 'nearest_upsample', 'unary', 'nearest_upsample',
            'unary', 'nearest_upsample', 'unary']


@pytest.mark.parametrize("native", [True, False])
@pytest.mark.parametrize("fixture,dl0,arch", [("pyramid_ref.npz", 0.24, PL_ARCH5[:8] + PL_ARCH5[10:16]),
                                              ("pyramid_dales_ref.npz", 0.4, PL_ARCH5)])
def test_pyramid_matches_reference_segmentation_inputs(native, fixture, dl0, arch, torch_cuda):
    """Whole device pyramid (one native call, or driven per operator from Python) against the reference's own
    segmentation_inputs output (golden; random grid orientation included) for the Vaihingen3D walk (dl 0.24, 4 layers
    in the fixture) and the DALES walk (dl 0.4, 5 layers, train_DALES_PseudoLabel.py:98-121): points and lengths
    bit-exact; index matrices IDENTICAL after canonicalising groups of exactly equal d2 (nanoflann's std::sort orders on
    d2 alone), 100 % membership, the only other admissible difference being which member of a tie group the
    neighbourhood-limit crop kept (conftest.assert_same_up_to_ties)."""
    torch = torch_cuda
    from weasal_b200 import pyramid
    g = np.load(os.path.join(GOLDEN, fixture))

    class Cfg:
        first_subsampling_dl = dl0
        conv_radius = 2.5
        deform_radius = 6.0
        architecture = arch

    np.random.seed(int(g["seed"]))
    li = pyramid.segmentation_inputs(g["in_pts"], None, None, g["in_lens"], Cfg(), neighborhood_limits=list(g["limits"]),
                                     native=native)
    L = int(g["L"])
    assert (len(li) - 2) // 5 == L
    explained = 0
    for l in range(L):
        pts = li[l].cpu().numpy()
        assert np.array_equal(pts, g[f"points{l}"]), f"points layer {l}"
        assert np.array_equal(li[4 * L + l].cpu().numpy(), g[f"lengths{l}"])
        for off, nm in ((L, "neighbors"), (2 * L, "pools"), (3 * L, "upsamples")):
            got, ref = li[off + l].cpu().numpy(), g[f"{nm}{l}"].astype(np.int64)
            assert got.shape == ref.shape, f"{nm}{l} shape {got.shape} vs {ref.shape}"
            assert li[off + l].dtype == torch.int64
            if ref.size:
                q = li[l + 1] if nm == "pools" else li[l]
                s_ = li[l + 1] if nm == "upsamples" else li[l]
                n_perm, n_crop = assert_same_up_to_ties(q.cpu().numpy(), s_.cpu().numpy(), got, ref, f"{nm}{l}")
                explained += n_perm
    print(f"{fixture}: {explained} rows differ from the reference by tie permutations only")


def _vaihingen_cfg():
    from weasal_b200.net import CfgView, net_config
    return CfgView(net_config("vaihingen_pl"))


def test_native_pyramid_equals_operator_driver_and_prefetcher(torch_cuda):
    """kp_pyramid_build_dev (one call), the per-operator Python driver and the prefetch thread produce identical
    batches (same kernels underneath): every tensor bit-equal, at BASELINE size, random grid orientation on."""
    torch = torch_cuda
    from weasal_b200 import pyramid
    cfg = _vaihingen_cfg()
    b = make_batch("vaihingen_pl", seed=3)
    feats = np.ones((len(b["points"]), 4), np.float32)
    labels = np.zeros(len(b["points"]), np.int64)
    np.random.seed(11)
    ref = pyramid.segmentation_inputs(b["points"], feats, labels, b["lengths"], cfg, native=False)
    np.random.seed(11)
    nat = pyramid.segmentation_inputs(b["points"], feats, labels, b["lengths"], cfg, native=True)
    pf = pyramid.PyramidPrefetcher(cfg, "cuda")
    np.random.seed(11)
    pf.submit(torch.from_numpy(b["points"]).pin_memory(), torch.from_numpy(feats).pin_memory(),
              torch.from_numpy(labels).pin_memory(), b["lengths"])
    pf.submit(b["points"], feats, labels, b["lengths"])  # a second one in flight, different orientations
    got = pf.get()
    second = pf.get()
    pf.close()
    pre = got.points + got.neighbors + got.pools + got.upsamples + got.lengths + [got.features, got.labels]
    assert len(ref) == len(nat) == len(pre)
    for k, (a, c, d) in enumerate(zip(ref, nat, pre)):
        assert a.shape == c.shape == d.shape and a.dtype == c.dtype == d.dtype, f"entry {k}"
        assert torch.equal(a, c), f"native entry {k}"
        assert torch.equal(a, d), f"prefetched entry {k}"
    assert second.points[0].shape == got.points[0].shape
    assert not torch.equal(second.points[1][:100], got.points[1][:100])  # new orientations were drawn
    # conv matrices are marked as their own transpose; check the claim itself on one layer
    for l in range(len(got.neighbors)):
        assert getattr(nat[len(got.points) + l], "_kp_symmetric", False)
    nb = got.neighbors[2].cpu().numpy()
    n = nb.shape[0]
    src = np.repeat(np.arange(n), nb.shape[1])
    ok = nb.ravel() < n
    fwd = set(zip(src[ok].tolist(), nb.ravel()[ok].tolist()))
    assert all((j, i) in fwd for i, j in list(fwd)[:20000])


def test_prefetcher_on_an_sm_partition(torch_cuda):
    """Build streams confined to a green-context SM partition (kp_sm_partition_streams): same pyramid, bit for bit."""
    torch = torch_cuda
    from weasal_b200 import pyramid
    cfg = _vaihingen_cfg()
    b = make_batch("vaihingen_pl", seed=6, batch_num=2, in_radius=10.0)
    feats = np.ones((len(b["points"]), 1), np.float32)
    labels = (np.arange(len(b["points"])) % 5).astype(np.int64)
    outs = []
    for part in (0, 64):
        pf = pyramid.PyramidPrefetcher(cfg, "cuda", workers=2, sm_partition=part)
        assert (pf.sm_partition >= 64) if part else (pf.sm_partition == 0)
        np.random.seed(21)
        pf.submit(b["points"], feats, labels, b["lengths"])
        pf.submit(b["points"], feats, labels, b["lengths"])
        g0, g1 = pf.get(), pf.get()
        pf.close()
        outs.append([t.clone() for g in (g0, g1) for t in g.points + g.neighbors + g.pools + g.upsamples + [g.features, g.labels]])
    assert len(outs[0]) == len(outs[1])
    for a, c in zip(*outs):
        assert torch.equal(a, c)


def test_native_pyramid_grows_cap_and_slab(torch_cuda):
    """A first call with a neighbour capacity / slab that is too small reports what it needs and the wrapper repeats."""
    torch = torch_cuda
    from weasal_b200 import pyramid
    cfg = _vaihingen_cfg()
    b = make_batch("vaihingen_pl", seed=4, batch_num=2, in_radius=10.0)
    P = torch.from_numpy(b["points"]).cuda()
    np.random.seed(5)
    want = pyramid.build_native(P, b["lengths"], cfg)
    pyramid._SLAB_HINT.clear()
    np.random.seed(5)
    got = pyramid.build_native(P, b["lengths"], cfg, cap=7)
    for a, c in zip(sum(want[:5], []), sum(got[:5], [])):
        assert torch.equal(a, c)
    pyramid._SLAB_HINT[(5, 8, 80)] = 40.0  # far too small a slab
    np.random.seed(5)
    got = pyramid.build_native(P, b["lengths"], cfg)
    for a, c in zip(sum(want[:5], []), sum(got[:5], [])):
        assert torch.equal(a, c)


def test_kpconv_backward_symmetric_table_shortcut(torch_cuda):
    """Backward through a conv matrix marked `_kp_symmetric` (its own transpose) equals the general path that builds
    the transposed CSR table."""
    torch = torch_cuda
    from weasal_b200 import ops
    b = make_batch("vaihingen_pl", seed=6, batch_num=2, in_radius=8.0)
    P = torch.from_numpy(b["points"]).cuda()
    idx = ops.batch_query(P, P, b["lengths"], b["lengths"], 0.6)
    rng = np.random.default_rng(0)
    cin, cout = 16, 32
    w0 = torch.from_numpy((rng.normal(size=(15, cin, cout)) / 4).astype(np.float32)).cuda()
    kp = torch.from_numpy((rng.normal(size=(15, 3)) * 0.25).astype(np.float32)).cuda()
    x0 = torch.from_numpy(rng.normal(size=(len(b["points"]), cin)).astype(np.float32)).cuda()
    do = torch.from_numpy(rng.normal(size=(len(b["points"]), cout)).astype(np.float32)).cuda()
    res = []
    for sym in (False, True):
        ii = idx.clone()
        if sym:
            ii._kp_symmetric = True
        x, w = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
        ops.kpconv(P, P, ii, x, w, kp, 0.24).backward(do)
        res.append((x.grad.cpu().numpy(), w.grad.cpu().numpy()))
    # (the dX sums run in a different order and are rounded to TF32 before the contraction)
    assert rel_max(res[1][0], res[0][0]) < 3e-4 and rel_max(res[1][1], res[0][1]) < 1e-5


def test_full_size_vaihingen_batch_properties(torch_cuda):
    """BASELINE config sizes (4 spheres of radius 24 m, ~40k points): properties instead of the O(N^2) oracle."""
    torch = torch_cuda
    from weasal_b200 import ops
    b = make_batch("vaihingen_pl", seed=0)
    P = torch.from_numpy(b["points"]).cuda()
    L = b["lengths"]
    r = 0.6
    nb = ops.batch_query(P, P, L, L, r).cpu().numpy()
    Pn = b["points"]
    n = len(Pn)
    assert (nb[:, 0] == np.arange(n)).all()
    pad = np.vstack([Pn, np.full((1, 3), 1e9, np.float32)])
    d2 = ((pad[nb] - Pn[:, None, :]) ** 2).sum(2)
    real = nb < n
    assert (d2[real] < r * r * (1 + 1e-5)).all()
    assert (np.diff(np.where(real, d2, np.float32(1e30)), axis=1) >= -1e-9).all()  # rows sorted by distance, shadows last
    # symmetry of the neighbour relation for q == s (checked on a sample of pairs)
    src = np.repeat(np.arange(n), nb.shape[1])[real.ravel()]
    dst = nb[real]
    sel = np.random.default_rng(1).choice(len(src), 5000, replace=False)
    for i, j in zip(src[sel], dst[sel]):
        assert i in nb[j]
    offs = np.concatenate([[0], np.cumsum(L)])
    bid = np.searchsorted(offs, np.arange(n), side="right") - 1
    assert (bid[nb[real]] == np.repeat(bid, nb.shape[1])[real.ravel()]).all()  # never crosses batch elements
    # sampled rows against the oracle
    rows = np.random.default_rng(0).choice(n, 300, replace=False)
    for i in rows:
        lo, hi = offs[bid[i]], offs[bid[i] + 1]
        want = oracle.batch_neighbors(Pn[i:i + 1], Pn[lo:hi], [1], [hi - lo], r)[0] + lo
        k = len(want)
        assert np.array_equal(nb[i, :k], want) and (nb[i, k:] == n).all()


# ----------------------------------------------------------------------------------------------------------- pooling
def test_pooling_ops_match_reference_formulation(torch_cuda):
    """max_pool / closest_pool (models/blocks.py:77-112) forward and backward against the reference's cat + gather
    formulation evaluated by PyTorch."""
    torch = torch_cuda
    from weasal_b200 import ops
    rng = np.random.default_rng(0)
    ns, nq, C_, H = 500, 320, 48, 9
    x = torch.from_numpy(rng.normal(size=(ns, C_)).astype(np.float32)).cuda()
    idx_np = rng.integers(0, ns + 1, (nq, H))  # includes shadow entries (== ns)
    idx_np[5] = ns                              # one row made of shadows only
    idx = torch.from_numpy(idx_np).cuda()
    g = torch.from_numpy(rng.normal(size=(nq, C_)).astype(np.float32)).cuda()

    def ref_max(xx):
        xp = torch.cat((xx, torch.zeros_like(xx[:1])), 0)
        return xp[idx].max(dim=1)[0]

    def ref_closest(xx):
        xp = torch.cat((xx, torch.zeros_like(xx[:1])), 0)
        return xp[idx[:, 0]]

    for mine, ref in ((ops.max_pool, ref_max), (ops.closest_pool, ref_closest)):
        x1 = x.clone().requires_grad_(True)
        x2 = x.clone().requires_grad_(True)
        y1, y2 = mine(x1, idx), ref(x2)
        assert torch.equal(y1, y2)
        y1.backward(g)
        y2.backward(g)
        assert torch.allclose(x1.grad, x2.grad, rtol=0, atol=1e-5)
    y = ops.max_pool(x, idx.to(torch.int32))
    assert torch.equal(y, ref_max(x))
    # closest_pool backward reads a gradient that is a column slice of a wider matrix (torch.cat backward) in place
    wide = torch.from_numpy(rng.normal(size=(nq, C_ + 11)).astype(np.float32)).cuda()
    gs = wide[:, 5:5 + C_]
    assert not gs.is_contiguous()
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ops.closest_pool(xa, idx).backward(gs)
    ops.closest_pool(xb, idx).backward(gs.contiguous())
    assert torch.allclose(xa.grad, xb.grad, rtol=0, atol=1e-5)


def test_batch_query_more_than_256_neighbours_escalates(torch_cuda):
    """Rows beyond the fast kernel's 256-hit staging: the call repeats itself with the 1024-hit variant; beyond 1024
    the library reports KP_ERR_TOO_DENSE instead of returning truncated rows."""
    torch = torch_cuda
    from weasal_b200 import ops, radius_neighbors as rn
    rng = np.random.default_rng(4)
    s = rng.uniform(0, 1, (900, 3)).astype(np.float32)
    L = np.array([900], np.int32)
    got = rn.batch_query(s, s, L, L, radius=0.7)
    assert got.shape[1] > 256
    assert np.array_equal(got, oracle.batch_neighbors(s, s, L, L, 0.7))
    # the same through the deferred (no-sync) path used by the pyramid builder
    d = torch.from_numpy(s).cuda()
    pend = ops.PendingSearches(d.device)
    pend.add(d, d, L, L, 0.7, dtype=torch.int32)
    pend.add(d, d, L, L, 0.1, dtype=torch.int32)
    a, b = pend.resolve()
    assert np.array_equal(a.cpu().numpy(), got)
    assert np.array_equal(b.cpu().numpy(), oracle.batch_neighbors(s, s, L, L, 0.1))
    big = rng.uniform(0, 1, (3000, 3)).astype(np.float32)
    with pytest.raises(RuntimeError, match="1024 neighbours"):
        rn.batch_query(big, big, np.array([3000], np.int32), np.array([3000], np.int32), radius=0.9)


@pytest.mark.parametrize("K,radius,cin,cout", [(15, 0.8, 8, 16), (9, 0.45, 12, 20), (1, 0.45, 16, 8)])
def test_kpconv_long_rows_and_fewer_kernel_points(K, radius, cin, cout, torch_cuda):
    """Rows wider than the 128-neighbour shared-memory staging of the influence kernel (the two-pass path, also taken by
    the transposed table) and kernel sizes below 15."""
    torch = torch_cuda
    rng = np.random.default_rng(K)
    n = 400
    s = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    L = np.array([n], np.int32)
    idx = oracle.batch_neighbors(s, s, L, L, radius)
    if radius > 0.7:
        assert idx.shape[1] > 128
    a = dict(q_pts=s, s_pts=s, idx=idx, x=rng.normal(size=(n, cin)).astype(np.float32),
             weights=(rng.normal(size=(K, cin, cout)) / np.sqrt(cin)).astype(np.float32),
             kernel_points=(rng.normal(size=(K, 3)) * 0.3).astype(np.float32), extent=np.float32(0.3),
             d_out=rng.normal(size=(n, cout)).astype(np.float32))
    out, dx, dw = _run_kpconv(torch, a, torch.int64)
    o_out = oracle.kpconv_forward(s, s, idx, a["x"], a["weights"], a["kernel_points"], 0.3)
    o_dx, o_dw = oracle.kpconv_backward(s, s, idx, a["x"], a["weights"], a["kernel_points"], 0.3, a["d_out"])
    assert rel_max(out, o_out) < KP_TOL and rel_max(dx, o_dx) < KP_TOL and rel_max(dw, o_dw) < KP_TOL


def test_batch_query_no_supports_at_all(torch_cuda):
    from weasal_b200 import radius_neighbors as rn
    q = np.random.default_rng(0).uniform(0, 1, (10, 3)).astype(np.float32)
    with pytest.raises(RuntimeError, match="^Error$"):
        rn.batch_query(q, np.zeros((0, 3), np.float32), [10], [0], radius=0.5)



# ------------------------------------------------------------------------------------------ static shapes + CUDA graph
@pytest.mark.parametrize("prefetched", [False, True])
def test_graphed_static_step_matches_eager_dynamic_step(prefetched, torch_cuda):
    """The training step replayed from a CUDA graph over padded static-shape batches computes what the eager step
    computes over the ordinary batches: same losses, same parameters after three SGD steps (dropout off).
    ``prefetched``: the KPConv influence lists / transposed tables come from the prefetch stage (plan.ConvPlans) and the
    weight images from one packing launch per step (plan.WeightPacker), dW and dX forked onto two streams."""
    torch = torch_cuda
    import copy
    import torch.nn.functional as F
    from weasal_b200 import pyramid
    from weasal_b200.engine import GraphedTrainStep, calibrate_static_caps
    from weasal_b200.kpconv import KPConv
    from weasal_b200.net import CfgView, KPFCNNHarness, net_config
    ncfg = dict(net_config("vaihingen_pl"), dropout=0.0)
    view = CfgView(ncfg)
    data = [make_batch("vaihingen_pl", seed=s, batch_num=2, in_radius=9.0) for s in (1, 2, 3)]
    dev = torch.device("cuda")
    P = [torch.from_numpy(b["points"]).to(dev) for b in data]
    Fe = [torch.from_numpy(b["features"]).to(dev) for b in data]
    Lb = [torch.from_numpy(b["labels"] % 9).to(dev) for b in data]
    n_cap, limits = calibrate_static_caps(view, P, [b["lengths"] for b in data], random_grid_orient=False)
    np.random.seed(0)
    torch.manual_seed(0)
    net_a = KPFCNNHarness(ncfg, KPConv).to(dev)
    net_b = copy.deepcopy(net_a)
    losses = []
    for net, caps in ((net_a, None), (net_b, n_cap)):
        opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-3)
        plans = packer = None
        if caps and prefetched:
            from weasal_b200.engine import calibrate_conv_plans
            from weasal_b200.net import fused_linear_weights
            from weasal_b200.plan import WeightPacker
            plans = calibrate_conv_plans(net, view, P, [b["lengths"] for b in data], n_cap, limits, random_grid_orient=False)
            packer = WeightPacker(net, fused_linear_weights(net))
            assert len(plans.specs) == 10 and len(packer.jobs) >= 20
        tr = GraphedTrainStep(net, opt, F.cross_entropy, clip_value=100.0, plans=plans, packer=packer)
        pf = pyramid.PyramidPrefetcher(view, dev, neighborhood_limits=limits if caps else None, n_cap=caps,
                                       random_grid_orient=False, plans=plans)
        ls = []
        for i in range(3):
            pf.submit(P[i], Fe[i], Lb[i], data[i]["lengths"])
            batch = pf.get()
            assert (batch.static_slab is not None) == (caps is not None)
            assert (batch.plan_buf is not None) == (plans is not None)
            ls.append(float(tr.step(batch)))
        pf.close()
        assert (tr.n_graphed, tr.n_eager) == ((3, 0) if caps else (0, 3))
        if caps:
            assert tr.launches_per_replay > 50
        losses.append(ls)
    assert np.allclose(losses[0], losses[1], rtol=2e-3), losses
    for (na, pa), (nb_, pb) in zip(net_a.named_parameters(), net_b.named_parameters()):
        d = float((pa - pb).abs().max())
        assert d <= 2e-3 * max(float(pa.abs().max()), 1e-3), (na, d)


def test_two_graph_data_parallel_step_and_eager_fallback_in_between(torch_cuda):
    """The data-parallel form of the step (graph A: forward/backward + gradients gathered into a flat buffer; one collective;
    graph B: clip + optimizer reading the flat buffer) with a pass-through reducer computes what the single-graph step
    computes, also when an eager fall-back step (which re-creates every p.grad) runs between two replays — the stale-gradient
    hazard of a tail that reads p.grad."""
    torch = torch_cuda
    import copy
    import torch.nn.functional as F
    from weasal_b200 import pyramid
    from weasal_b200.engine import GraphedTrainStep, calibrate_static_caps
    from weasal_b200.kpconv import KPConv
    from weasal_b200.net import CfgView, KPFCNNHarness, net_config

    class PassThroughReducer:   # what GradAllReducer does on one rank
        calls = 0

        def step(self):
            pass

        def allreduce_flat(self, flat):
            PassThroughReducer.calls += 1

    ncfg = dict(net_config("vaihingen_pl"), dropout=0.0)
    view = CfgView(ncfg)
    data = [make_batch("vaihingen_pl", seed=s, batch_num=2, in_radius=9.0) for s in (1, 2, 3, 4)]
    dev = torch.device("cuda")
    P = [torch.from_numpy(b["points"]).to(dev) for b in data]
    Fe = [torch.from_numpy(b["features"]).to(dev) for b in data]
    Lb = [torch.from_numpy(b["labels"] % 9).to(dev) for b in data]
    n_cap, limits = calibrate_static_caps(view, P, [b["lengths"] for b in data], random_grid_orient=False)
    np.random.seed(0)
    torch.manual_seed(0)
    net_a = KPFCNNHarness(ncfg, KPConv).to(dev)
    net_b = copy.deepcopy(net_a)
    losses = []
    for net, reducer in ((net_a, None), (net_b, PassThroughReducer())):
        opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-3)
        tr = GraphedTrainStep(net, opt, F.cross_entropy, reducer=reducer, clip_value=100.0)
        pf = pyramid.PyramidPrefetcher(view, dev, neighborhood_limits=limits, n_cap=n_cap, random_grid_orient=False)
        ls = []
        for i in range(4):
            pf.submit(P[i], Fe[i], Lb[i], data[i]["lengths"])
            batch = pf.get()
            if i == 2:
                batch.no_crop = False   # force the eager step for this batch
            ls.append(float(tr.step(batch)))
        pf.close()
        assert (tr.n_graphed, tr.n_eager) == (3, 1)
        assert (tr.graph_tail is not None) == (reducer is not None)
        losses.append(ls)
    assert PassThroughReducer.calls == 3
    assert np.allclose(losses[0], losses[1], rtol=2e-3), losses
    for (na, pa), (nb_, pb) in zip(net_a.named_parameters(), net_b.named_parameters()):
        d = float((pa.detach() - pb.detach()).abs().max())
        assert d <= 2e-3 * max(float(pa.detach().abs().max()), 1e-3), (na, d)


def test_static_pyramid_layout_and_overflow_fallback(torch_cuda):
    """Static layout: real rows equal the ordinary pyramid, padded rows are all-shadow / 1e6 / ignore_index; a batch
    that outgrows a capacity comes back in the ordinary layout."""
    torch = torch_cuda
    from weasal_b200 import pyramid
    from weasal_b200.engine import calibrate_static_caps
    cfg = _vaihingen_cfg()
    b = make_batch("vaihingen_pl", seed=8, batch_num=2, in_radius=9.0)
    dev = torch.device("cuda")
    P, Fe, Lb = (torch.from_numpy(b[k]).to(dev) for k in ("points", "features", "labels"))
    n_cap, limits = calibrate_static_caps(cfg, [P], [b["lengths"]], random_grid_orient=False)
    want = pyramid.build_native(P, b["lengths"], cfg, random_grid_orient=False)
    pf = pyramid.PyramidPrefetcher(cfg, dev, neighborhood_limits=limits, n_cap=n_cap, random_grid_orient=False)
    pf.submit(P, Fe, Lb, b["lengths"])
    got = pf.get()
    assert got.static_slab is not None and got.no_crop
    assert got.static_slab.numel() >= int(got.build.need_bytes[0]) > 0  # the host's layout size covers the builder's
    for l in range(len(want[0])):
        n = want[0][l].shape[0]
        assert got.points[l].shape[0] == n_cap[l]
        assert torch.equal(got.points[l][:n], want[0][l]) and bool((got.points[l][n:] == 1e6).all())
        for k, (mine, ref) in enumerate(((got.neighbors[l], want[1][l]), (got.pools[l], want[2][l]), (got.upsamples[l], want[3][l]))):
            if ref.shape[0] == 0:
                assert mine.shape[0] == 0
                continue
            rows, w = ref.shape
            ns_ref = want[0][l + 1].shape[0] if k == 2 else n          # shadow value of the ordinary matrix
            ns_cap = n_cap[l + 1] if k == 2 else n_cap[l]               # ... and of the static one
            m = mine[:rows, :w]
            assert torch.equal(torch.where(ref == ns_ref, torch.full_like(ref, ns_cap), ref), m)
            assert bool((mine[:rows, w:] == ns_cap).all()) and bool((mine[rows:] == ns_cap).all())
    # max_pool on the fixed-width matrix with its true width == max_pool on the ordinary matrix, bit for bit; without
    # the width, rows as wide as the matrix get the extra zero candidate of a shadow column
    from weasal_b200 import ops
    x0 = torch.randn(len(P), 16, device=dev) - 0.5
    xpad = torch.cat([x0, torch.zeros(n_cap[0] - len(P), 16, device=dev)])
    m_dyn = ops.max_pool(x0, want[2][0])
    assert int(got.pool_widths[0]) == want[2][0].shape[1]
    assert torch.equal(m_dyn, ops.max_pool(xpad, got.pools[0], got.pool_widths[0])[:m_dyn.shape[0]])
    full = (want[2][0] < len(P)).all(1)
    m_wide = ops.max_pool(xpad, got.pools[0])[:m_dyn.shape[0]]
    assert torch.equal(torch.clamp(m_dyn[full], min=0), m_wide[full]) and torch.equal(m_dyn[~full], m_wide[~full])
    assert torch.equal(got.features[:len(P)], Fe) and bool((got.features[len(P):] == 0).all())
    assert torch.equal(got.labels[:len(P)], Lb) and bool((got.labels[len(P):] == -100).all())
    pf.close()
    small = [max(c // 2, 256) for c in n_cap]
    pf = pyramid.PyramidPrefetcher(cfg, dev, neighborhood_limits=limits, n_cap=small, random_grid_orient=False)
    pf.submit(P, Fe, Lb, b["lengths"])
    got = pf.get()
    pf.close()
    assert got.static_slab is None and got.points[0].shape[0] == len(P)
    assert torch.equal(got.neighbors[0], want[1][0][:, :got.neighbors[0].shape[1]])


# ------------------------------------------------------------------------------------------------------ unary blocks
@pytest.mark.parametrize("n,cin,cout,bias,slope", [(5000, 32, 16, False, 0.1), (3000, 64, 9, True, 1.0),
                                                   (2000, 1536, 512, False, 0.1), (300, 512, 1024, False, 1.0),
                                                   (777, 6, 10, True, 0.1), (40000, 16, 64, False, 1.0),
                                                   (129, 64, 64, True, 0.1)])
def test_linear_act_matches_fp32_reference(n, cin, cout, bias, slope, torch_cuda):
    """Unary block kernel (Linear + bias + LeakyReLU, forward / dX / dW / db) against the plain PyTorch fp32 expression
    (models/blocks.py:467-507 on 2-D features); TF32 operands, 1e-3 relative like the KPConv contraction."""
    torch = torch_cuda
    import torch.nn.functional as F
    from weasal_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(n + cin)
    x0 = torch.randn(n, cin, device="cuda", generator=g)
    w0 = torch.randn(cout, cin, device="cuda", generator=g) / np.sqrt(cin)
    b0 = torch.randn(cout, device="cuda", generator=g) if bias else None
    dy = torch.randn(n, cout, device="cuda", generator=g)
    x, w = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True)
    b = b0.clone().requires_grad_(True) if bias else None
    y = ops.linear_act(x, w, b, slope)
    y.backward(dy)
    pre = F.linear(x0, w0, b0)
    y_ref = F.leaky_relu(pre, slope) if slope != 1.0 else pre
    # the LeakyReLU derivative is taken where OUR output is positive: a TF32 forward and an fp32 forward disagree on the
    # sign of the few outputs within rounding of zero, and each such flip would show as a 0.9 * dy difference
    g = dy * torch.where(y.detach() > 0, 1.0, slope) if slope != 1.0 else dy
    ref = [y_ref, g @ w0, g.t() @ x0] + ([g.sum(0)] if bias else [])
    got = [y.detach(), x.grad, w.grad] + ([b.grad] if bias else [])
    near_zero = (pre.abs() < 2e-3 * pre.abs().max())
    assert float(((y.detach() > 0) != (pre > 0))[~near_zero].float().sum()) == 0  # signs agree away from zero
    for name, a, r in zip(("y", "dx", "dw", "db"), got, ref):
        a, r = a.cpu().numpy(), r.cpu().numpy()
        assert a.shape == r.shape
        assert rel_max(a, r) < KP_TOL and rel_l2(a, r) < KP_TOL, (name, rel_max(a, r), rel_l2(a, r))


# ------------------------------------------------------------------------------------------------- sphere voting
def test_sharded_sphere_voting_equals_single_rank(torch_cuda):
    """Votes accumulated by two 'ranks' taking alternate sphere batches and summed equal the single-rank votes
    (sum / count form: order independent), every cloud point of the tile interior is voted on, and only points within
    0.7 * in_radius of a centre receive votes."""
    torch = torch_cuda
    from weasal_b200.kpconv import KPConv
    from weasal_b200.net import CfgView, KPFCNNHarness, net_config
    from weasal_b200.voting import vote_cloud
    ncfg = net_config("vaihingen_pl")
    R = 6.0
    tile, inten, _ = make_als_tile(3, 6 * R, 6.0)
    dev = torch.device("cuda")
    cloud = torch.from_numpy(tile).to(dev)
    feats = torch.from_numpy(np.stack([np.ones(len(tile), np.float32), inten, tile[:, 2]], 1)).to(dev)
    np.random.seed(0)
    torch.manual_seed(0)
    net = KPFCNNHarness(ncfg, KPConv).to(dev)
    view = CfgView(ncfg)
    p1, v1, s1, n1 = vote_cloud(net, view, cloud, feats, R, 2, num_votes=1, random_grid_orient=False)
    stats = dict(vote_cloud.last_stats)
    assert stats["graphed"] > 0 and stats["graphed"] >= 4 * stats["eager"], stats  # the CUDA-graph path is the one that ran
    # ... and it equals the eager path (same kernels; the padded layout changes nothing for the real rows)
    p0, v0, s0, n0 = vote_cloud(net, view, cloud, feats, R, 2, num_votes=1, random_grid_orient=False, graph=False)
    assert (s0, n0) == (s1, n1) and torch.equal(v0, v1)
    assert float((p0 - p1).abs().max()) < 1e-3
    sums, votes, spheres = 0, 0, 0
    for r in range(2):
        p, v, s, n = vote_cloud(net, view, cloud, feats, R, 2, num_votes=1, rank=r, world_size=2,
                                random_grid_orient=False)
        sums = sums + p * v.unsqueeze(1)
        votes = votes + v
        spheres += s
    assert spheres == s1 and torch.equal(votes, v1)
    assert float((v1 > 0).float().mean()) > 0.8  # (the kept ball is 3-D, tester:188-191; tall vegetation at R = 6 m falls outside)
    got = sums / votes.clamp_min(1e-12).unsqueeze(1)
    # same spheres, same network, same pyramids: equal up to the order of the float additions
    assert float((got - p1).abs().max()) < 1e-3
    assert torch.allclose(got.sum(1)[votes > 0], torch.ones_like(got.sum(1)[votes > 0]), atol=1e-4)
