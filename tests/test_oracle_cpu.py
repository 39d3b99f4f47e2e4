"""CPU tests: the oracle (oracle/kp_oracle.c) against the golden vectors produced from the unmodified reference
(tests/golden/make_golden.py) and, when oracle/_ref is present, against the compiled reference itself."""
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, load_case
from weasal_b200.synthetic import make_als_tile, make_batch

KP_CASES = ["c4_32", "c16_16", "c64_64", "c32_128", "c3_64", "strided16"]


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def kp_golden():
    return np.load(os.path.join(GOLDEN, "kpconv_ref.npz"))


@pytest.fixture(scope="module")
def pre_golden():
    return np.load(os.path.join(GOLDEN, "precompute_ref.npz"))


@pytest.mark.parametrize("name", KP_CASES)
def test_oracle_kpconv_matches_reference_kpconv(kp_golden, name):
    a = load_case(kp_golden, name)
    out = oracle.kpconv_forward(a["q_pts"], a["s_pts"], a["idx"], a["x"], a["weights"], a["kernel_points"],
                                float(a["extent"]))
    dx, dw = oracle.kpconv_backward(a["q_pts"], a["s_pts"], a["idx"], a["x"], a["weights"], a["kernel_points"],
                                    float(a["extent"]), a["d_out"])
    # the reference is fp32 PyTorch, the oracle accumulates in f64: agreement to fp32 round-off
    assert rel_err(out, a["out"]) < 2e-6
    assert rel_err(dx, a["dx"]) < 2e-6
    assert rel_err(dw, a["dw"]) < 2e-6


def test_oracle_neighbors_match_reference(pre_golden):
    g = pre_golden
    pts, lens = g["pts"], g["lens"]
    mine = oracle.batch_neighbors(pts, pts, lens, lens, 0.6)
    assert np.array_equal(mine, g["nbr_r0.6_ordered"])  # bit-exact vs batch_ordered_neighbors
    # wired-in nanoflann path: same shape, same membership per row (order may differ inside exact-d2 ties only)
    nf = g["nbr_r0.6_nanoflann"]
    assert nf.shape == mine.shape
    assert np.array_equal(np.sort(nf, 1), np.sort(mine, 1))
    sp, sl = g["sub0.48_pts"], g["sub0.48_lens"]
    pool = oracle.batch_neighbors(sp, pts, sl, lens, 0.6)
    assert np.array_equal(np.sort(pool, 1), np.sort(g["pool_r0.6_nanoflann"], 1))
    up = oracle.batch_neighbors(pts, sp, lens, sl, 1.2)
    assert np.array_equal(np.sort(up, 1), np.sort(g["up_r1.2_nanoflann"], 1))
    # rows that differ from nanoflann must differ only by permuting equal-distance neighbours
    for mat, ref, q, s in ((mine, nf, pts, pts), (pool, g["pool_r0.6_nanoflann"], sp, pts)):
        bad = np.nonzero((mat != ref).any(1))[0]
        for i in bad:
            sel = mat[i] < len(s)
            d_a = ((q[i] - s[mat[i][sel]]) ** 2).sum(1)
            d_b = ((q[i] - s[ref[i][sel]]) ** 2).sum(1)
            assert np.allclose(d_a, d_b, rtol=0, atol=0)


def test_oracle_subsampling_matches_reference(pre_golden):
    g = pre_golden
    sp, sl = oracle.grid_subsample_batch(g["pts"], g["lens"], sampleDl=0.48)
    assert np.array_equal(sl, g["sub0.48_lens"])
    assert np.array_equal(sp, g["sub0.48_pts"])  # bit-exact, including the unordered_map output order
    p2, f2, c2 = oracle.grid_subsample(g["pts"], features=g["feats"], classes=g["labels"], sampleDl=0.9)
    assert np.array_equal(p2, g["sub0.9_pts"])
    assert np.array_equal(f2, g["sub0.9_feats"])
    assert np.array_equal(c2, g["sub0.9_classes"])


@pytest.mark.parametrize("n", [0, 1, 2, 13, 14, 15, 29, 30, 200, 5000, 70000])
def test_unordered_map_order_model_vs_live_container(n):
    rng = np.random.default_rng(n)
    keys = rng.choice(1 << 40, size=n, replace=False).astype(np.uint64)
    assert np.array_equal(oracle.umap_order(keys), oracle.umap_live_order(keys))
    keys = (np.arange(n, dtype=np.uint64) * np.uint64(13))  # bucket collisions at 13 buckets
    assert np.array_equal(oracle.umap_order(keys), oracle.umap_live_order(keys))


def test_oracle_first_occurrence_order_is_a_permutation_of_reference_order():
    pts, _, _ = make_als_tile(3, 30.0, 10.0)
    a = oracle.grid_subsample(pts, sampleDl=0.7, order="reference")
    b = oracle.grid_subsample(pts, sampleDl=0.7, order="first")
    assert a.shape == b.shape
    assert np.array_equal(a[np.lexsort(a.T)], b[np.lexsort(b.T)])


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed,dl", [(0, 0.24), (1, 0.4), (2, 1.3), (3, 2.4)])
def test_oracle_vs_compiled_reference_random(seed, dl):
    pts, inten, lab = make_als_tile(seed, 40.0, 12.0)
    feats = np.stack([inten, pts[:, 2]], 1)
    rp, rf, rc = oracle.ref_subsample(pts, features=feats, classes=lab, sampleDl=dl)
    op, of, oc = oracle.grid_subsample(pts, features=feats, classes=lab, sampleDl=dl)
    assert np.array_equal(rp, op) and np.array_equal(rf, of) and np.array_equal(rc, oc)
    b = make_batch("vaihingen_pl", seed=seed, batch_num=3, in_radius=5.0)
    P, L = b["points"], b["lengths"]
    ro = oracle.ref_batch_neighbors(P, P, L, L, 2.5 * 0.24, ordered=True)
    assert np.array_equal(ro, oracle.batch_neighbors(P, P, L, L, 2.5 * 0.24))
    rs, rl = oracle.ref_subsample_batch(P, L, sampleDl=dl, max_p=50)
    os_, ol = oracle.grid_subsample_batch(P, L, sampleDl=dl, max_p=50)
    assert np.array_equal(rl, ol) and np.array_equal(rs, os_)


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_label_ties_follow_unordered_map_order():
    # many labels per voxel, including negative ones and more than 13 distinct values (forces a histogram rehash)
    rng = np.random.default_rng(5)
    pts = rng.uniform(0, 4, (4000, 3)).astype(np.float32)
    lab = rng.integers(-3, 40, 4000).astype(np.int32)
    rp, rc = oracle.ref_subsample(pts, classes=lab, sampleDl=1.0)
    op, oc = oracle.grid_subsample(pts, classes=lab, sampleDl=1.0)
    assert np.array_equal(rp, op) and np.array_equal(rc, oc)


def test_rotation_restatement():
    rng = np.random.default_rng(0)
    p = rng.normal(size=(500, 3)).astype(np.float32)
    R = np.linalg.qr(rng.normal(size=(3, 3)))[0].astype(np.float32)
    ref = np.sum(np.expand_dims(p, 2) * R, axis=1)  # the expression of datasets/common.py:118
    assert np.array_equal(oracle.rotate(p, R), ref)
    ref_t = np.sum(np.expand_dims(p, 2) * R.T, axis=1)  # common.py:134
    assert np.array_equal(oracle.rotate(p, R, transpose=True), ref_t)
