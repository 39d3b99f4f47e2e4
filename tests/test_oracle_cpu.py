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


PL_ARCH5 = ['simple', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb_strided', 'resnetb',
            'resnetb_strided', 'resnetb', 'nearest_upsample', 'unary', 'nearest_upsample', 'unary', 'nearest_upsample',
            'unary', 'nearest_upsample', 'unary']


@pytest.mark.parametrize("fixture,dl0,arch", [("pyramid_ref.npz", 0.24, PL_ARCH5[:8] + PL_ARCH5[10:16]),
                                              ("pyramid_dales_ref.npz", 0.4, PL_ARCH5)])
def test_oracle_pyramid_equals_reference_up_to_exact_ties(fixture, dl0, arch):
    """The restated pyramid walk on the restated cores ((d2, index) order) against the reference's own
    segmentation_inputs output, Vaihingen3D and DALES walks: points bit-exact, index matrices identical once groups of
    exactly equal d2 are canonicalised; differences are admitted only inside the tie group a crop column cut."""
    from conftest import assert_same_up_to_ties
    from oracle.pyramid_ref import segmentation_inputs_cpu
    g = np.load(os.path.join(GOLDEN, fixture))

    class Cfg:
        first_subsampling_dl = dl0
        conv_radius = 2.5
        deform_radius = 6.0
        architecture = arch

    np.random.seed(int(g["seed"]))
    li = segmentation_inputs_cpu(g["in_pts"], None, None, g["in_lens"], Cfg(), neighborhood_limits=list(g["limits"]),
                                 use_ref=False)
    L = int(g["L"])
    assert (len(li) - 2) // 5 == L
    for l in range(L):
        assert np.array_equal(li[l], g[f"points{l}"])
        for off, nm in ((L, "neighbors"), (2 * L, "pools"), (3 * L, "upsamples")):
            ref = g[f"{nm}{l}"].astype(np.int64)
            if ref.size:
                q = li[l + 1] if nm == "pools" else li[l]
                s = li[l + 1] if nm == "upsamples" else li[l]
                assert_same_up_to_ties(q, s, li[off + l], ref, f"{nm}{l}")


def test_tie_canonicalisation_detects_real_mismatches():
    from conftest import assert_same_up_to_ties
    rng = np.random.default_rng(0)
    base = rng.uniform(0, 2, (60, 3)).astype(np.float32)
    s = np.concatenate([base, base], 0)  # every point twice: exact ties everywhere
    L = np.array([len(s)], np.int32)
    mine = oracle.batch_neighbors(s, s, L, L, 0.6).astype(np.int64)
    swapped = mine.copy()
    swapped[:, [0, 1]] = swapped[:, [1, 0]]  # self and its duplicate: d2 = 0 both, order is arbitrary in the reference
    n_perm, n_crop = assert_same_up_to_ties(s, s, mine, swapped, "dup")
    assert n_perm == len(s) and n_crop == 0
    wrong = mine.copy()
    i = int(np.argmax((mine < len(s)).sum(1)))
    wrong[i, 2] = (wrong[i, 2] + 1) % len(s) if (wrong[i, 2] + 1) % len(s) not in wrong[i] else wrong[i, 2]
    far = int(np.setdiff1d(np.arange(len(s)), mine[i])[0])
    wrong[i, 2] = far  # a support that is NOT within the radius
    with pytest.raises(AssertionError):
        assert_same_up_to_ties(s, s, mine, wrong, "wrong")


def test_oracle_kpconv_matches_reference_wide_golden():
    """One of the wide cases (256 -> 32 on 17k points) through the f64 oracle against the reference's sampled fp32 outputs."""
    g = np.load(os.path.join(GOLDEN, "kpconv_wide_ref.npz"))
    name, cin, cout = "w256_32", 256, 32
    pts, idx, rows = g["pts"], g["idx"].astype(np.int64), g["rows"].astype(np.int64)
    rng = np.random.default_rng(int(g[f"{name}.seed"]))
    x = rng.standard_normal((len(pts), cin), dtype=np.float32)
    w = (rng.standard_normal((15, cin, cout), dtype=np.float32) / np.float32(np.sqrt(cin * 4.0))).astype(np.float32)
    kp = g[f"{name}.kernel_points"]
    out = oracle.kpconv_forward(pts[rows], pts, idx[rows], x, w, kp, float(g["extent"]))
    assert rel_err(out, g[f"{name}.out_rows"]) < 3e-6
