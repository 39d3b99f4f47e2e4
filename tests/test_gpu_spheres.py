"""GPU tests of the stages either side of the network: sphere extraction, augmentation + feature
assembly, vote update / reprojection / confusion — each against the reference's own formulation in numpy / sklearn."""
import numpy as np
import pytest

from weasal_b200.synthetic import make_als_tile

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def test_extract_spheres_matches_sklearn_query_radius(torch_cuda):
    """Membership = sklearn's KDTree.query_radius (what potential_item calls, Vaihingen3D_PseudoLabel.py:345-348), points
    re-centred like :365; rows in ascending cloud index."""
    torch = torch_cuda
    from sklearn.neighbors import KDTree
    from weasal_b200.spheres import extract_spheres
    pts, _, _ = make_als_tile(3, 80.0, 8.0)
    tree = KDTree(pts, leaf_size=10)
    rng = np.random.default_rng(0)
    centres = pts[rng.choice(len(pts), 5, replace=False)].astype(np.float64) + rng.normal(0, 1.0, (5, 3))
    centres[4] += 1000.0  # an empty sphere
    got_p, lens, got_i = extract_spheres(torch.from_numpy(pts).cuda(), centres, 12.0, cap=1000)  # cap too small: grows
    got_p, got_i = got_p.cpu().numpy(), got_i.cpu().numpy()
    i0 = 0
    for b, c in enumerate(centres):
        want = np.sort(tree.query_radius(c.reshape(1, -1), r=12.0)[0])
        assert lens[b] == len(want)
        assert np.array_equal(got_i[i0:i0 + lens[b]], want)
        assert np.array_equal(got_p[i0:i0 + lens[b]], (pts[want].astype(np.float64) - c.reshape(1, -1)).astype(np.float32))
        i0 += lens[b]
    assert lens[4] == 0 and i0 == len(got_i)


def test_augmentation_and_features_bit_exact_vs_numpy(torch_cuda):
    """augmented = np.sum(np.expand_dims(points, 2) * R, axis=1) * scale + noise (datasets/common.py:318) and the feature
    stack [1, colours * keep, z + centre_z, z] (Vaihingen3D_PseudoLabel.py:383, 423-430), bit for bit, with the draws made in
    the reference's np.random order."""
    torch = torch_cuda
    from weasal_b200.spheres import augment, draw_augmentation

    class Cfg:
        augment_rotation = 'vertical'
        augment_scale_anisotropic = True
        augment_symmetries = [True, True, True]
        augment_scale_min, augment_scale_max, augment_noise = 0.2, 1.8, 0.06

    rng = np.random.default_rng(1)
    lens = np.array([700, 0, 1300], np.int32)
    n = int(lens.sum())
    p = rng.normal(0, 5, (n, 3)).astype(np.float32)
    colors = rng.uniform(0, 1, (5000, 1)).astype(np.float32)
    inds = rng.choice(5000, n).astype(np.int64)
    centre_z = np.array([3.5, 0.0, -1.25], np.float32)
    keep = np.array([1.0, 1.0, 0.0], np.float32)
    np.random.seed(5)
    R, scale, noise = draw_augmentation(lens, Cfg())
    # the same draws, straight from the reference's lines (datasets/common.py:262-304)
    np.random.seed(5)
    want_p, i0 = [], 0
    for b, m in enumerate(lens):
        theta = np.random.rand() * 2 * np.pi
        c, s = np.cos(theta), np.sin(theta)
        Rb = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float32)
        sc = np.random.rand(3) * (1.8 - 0.2) + 0.2
        sym = np.array([True, True, True]).astype(np.int32) * np.random.randint(2, size=3)
        sc = (sc * (1 - sym * 2)).astype(np.float32)
        nz = (np.random.randn(m, 3) * 0.06).astype(np.float32)
        assert np.array_equal(Rb, R[b]) and np.array_equal(sc, scale[b]) and np.array_equal(nz, noise[i0:i0 + m])
        want_p.append(np.sum(np.expand_dims(p[i0:i0 + m], 2) * Rb, axis=1) * sc + nz)
        i0 += m
    want_p = np.concatenate(want_p, 0)
    got_p, got_f = augment(torch.from_numpy(p).cuda(), lens, R, scale, noise, colors=torch.from_numpy(colors).cuda(),
                           input_inds=torch.from_numpy(inds).cuda(), centre_z=centre_z, color_keep=keep, fdim=4)
    assert np.array_equal(got_p.cpu().numpy(), want_p)
    b_of = np.repeat(np.arange(3), lens)
    want_f = np.hstack([np.ones((n, 1), np.float32), colors[inds] * keep[b_of, None], want_p[:, 2:] + centre_z[b_of, None],
                        want_p[:, 2:]]).astype(np.float32)
    assert np.array_equal(got_f.cpu().numpy(), want_f)


def test_vote_update_reprojection_and_confusion(torch_cuda):
    """test_probs update (utils/tester_PseudoLabel.py:176-195, spheres in order, overlapping), the order-independent
    accumulation, reprojection (:270-283) and fast_confusion (utils/metrics.py:35-118) against numpy."""
    torch = torch_cuda
    from weasal_b200.spheres import VoteBuffer
    rng = np.random.default_rng(2)
    N, Cc, in_r = 5000, 9, 10.0
    lens = np.array([900, 1100, 800], np.int32)
    n = int(lens.sum())
    inds = np.concatenate([rng.choice(N, m, replace=False) for m in lens]).astype(np.int64)  # spheres overlap in the cloud
    pts = rng.normal(0, 5, (n, 3)).astype(np.float32)
    probs = rng.dirichlet(np.ones(Cc), n).astype(np.float32)
    want = np.zeros((N, Cc), np.float32)
    wsum, wcnt = np.zeros((N, Cc), np.float32), np.zeros(N, np.float32)
    i0 = 0
    for m in lens:
        mask = np.sum(pts[i0:i0 + m] ** 2, axis=1) < (0.7 * in_r) ** 2
        ii, pp = inds[i0:i0 + m][mask], probs[i0:i0 + m][mask]
        want[ii] = np.float32(0.95) * want[ii] + np.float32(1 - 0.95) * pp
        wsum[ii] += pp
        wcnt[ii] += 1
        i0 += m
    dev = torch.device("cuda")
    t = lambda a: torch.from_numpy(a).to(dev)
    ema = VoteBuffer(N, Cc, dev, mode="ema", smooth=0.95)
    ema.update(t(probs), t(pts), t(inds), lens, radius_limit=0.7 * in_r)
    assert np.allclose(ema.probs.cpu().numpy(), want, rtol=0, atol=1e-7)
    acc = VoteBuffer(N, Cc, dev, mode="sum")
    acc.update(t(probs), t(pts), t(inds), lens, radius_limit=0.7 * in_r)
    assert np.array_equal(acc.weight.cpu().numpy(), wcnt)
    assert np.allclose(acc.probs.cpu().numpy(), wsum, rtol=0, atol=1e-6)
    # reprojection on evaluation points + confusion
    proj = rng.choice(N, 20000).astype(np.int64)
    truth = rng.integers(0, Cc, 20000).astype(np.int32)
    out, pred, conf = ema.reproject(t(proj), t(truth))
    assert np.array_equal(out.cpu().numpy(), ema.probs.cpu().numpy()[proj])
    want_pred = np.argmax(want[proj], axis=1).astype(np.int32)
    assert np.array_equal(pred.cpu().numpy(), np.argmax(ema.probs.cpu().numpy()[proj], axis=1).astype(np.int32))
    vec = np.bincount(truth.astype(np.int64) * Cc + pred.cpu().numpy(), minlength=Cc * Cc).reshape(Cc, Cc)  # fast_confusion
    assert np.array_equal(conf.cpu().numpy(), vec)
    assert (want_pred == pred.cpu().numpy()).mean() > 0.999
    out2, pred2, _ = acc.reproject()
    ref2 = wsum / np.maximum(wcnt, 1e-12)[:, None]
    assert np.allclose(out2.cpu().numpy(), ref2, rtol=1e-6, atol=1e-7)
