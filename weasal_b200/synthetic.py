"""Synthetic ALS (airborne laser scanning) clouds shaped like the reference's datasets.

There is no network for Vaihingen3D / DALES, so tests and bench.py draw clouds from this generator
(SURVEY.md §8d). Shapes follow the reference's configs:
  * Vaihingen3D PseudoLabel: first_subsampling_dl 0.24, in_radius 24, batch_num 4, 4 input features
    ``[1, intensity, z_abs, z_rel]`` (train_Vaihingen3D_PseudoLabel.py:98-121, datasets/Vaihingen3D_PseudoLabel.py:423-430);
  * DALES PseudoLabel: first_subsampling_dl 0.4, in_radius 18, batch_num 4, 3 input features ``[1, z_abs, z_rel]``
    (train_DALES_PseudoLabel.py:98-121).
Sphere extraction mirrors ``potential_item`` (datasets/Vaihingen3D_PseudoLabel.py:318-365): all cloud points within
``in_radius`` of a centre, re-centred on it.
"""
import numpy as np

CONFIGS = {
    # name: (first_subsampling_dl, in_radius, batch_num, in_features_dim, first_features_dim, density pts/m^2, classes)
    "vaihingen_pl": dict(dl=0.24, in_radius=24.0, batch_num=4, in_features=4, first_features_dim=64,
                         density=6.0, num_classes=9, conv_radius=2.5, num_layers=5),
    "dales_pl": dict(dl=0.4, in_radius=18.0, batch_num=4, in_features=3, first_features_dim=128,
                     density=14.0, num_classes=8, conv_radius=2.5, num_layers=5),
    "vaihingen_wl": dict(dl=0.24, in_radius=18.0, batch_num=3, in_features=4, first_features_dim=64,
                         density=6.0, num_classes=9, conv_radius=2.5, num_layers=3),
}


def make_als_tile(seed, extent_m, density_pts_m2):
    """Terrain + roof plateaus + vegetation. Returns points f32 [N,3], intensity f32 [N], labels i32 [N]."""
    rng = np.random.default_rng(seed)
    n = int(extent_m * extent_m * density_pts_m2)
    x = rng.uniform(0.0, extent_m, n)
    y = rng.uniform(0.0, extent_m, n)
    z = 2.0 * np.sin(x / 7.0) + 1.5 * np.cos(y / 5.0)
    labels = np.zeros(n, np.int32)
    cx = np.floor(x / 12.0).astype(np.int64)
    cy = np.floor(y / 12.0).astype(np.int64)
    roof = ((cx + cy) % 3) == 0
    z = z + np.where(roof, 6.0, 0.0)
    labels[roof] = 1
    veg = rng.uniform(0.0, 1.0, n) < 0.25
    z = z + np.where(veg, rng.uniform(0.0, 8.0, n), 0.0)
    labels[veg] = 2 + (rng.integers(0, 3, n)[veg])
    pts = np.stack([x, y, z], axis=1)
    pts = (pts - pts[0]).astype(np.float32)  # mirrors coord_offset, Vaihingen3D_PseudoLabel.py:661-693
    intensity = rng.uniform(0.0, 1.0, n).astype(np.float32)
    return pts, intensity, labels


def pick_centres(points, num, in_radius, seed):
    """Sphere centres on a coarse grid of the tile interior (stand-in for the potential-based picker)."""
    rng = np.random.default_rng(seed)
    lo = points[:, :2].min(0) + in_radius
    hi = points[:, :2].max(0) - in_radius
    if np.any(hi <= lo):
        lo, hi = points[:, :2].min(0), points[:, :2].max(0)
    c = np.empty((num, 3), np.float32)
    for i in range(num):
        xy = rng.uniform(lo, hi)
        d = np.sum((points[:, :2] - xy) ** 2, axis=1)
        c[i] = points[int(np.argmin(d))]
    return c


def extract_spheres(points, centres, in_radius):
    """Returns (stacked centred points f32 [N,3], lengths i32 [B], input_inds i64 [N])."""
    pts, lens, inds = [], [], []
    r2 = np.float32(in_radius) ** 2
    for c in centres:
        d2 = np.sum((points - c) ** 2, axis=1)
        sel = np.nonzero(d2 < r2)[0]
        pts.append((points[sel] - c).astype(np.float32))
        lens.append(len(sel))
        inds.append(sel)
    return np.concatenate(pts, 0), np.asarray(lens, np.int32), np.concatenate(inds, 0).astype(np.int64)


def make_batch(config="vaihingen_pl", seed=0, batch_num=None, in_radius=None, density=None):
    """One stacked batch of input spheres in the reference's collate layout.

    Returns dict(points [N,3] f32, lengths [B] i32, features [N,Cin0] f32, labels [N] i64, cfg).
    """
    cfg = dict(CONFIGS[config])
    if batch_num is not None:
        cfg["batch_num"] = batch_num
    if in_radius is not None:
        cfg["in_radius"] = in_radius
    if density is not None:
        cfg["density"] = density
    R = cfg["in_radius"]
    extent = max(4.0 * R, 2.5 * R + 20.0)
    tile, inten, labels = make_als_tile(seed, extent, cfg["density"])
    centres = pick_centres(tile, cfg["batch_num"], R, seed + 1)
    pts, lens, inds = extract_spheres(tile, centres, R)
    z_abs = tile[inds, 2:3]
    z_rel = pts[:, 2:3]
    ones = np.ones((len(pts), 1), np.float32)
    if cfg["in_features"] == 4:
        feats = np.hstack([ones, inten[inds, None], z_abs, z_rel]).astype(np.float32)
    elif cfg["in_features"] == 3:
        feats = np.hstack([ones, z_abs, z_rel]).astype(np.float32)
    else:
        feats = ones
    return dict(points=pts, lengths=lens, features=feats, labels=labels[inds].astype(np.int64), cfg=cfg,
                input_inds=inds)
