"""Test-time sphere voting over a whole cloud, spheres sharded across ranks — the hot path's inference caller
(utils/tester_PseudoLabel.py:87-195 ``cloud_segmentation_test``; BASELINE.json configs[3]).

Per batch of spheres the reference does: extract all cloud points within ``in_radius`` of each centre and re-centre
them (datasets/Vaihingen3D_PseudoLabel.py:318-365), build the pyramid, run the network, softmax, keep the points
within ``0.7 * in_radius`` of the centre (tester:188-191) and blend their probabilities into ``test_probs`` (tester:194).
Here every step stays on the device: sphere extraction is three kernels over the cloud (weasal_b200/spheres.py), the
pyramid comes from the prefetch thread (one native call per batch), and votes are accumulated by a kernel as (sum of
probabilities, number of votes) —
the order-independent form of the reference's visit-order dependent EMA (weasal_b200/distributed.py) — so that ranks
can take disjoint sets of spheres and meet in one all-reduce at the end.
"""
import numpy as np
import torch

from .distributed import shard_indices
from .spheres import VoteBuffer
from .pyramid import PyramidPrefetcher


def vote_centres(points_xy_min, points_xy_max, in_radius, num_votes, seed=0):
    """Deterministic visiting schedule standing in for the potential-based picker: ``num_votes`` jittered passes over
    a grid whose pitch (0.7 * in_radius * sqrt(2)) lets the kept discs of one pass cover the tile. float32 [S, 2]."""
    rng = np.random.default_rng(seed)
    pitch = 0.7 * in_radius * np.sqrt(2.0) * 0.98
    lo, hi = np.asarray(points_xy_min, np.float64), np.asarray(points_xy_max, np.float64)
    nx, ny = (np.maximum(np.ceil((hi - lo) / pitch), 1)).astype(int)
    out = []
    for v in range(num_votes):
        jitter = rng.uniform(-0.5, 0.5, 2) * pitch if v else np.zeros(2)
        gx = lo[0] + (np.arange(nx) + 0.5) * (hi[0] - lo[0]) / nx + jitter[0]
        gy = lo[1] + (np.arange(ny) + 0.5) * (hi[1] - lo[1]) / ny + jitter[1]
        out.append(np.stack(np.meshgrid(gx, gy, indexing="ij"), -1).reshape(-1, 2))
    return np.concatenate(out, 0).astype(np.float32)


def extract_spheres_device(cloud, feats, centres_xy, in_radius, keep_frac=0.7):
    """All points within ``in_radius`` of each centre (the centre's z is that of its nearest cloud point in xy), stacked
    and re-centred, by the sphere-extraction kernels (spheres.extract_spheres). Returns (points [N,3], features [N,C],
    lengths int32 [B] (host), cloud indices [N]) or None when every sphere is empty; empty spheres are dropped."""
    from .spheres import extract_spheres
    d_xy = ((cloud[:, None, :2] - centres_xy[None, :, :]) ** 2).sum(-1)          # [N, B]
    c3 = cloud[torch.argmin(d_xy, 0)].double().cpu().numpy()                     # [B, 3]
    p, lens, inds = extract_spheres(cloud, c3, in_radius)
    if lens.sum() == 0:
        return None
    f = torch.cat([feats[inds], p[:, 2:3]], 1)
    return p, f, lens[lens > 0], inds


@torch.no_grad()
def vote_cloud(net, config, cloud, feats, in_radius, batch_num, num_votes=1, num_classes=None, rank=0, world_size=1,
               seed=0, group=None, neighborhood_limits=None, random_grid_orient=True):
    """Votes of this rank's share of the spheres, all-reduced: returns (probabilities [N, classes], votes [N],
    spheres this rank ran, points this rank pushed through the network). ``cloud`` [N,3] / ``feats`` [N,C-1] are CUDA
    tensors holding the whole (subsampled) cloud on every rank; the network sees ``[feats, z_rel]`` per point."""
    dev = cloud.device
    lo, hi = cloud[:, :2].min(0)[0].cpu().numpy(), cloud[:, :2].max(0)[0].cpu().numpy()
    centres = vote_centres(lo, hi, in_radius, num_votes, seed)
    batches = [centres[i:i + batch_num] for i in range(0, len(centres), batch_num)]
    mine = [batches[i] for i in shard_indices(len(batches), rank, world_size)]
    centres_dev = [torch.from_numpy(b).to(dev) for b in mine]
    net.eval()
    acc = None
    pf = PyramidPrefetcher(config, dev, neighborhood_limits=neighborhood_limits, random_grid_orient=random_grid_orient)
    pending = []
    n_spheres = n_points = 0

    def submit(k):
        ex = extract_spheres_device(cloud, feats, centres_dev[k], in_radius)
        if ex is None:
            return False
        p, f, lens, inds = ex
        pf.submit(p, f, None, lens, extras=dict(input_inds=inds, in_points=p, in_lengths=lens))
        return True

    todo = iter(range(len(mine)))

    def submit_next():
        for k in todo:
            if submit(k):
                return True
        return False

    inflight = submit_next()
    while inflight:
        batch = pf.get()
        inflight = submit_next()  # the next batch's extraction + pyramid overlap this batch's forward pass
        probs = torch.softmax(net(batch), 1)
        if acc is None:
            acc = VoteBuffer(cloud.shape[0], probs.shape[1] if num_classes is None else num_classes, dev, mode="sum")
        # (tester_PseudoLabel.py:188-194: only the points within 0.7 * in_radius of the sphere centre vote)
        acc.update(probs, batch.in_points, batch.input_inds, batch.in_lengths, radius_limit=0.7 * in_radius)
        n_spheres += len(batch.lengths[0])
        n_points += batch.points[0].shape[0]
    pf.close()
    if acc is None:
        acc = VoteBuffer(cloud.shape[0], num_classes or 1, dev, mode="sum")
    acc.reduce(group)  # all-reduces (sum of probabilities, votes) in place
    probs, _, _ = acc.reproject()
    return probs, acc.weight, n_spheres, n_points
