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
               seed=0, group=None, neighborhood_limits=None, random_grid_orient=True, graph=True, calib_batches=8,
               workers=3, state=None):
    """Votes of this rank's share of the spheres, all-reduced: returns (probabilities [N, classes], votes [N],
    spheres this rank ran, points this rank pushed through the network). ``cloud`` [N,3] / ``feats`` [N,C-1] are CUDA
    tensors holding the whole (subsampled) cloud on every rank; the network sees ``[feats, z_rel]`` per point.

    ``graph`` (default, needs no ``neighborhood_limits``): row capacities, matrix widths and list capacities are
    calibrated on ``calib_batches`` batches spread over this rank's schedule (like the reference's sampler calibration),
    the batches come in the static layout with their forward influence lists prepared by the prefetch workers, and the
    forward pass + softmax is ONE CUDA graph replay per batch (engine.GraphedForward); a batch that does not fit runs
    eagerly. Sphere extraction runs on its own stream, so its host read-backs (sphere sizes) never wait for the network.
    ``state``: the calibration + captured graph of an earlier call on the same network and cloud density
    (``vote_cloud.last_state``), reused instead of calibrating again (the reference calibrates once per dataset and then
    runs many voting passes, tester_PseudoLabel.py:149-160)."""
    dev = cloud.device
    lo, hi = cloud[:, :2].min(0)[0].cpu().numpy(), cloud[:, :2].max(0)[0].cpu().numpy()
    centres = vote_centres(lo, hi, in_radius, num_votes, seed)
    batches = [centres[i:i + batch_num] for i in range(0, len(centres), batch_num)]
    mine = [batches[i] for i in shard_indices(len(batches), rank, world_size)]
    centres_dev = [torch.from_numpy(b).to(dev) for b in mine]
    net.eval()
    acc = None
    softmax = lambda y: torch.softmax(y, 1)
    fwd = None
    graph = graph and neighborhood_limits is None and len(mine) > 0
    if graph:
        from .engine import GraphedForward, calibrate_conv_plans, calibrate_static_caps
        from .net import fused_linear_weights
        from .plan import WeightPacker
        if state is None:
            picks = sorted(set(int(round(v)) for v in np.linspace(0, len(mine) - 1, min(calib_batches, len(mine)))))
            cal = [extract_spheres_device(cloud, feats, centres_dev[k], in_radius) for k in picks]
            cal = [c for c in cal if c is not None]
            if cal:
                n_cap, limits = calibrate_static_caps(config, [c[0] for c in cal], [c[2] for c in cal], row_margin=1.25,
                                                      width_margin=0.25, random_grid_orient=random_grid_orient)
                plans = calibrate_conv_plans(net, config, [c[0] for c in cal], [c[2] for c in cal], n_cap, limits,
                                             random_grid_orient=random_grid_orient, margin=1.5, forward_only=True)
                state = dict(n_cap=n_cap, limits=limits, plans=plans, net=net,
                             fwd=GraphedForward(net, post=softmax, plans=plans,
                                                packer=WeightPacker(net, fused_linear_weights(net))))
            del cal
        if state is not None:
            assert state["net"] is net, "vote_cloud: state belongs to another network"
            fwd = state["fwd"]
            pf = PyramidPrefetcher(config, dev, neighborhood_limits=state["limits"], random_grid_orient=random_grid_orient,
                                   n_cap=state["n_cap"], plans=state["plans"], workers=workers)
    vote_cloud.last_state = state
    if fwd is None:
        workers = 1
        pf = PyramidPrefetcher(config, dev, neighborhood_limits=neighborhood_limits, random_grid_orient=random_grid_orient)
    n_spheres = n_points = 0
    ex_stream = torch.cuda.Stream(dev)
    ex_stream.wait_stream(torch.cuda.current_stream(dev))

    def submit(k):
        with torch.cuda.stream(ex_stream):
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

    inflight = sum(1 for _ in range(workers) if submit_next())
    while inflight:
        batch = pf.get()
        inflight -= 1
        if submit_next():  # the next batches' extraction + pyramid overlap this batch's forward pass
            inflight += 1
        cur = torch.cuda.current_stream(dev)
        for t in (batch.in_points, batch.input_inds):  # allocated on the extraction stream, read on this one
            t.record_stream(cur)
        probs = fwd.run(batch) if fwd is not None else softmax(net(batch))
        if acc is None:
            acc = VoteBuffer(cloud.shape[0], probs.shape[1] if num_classes is None else num_classes, dev, mode="sum")
        # (tester_PseudoLabel.py:188-194: only the points within 0.7 * in_radius of the sphere centre vote)
        acc.update(probs, batch.in_points, batch.input_inds, batch.in_lengths, radius_limit=0.7 * in_radius)
        n_spheres += len(batch.in_lengths)
        n_points += int(batch.in_points.shape[0])
    pf.close()
    g0, e0 = getattr(vote_cloud, "_seen", {}).get(id(fwd), (0, 0)) if fwd is not None else (0, 0)
    vote_cloud.last_stats = dict(graphed=fwd.n_graphed - g0 if fwd is not None else 0,
                                 eager=fwd.n_eager - e0 if fwd is not None else -(-n_spheres // max(batch_num, 1)))
    if fwd is not None:
        vote_cloud._seen = {id(fwd): (fwd.n_graphed, fwd.n_eager)}
    if acc is None:
        acc = VoteBuffer(cloud.shape[0], num_classes or 1, dev, mode="sum")
    acc.reduce(group)  # all-reduces (sum of probabilities, votes) in place
    probs, _, _ = acc.reproject()
    return probs, acc.weight, n_spheres, n_points
