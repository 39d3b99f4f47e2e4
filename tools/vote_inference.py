"""Sphere-voting inference over a whole synthetic tile, spheres sharded across the ranks (BASELINE.json configs[3]).

    python tools/vote_inference.py [--config vaihingen_pl|dales_pl] [--votes 2] [--extent 8]      (1 GPU)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/vote_inference.py ...

Prints one JSON line on rank 0: voted points/s (points pushed through the network by all ranks / max-over-ranks time),
coverage (share of cloud points that received a vote) and the agreement of a sharded run with itself.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from weasal_b200 import grid_subsampling  # noqa: E402
from weasal_b200.kpconv import KPConv  # noqa: E402
from weasal_b200.net import CfgView, KPFCNNHarness, net_config  # noqa: E402
from weasal_b200.synthetic import CONFIGS, make_als_tile  # noqa: E402
from weasal_b200.voting import vote_cloud  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="vaihingen_pl", choices=["vaihingen_pl", "dales_pl"])
    ap.add_argument("--votes", type=int, default=2)
    ap.add_argument("--extent", type=float, default=8.0, help="tile edge in units of in_radius")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--workers", type=int, default=3, help="pyramid builds in flight (prefetch threads / side streams)")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cuda.matmul.allow_tf32 = True
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = CONFIGS[args.config]
    tile, inten, labels = make_als_tile(args.seed, args.extent * cfg["in_radius"], cfg["density"])
    f0 = np.stack([inten, tile[:, 2]], 1)
    sub_p, sub_f, _ = grid_subsampling.subsample(tile, features=f0, classes=labels, sampleDl=cfg["dl"])
    cloud = torch.from_numpy(sub_p).to(dev)
    ones = np.ones((len(sub_p), 1), np.float32)
    feats = np.hstack([ones, sub_f]) if cfg["in_features"] == 4 else np.hstack([ones, sub_f[:, 1:2]])
    feats = torch.from_numpy(feats.astype(np.float32)).to(dev)  # vote_cloud appends z_rel per sphere
    ncfg = net_config(args.config)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    net = KPFCNNHarness(ncfg, KPConv).to(dev)
    view = CfgView(ncfg)
    # warm-up pass (allocator pools, scratch arenas; it also calibrates the static batch layout and captures the forward
    # graph, like the reference's one-off sampler calibration), then the timed pass, which reuses that state
    vote_cloud(net, view, cloud, feats, cfg["in_radius"], cfg["batch_num"], 1, rank=rank, world_size=world, seed=1,
               workers=args.workers)
    state = vote_cloud.last_state
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    probs, votes, n_sph, n_pts = vote_cloud(net, view, cloud, feats, cfg["in_radius"], cfg["batch_num"], args.votes,
                                            rank=rank, world_size=world, seed=args.seed, state=state, workers=args.workers)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), float(n_pts), float(n_sph)], dtype=torch.float64, device=dev)
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({
            "metric": "voted points/sec", "value": float(t[1]) / (float(tmax[0]) / 1e3), "unit": "points/s",
            "n_gpus": world, "ms_total": float(tmax[0]), "spheres": int(t[2]), "points_through_network": int(t[1]),
            "cloud_points": int(cloud.shape[0]), "coverage": float((votes > 0).float().mean()),
            "mean_votes_per_point": float(votes.mean()), "classes_predicted": int(probs.argmax(1).unique().numel()),
            "batches_rank0": vote_cloud.last_stats,
            "config": {"workload": f"{args.config}: sphere voting over a {args.extent:g} x {args.extent:g} in_radius tile, "
                                   f"{args.votes} passes, spheres sharded over {world} rank(s)", "data": "synthetic"}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
