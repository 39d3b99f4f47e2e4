"""Subprocess body of tests/test_gpu_dropin.py (its own process: the import harness rebinds `datasets`, `models`, ...).

Runs the UNMODIFIED reference network (models.architectures.KPFCNN built by the reference's block_decider, the
reference's Config subclass from its train_*.py, the reference's <DS>CustomBatch) on one CUDA batch
  (1) with the reference's stock aten-chain KPConv / max_pool / closest_pool, and
  (2) after weasal_b200.dropin.install() swapped KPConv and the pooling gathers for the sm_100a kernels,
same parameters (state_dict copied), same dropout mask (same torch seed), and prints one JSON line with the relative
differences of the logits, the loss and every parameter gradient. The batch itself is produced twice as well:
  (a) by the reference's own ``segmentation_inputs`` (datasets/common.py:461-577) calling the product's numpy drop-ins
      for the two extension modules, collated by the unchanged CustomBatch and moved with ``batch.to('cuda')``;
  (b) by the device pyramid (kp_pyramid_build_dev) handed to the same unchanged CustomBatch through
      ``dropin.collate_device``;
and (a) and (b) must agree tensor for tensor.

usage: ref_dropin_script.py <vaihingen_pl|dales_pl> <in_radius> <batch_num>
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    cfg_name, in_radius, batch_num = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
    import torch
    from oracle import ref_harness
    from weasal_b200.synthetic import make_batch

    root = ref_harness.find_root()
    assert root is not None, "no copy of the reference on this machine (tools/install_reference.py)"
    ref_harness.install(root, backend="weasal_b200")
    script, cls, ds_mod, batch_cls = {
        "vaihingen_pl": ("train_Vaihingen3D_PseudoLabel", "Vaihingen3DPLConfig", "datasets.Vaihingen3D_PseudoLabel",
                         "Vaihingen3DPLCustomBatch"),
        "dales_pl": ("train_DALES_PseudoLabel", "DALESPLConfig", "datasets.DALES_PseudoLabel", "DALESPLCustomBatch"),
    }[cfg_name]
    cfg = getattr(importlib.import_module(script), cls)()
    Batch = getattr(importlib.import_module(ds_mod), batch_cls)
    from datasets.common import PointCloudDataset
    import models.blocks as B
    from models.architectures import KPFCNN
    from weasal_b200 import dropin, kpconv, pyramid

    dev = torch.device("cuda", 0)
    b = make_batch(cfg_name, seed=5, batch_num=batch_num, in_radius=in_radius)
    nb = len(b["lengths"])
    n_cls = int(b["cfg"]["num_classes"])
    cfg.num_classes = n_cls
    cfg.class_w = [1.0] * n_cls
    labels = (b["labels"] % n_cls).astype(np.int64)
    extras = [np.ones((nb, 3), np.float32), np.tile(np.eye(3, dtype=np.float32), (nb, 1, 1)), np.zeros(nb, np.int32),
              np.zeros(nb, np.int32), np.arange(len(b["points"]), dtype=np.int64)]

    # (a) the reference's own pyramid walk on the product's numpy drop-ins
    ds = PointCloudDataset("x")
    ds.config = cfg
    ds.neighborhood_limits = []
    np.random.seed(77)
    li = ds.segmentation_inputs(b["points"], b["features"], labels, b["lengths"])
    batch_a = Batch([li + extras]).to(dev)
    # (b) the device pyramid through the unchanged CustomBatch
    np.random.seed(77)
    dl = pyramid.segmentation_inputs(b["points"], b["features"], labels, b["lengths"], cfg, native=True)
    batch_b = dropin.collate_device(Batch, dl + [torch.from_numpy(e).to(dev) for e in extras])
    same = True
    n_tensors = 0
    for name in ("points", "neighbors", "pools", "upsamples", "lengths"):
        for ta, tb in zip(getattr(batch_a, name), getattr(batch_b, name)):
            n_tensors += 1
            same = same and ta.shape == tb.shape and ta.dtype == tb.dtype and bool(torch.equal(ta, tb))
    same = same and bool(torch.equal(batch_a.features, batch_b.features)) and bool(torch.equal(batch_a.labels, batch_b.labels))
    L = len(batch_b.points)

    def run(net, batch, slope=None):
        """forward + loss + backward with a fixed dropout mask; ``slope`` overrides every LeakyReLU's negative_slope
        (an attribute of the reference's own modules: with slope 1 the network has no activation kinks left)"""
        for m in net.modules():
            if isinstance(m, torch.nn.LeakyReLU):
                m.negative_slope = 0.1 if slope is None else slope
        net.zero_grad(set_to_none=True)
        torch.manual_seed(11)
        out = net(batch, cfg)
        loss = net.loss(out, batch.labels)
        loss.backward()
        return out.detach().clone(), float(loss), {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}

    # (1) stock reference operator
    np.random.seed(3)
    torch.manual_seed(3)
    net_ref = KPFCNN(cfg, list(range(n_cls)), []).to(dev)
    net_ref.train()
    stock_kpconv = B.KPConv
    out_ref, loss_ref, g_ref = run(net_ref, batch_a)
    out_ref_lin, _, g_ref_lin = run(net_ref, batch_a, slope=1.0)
    # sensitivity of the STOCK network to a perturbation of the size of one TF32 rounding (2^-11 relative) of its input
    # features: how far parameter gradients move when a LeakyReLU / max-pool decision flips somewhere
    feats0 = batch_a.features.clone()
    torch.manual_seed(99)
    batch_a.features = feats0 * (1 + 2.0 ** -11 * torch.randn_like(feats0))
    out_pert, _, g_pert = run(net_ref, batch_a)
    batch_a.features = feats0

    # (2) the drop-in
    assert dropin.install() is True and B.KPConv is kpconv.KPConv and B.KPConv is not stock_kpconv
    np.random.seed(3)
    torch.manual_seed(3)
    net_new = KPFCNN(cfg, list(range(n_cls)), []).to(dev)
    n_conv = sum(isinstance(m, kpconv.KPConv) for m in net_new.modules())
    net_new.load_state_dict(net_ref.state_dict(), strict=True)
    net_new.train()
    out_new, loss_new, g_new = run(net_new, batch_b)
    out_new_lin, _, g_new_lin = run(net_new, batch_b, slope=1.0)
    torch.cuda.synchronize()

    def rel(a, r):
        return float((a - r).abs().max() / r.abs().max().clamp_min(1e-30))

    def grad_rel(ga, gr):
        assert set(ga) == set(gr)   # the identity BatchNorm layers have no grad on either side
        per = {n: rel(ga[n], gr[n]) for n in gr}
        worst = max(per, key=per.get)
        va, vr = torch.cat([ga[n].flatten() for n in gr]), torch.cat([gr[n].flatten() for n in gr])
        return per[worst], worst, float((va - vr).norm() / vr.norm())

    gmax, gworst, gl2 = grad_rel(g_new, g_ref)
    lmax, lworst, ll2 = grad_rel(g_new_lin, g_ref_lin)
    pmax, pworst, pl2 = grad_rel(g_pert, g_ref)
    print(json.dumps({
        "config": cfg_name, "points": int(len(b["points"])), "layers": L, "pyramid_tensors_compared": n_tensors,
        "pyramid_paths_identical": bool(same), "kpconv_modules_swapped": int(n_conv),
        "state_dict_keys": len(net_ref.state_dict()), "logits_rel": rel(out_new, out_ref),
        "loss_ref": loss_ref, "loss_new": loss_new, "n_grads": len(g_ref),
        "grad_rel_max": gmax, "grad_rel_worst": gworst, "grad_rel_l2": gl2,
        "linear_logits_rel": rel(out_new_lin, out_ref_lin), "linear_grad_rel_max": lmax, "linear_grad_rel_worst": lworst,
        "linear_grad_rel_l2": ll2,
        "stock_perturbed_logits_rel": rel(out_pert, out_ref), "stock_perturbed_grad_rel_max": pmax,
        "stock_perturbed_grad_rel_l2": pl2}))


if __name__ == "__main__":
    main()
