"""One profiled training step for ncu (use with --profile-from-start off): warm-up steps run unprofiled, then
cudaProfilerStart / one step / cudaProfilerStop."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from weasal_b200 import grid_subsampling, pyramid  # noqa: E402
from weasal_b200.kpconv import KPConv  # noqa: E402
from weasal_b200.net import CfgView, KPFCNNHarness, net_config  # noqa: E402

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "vaihingen_pl"
dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
cfg, batches = bench.build_batches(cfg_name, 0, 2,
                                   lambda p, f, l, dl: grid_subsampling.subsample(p, features=f, classes=l, sampleDl=dl))
ncfg = net_config(cfg_name)
view = CfgView(ncfg)
np.random.seed(0)
torch.manual_seed(0)
net = KPFCNNHarness(ncfg, KPConv).to(dev).train()
opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.98)
dbs = [{k: torch.from_numpy(v).to(dev) for k, v in b.items() if k != "lengths"} for b in batches]


def step(i):
    b, d = batches[i % 2], dbs[i % 2]
    li = pyramid.segmentation_inputs(d["points"], d["features"], d["labels"], b["lengths"], view, device=dev)
    batch = pyramid.DeviceBatch(li)
    loss = F.cross_entropy(net(batch), batch.labels)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    torch.nn.utils.clip_grad_value_(net.parameters(), 100.0)
    opt.step()
    return loss


for i in range(3):
    step(i)
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = step(3)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss.detach()))
