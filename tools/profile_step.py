"""Where does a bench step spend its time? Wall-clock per phase (with syncs) + torch.profiler top ops."""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from weasal_b200 import grid_subsampling, pyramid  # noqa: E402
from weasal_b200.kpconv import KPConv  # noqa: E402
from weasal_b200.net import CfgView, KPFCNNHarness, net_config  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = True
cfg, batches = bench.build_batches("vaihingen_pl", 0, 4,
                                   lambda p, f, l, dl: grid_subsampling.subsample(p, features=f, classes=l, sampleDl=dl))
ncfg = net_config("vaihingen_pl")
view = CfgView(ncfg)
np.random.seed(0)
torch.manual_seed(0)
net = KPFCNNHarness(ncfg, KPConv).to(dev).train()
opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.98)
dbs = [{k: torch.from_numpy(v).to(dev) for k, v in b.items() if k != "lengths"} for b in batches]


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


def step(i, verbose=False):
    b, d = batches[i % 4], dbs[i % 4]
    t0 = sync()
    li = pyramid.segmentation_inputs(d["points"], d["features"], d["labels"], b["lengths"], view, device=dev)
    t1 = sync()
    batch = pyramid.DeviceBatch(li)
    loss = F.cross_entropy(net(batch), batch.labels)
    t2 = sync()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    t3 = sync()
    torch.nn.utils.clip_grad_value_(net.parameters(), 100.0)
    opt.step()
    t4 = sync()
    if verbose:
        print(f"pyramid {1e3*(t1-t0):.2f} ms | fwd {1e3*(t2-t1):.2f} | bwd {1e3*(t3-t2):.2f} | opt {1e3*(t4-t3):.2f}", flush=True)


for i in range(6):
    step(i, verbose=True)
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(0)
    step(1)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=15, max_name_column_width=60))
