"""Unary-block kernels (kp_linear_*_dev) against the library path (cuBLAS TF32 GEMM + separate LeakyReLU kernels) on the
shapes of the Vaihingen3D-PL harness network: forward and forward+backward times, CUDA events, L2 flushed.

    python tools/bench_linear.py > gpurun_out/linear_shapes.txt
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from weasal_b200 import ops  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = True
SHAPES = [(39936, 32, 16), (39936, 16, 64), (39936, 32, 64), (39936, 64, 16), (31488, 64, 32), (31488, 32, 128),
          (31488, 64, 128), (31488, 128, 32), (18432, 128, 64), (18432, 64, 256), (18432, 128, 256), (18432, 256, 64),
          (8192, 256, 128), (8192, 128, 512), (8192, 256, 512), (8192, 512, 128), (2048, 512, 256), (2048, 256, 1024),
          (2048, 512, 1024), (8192, 1536, 512), (18432, 768, 256), (31488, 384, 128), (39936, 192, 64), (39936, 64, 64),
          (39936, 64, 9)]
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")


def timeit(fn, reps=20):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


print(f"{'n':>6} {'cin':>5} {'cout':>5} | {'ours fwd':>9} {'lib fwd':>9} | {'ours f+b':>9} {'lib f+b':>9}  (us, median of 20)")
for n, cin, cout in SHAPES:
    x = torch.randn(n, cin, device="cuda", requires_grad=True)
    w = (torch.randn(cout, cin, device="cuda") / cin ** 0.5).requires_grad_(True)
    dy = torch.randn(n, cout, device="cuda")

    def ours_f():
        with torch.no_grad():
            return ops.linear_act(x, w, None, 0.1)

    def lib_f():
        with torch.no_grad():
            return F.leaky_relu(F.linear(x, w), 0.1)

    def ours_fb():
        x.grad = w.grad = None
        ops.linear_act(x, w, None, 0.1).backward(dy)

    def lib_fb():
        x.grad = w.grad = None
        F.leaky_relu(F.linear(x, w), 0.1).backward(dy)

    for f in (ours_f, lib_f, ours_fb, lib_fb):
        f()
    torch.cuda.synchronize()
    print(f"{n:6d} {cin:5d} {cout:5d} | {timeit(ours_f):9.1f} {timeit(lib_f):9.1f} | {timeit(ours_fb):9.1f} {timeit(lib_fb):9.1f}")
