"""bench.py — train points/sec of the KPConv hot path on synthetic Vaihingen3D-shaped ALS spheres.

A "step" is one pass of the hot path over one batch: device pyramid (batch radius search x13 + grid subsampling x4,
datasets/common.py:461-577) -> KPFCNN-shaped network forward (10 KPConv) -> loss -> backward -> gradient clip ->
SGD step (+ one gradient all-reduce when N > 1). Workload = BASELINE.json configs[1] (Vaihingen3D PseudoLabel:
in_radius 24 m, first_subsampling_dl 0.24, batch_num 4, 4 input features, first_features_dim 64, 5 layers).

  python bench.py --gpus N --steps K --warmup W            our arm (hand-written sm_100a kernels)
  python bench.py --impl reference --steps K --warmup W    the reference's CPU implementation of the same path
                                                           (oracle/_ref C++ cores + PyTorch CPU operator chain)
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train points/sec"
UNIT = "points/s"
N_BATCHES = 8           # distinct pre-extracted sphere batches cycled through the steps
L2_FLUSH_BYTES = 256 << 20


# ------------------------------------------------------------------------------------------------------------ data
def build_batches(cfg_name, seed, n_batches, subsample_fn):
    """Synthetic tile -> whole-cloud grid subsampling at first_subsampling_dl (the dataset-prep call site,
    datasets/Vaihingen3D_PseudoLabel.py:795-805) -> n_batches stacks of batch_num spheres."""
    from weasal_b200.synthetic import CONFIGS, extract_spheres, make_als_tile, pick_centres
    cfg = dict(CONFIGS[cfg_name])
    R = cfg["in_radius"]
    tile, inten, labels = make_als_tile(seed, 4.0 * R, cfg["density"])
    feats = np.stack([inten, tile[:, 2]], 1)
    sub_p, sub_f, sub_l = subsample_fn(tile, feats, labels, cfg["dl"])
    sub_l = sub_l.reshape(-1)
    batches = []
    for b in range(n_batches):
        centres = pick_centres(sub_p, cfg["batch_num"], R, seed * 1000 + b)
        pts, lens, inds = extract_spheres(sub_p, centres, R)
        ones = np.ones((len(pts), 1), np.float32)
        if cfg["in_features"] == 4:
            f = np.hstack([ones, sub_f[inds, 0:1], sub_f[inds, 1:2], pts[:, 2:3]]).astype(np.float32)
        else:
            f = np.hstack([ones, sub_f[inds, 1:2], pts[:, 2:3]]).astype(np.float32)
        batches.append(dict(points=pts, lengths=lens, features=f, labels=sub_l[inds].astype(np.int64)))
    return cfg, batches


# ---------------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons read through NVML from the benchmark thread itself, a few times inside the timed
    region while kernels are in flight. (A background `nvidia-smi -lms` loop slowed the launch-bound step down by up
    to 2x and a polling NVML thread by 10-30 %: both contend with kernel launches for driver locks / the GIL; a
    handful of in-line reads cost well under 1 %.)"""
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, gpu_index):
        self.rows, self.h, self.max_mhz = [], None, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def sample(self):
        if self.h is None:
            return
        nv = self.nv
        try:
            sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            try:
                why = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                why = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.rows.append((float(sm), int(why)))
        except Exception:
            pass

    def summary(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        sm = [r[0] for r in self.rows]
        reasons = sorted({n for r in self.rows for n, b in self.BITS.items() if r[1] & b})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------- algorithmic
def layer_rows(batch):
    """real (unpadded) points per layer"""
    if getattr(batch, "build", None) is not None:
        return [int(n) for n in batch.build.n_out]
    return [p.shape[0] for p in batch.points]


def conv_shapes(batch, net):
    """(Nq, Ns, H, Cin, Cout) of every KPConv call in one forward, from the built pyramid (real rows, stored widths)."""
    out = []
    rows = layer_rows(batch)
    for blk in net.encoder:
        l = blk.layer
        if blk.strided:
            nq, ns, H = rows[l + 1], rows[l], batch.pools[l].shape[1]
        else:
            nq, ns, H = rows[l], rows[l], batch.neighbors[l].shape[1]
        out.append((nq, ns, H, blk.conv.in_channels, blk.conv.out_channels))
    return out


def algorithmic(shapes, K=15, idx_bytes=8):
    """compulsory bytes / contraction flops of the KPConv calls of one step."""
    fwd_b = sum(idx_bytes * nq * H + 12 * (nq + ns) + 4 * ns * ci + 4 * nq * co + 4 * K * ci * co + 12 * K
                for nq, ns, H, ci, co in shapes)
    fwd_f = sum(2 * nq * K * (H * ci + ci * co) for nq, ns, H, ci, co in shapes)
    dx_b = sum(idx_bytes * nq * H + 12 * (nq + ns) + 4 * nq * co + 4 * ns * ci + 4 * K * ci * co
               for nq, ns, H, ci, co in shapes)
    dw_b = sum(idx_bytes * nq * H + 12 * (nq + ns) + 4 * ns * ci + 4 * nq * co + 4 * K * ci * co
               for nq, ns, H, ci, co in shapes)
    mma_f = sum(2 * nq * K * ci * co for nq, ns, H, ci, co in shapes)  # the tensor-core contraction alone
    return dict(kp_fwd=fwd_b, kp_fwd_dx=dx_b, kp_dw=dw_b, fwd_flops=fwd_f, mma_flops=mma_f)


def search_bytes(batch, idx_bytes=8):
    """12*Nq + 12*Ns + 8*B + idx_bytes*Nq*Hmax per call, summed over the pyramid's 13 searches."""
    tot = 0
    L = len(batch.points)
    rows = layer_rows(batch)
    for l in range(L):
        n = rows[l]
        B = batch.lengths[l].shape[0]
        tot += 24 * n + 8 * B + idx_bytes * n * batch.neighbors[l].shape[1]
        if l + 1 < L:
            m = rows[l + 1]
            tot += 12 * (n + m) + 8 * B + idx_bytes * m * batch.pools[l].shape[1]
            tot += 12 * (n + m) + 8 * B + idx_bytes * n * batch.upsamples[l].shape[1]
    return tot


# ---------------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; the product path has no CPU fallback "
                           "(use --impl reference for the CPU reference arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # the harness network's Linear layers use the same operand precision as the KPConv contraction (TF32 in, fp32 out)
    torch.backends.cuda.matmul.allow_tf32 = True
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from weasal_b200 import _lib, grid_subsampling, pyramid
    from weasal_b200.kpconv import KPConv
    from weasal_b200.distributed import GradAllReducer
    from weasal_b200.net import CfgView, KPFCNNHarness, net_config
    import ctypes as C

    def gpu_subsample(p, f, l, dl):
        return grid_subsampling.subsample(p, features=f, classes=l, sampleDl=dl)

    cfg, batches = build_batches(args.config, args.seed + 17 * rank, N_BATCHES, gpu_subsample)
    ncfg = net_config(args.config)
    view = CfgView(ncfg)
    np.random.seed(args.seed + rank)
    torch.manual_seed(args.seed)  # identical initial weights on every rank
    net = KPFCNNHarness(ncfg, KPConv).to(dev)
    net.train()
    # (fused: weight decay + momentum + update of all parameters in one multi-tensor kernel)
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.98, weight_decay=1e-3,
                          fused=os.environ.get("WEASAL_BENCH_FUSED_SGD", "1") == "1")
    reducer = GradAllReducer(net.parameters())

    dev_batches = [{k: torch.from_numpy(v).to(dev) for k, v in b.items() if k != "lengths"} for b in batches]
    pin_batches = [{k: torch.from_numpy(v).pin_memory() for k, v in b.items() if k != "lengths"} for b in batches]
    flush = torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)

    from weasal_b200.engine import GraphedTrainStep, calibrate_static_caps
    # Static shapes + one CUDA graph per step (weasal_b200/engine.py). Capacities come from a calibration pass over
    # the batches, like the reference's sampler calibration (batch_limit / neighborhood_limits): rows padded to a
    # per-layer capacity, neighbourhood limits chosen so that no row is cropped (results equal the unlimited pyramid).
    use_graph = os.environ.get("WEASAL_BENCH_GRAPH", "1") != "0"  # "eager": static batches, eager launches (ncu lists)
    n_cap = limits = plans = None
    from weasal_b200.engine import calibrate_conv_plans
    from weasal_b200.net import fused_linear_weights
    from weasal_b200.plan import WeightPacker
    # weight-only work (TF32 operand images of every KPConv / unary block): one launch per step
    packer = WeightPacker(net, fused_linear_weights(net)) if os.environ.get("WEASAL_BENCH_PACKER", "1") == "1" else None
    if use_graph:
        n_cap, limits = calibrate_static_caps(view, [b["points"] for b in dev_batches], [b["lengths"] for b in batches])
        # geometry-only work (influence lists, transposed tables of all 10 KPConv): built by the prefetch stage
        if os.environ.get("WEASAL_BENCH_PLANS", "1") == "1":
            plans = calibrate_conv_plans(net, view, [b["points"] for b in dev_batches], [b["lengths"] for b in batches],
                                         n_cap, limits)
    prefetch = pyramid.PyramidPrefetcher(view, dev, neighborhood_limits=limits, n_cap=n_cap, plans=plans)
    trainer = GraphedTrainStep(net, opt, F.cross_entropy, reducer=reducer if world > 1 else None, clip_value=100.0,
                               use_graph=os.environ.get("WEASAL_BENCH_GRAPH", "1") == "1", plans=plans, packer=packer)
    eager = GraphedTrainStep(net, opt, F.cross_entropy, reducer=None, clip_value=100.0, packer=packer)  # profile leg: no collective
    if use_graph:  # capture before the prefetch pipeline runs (nothing else issues CUDA work meanwhile)
        prefetch.submit(dev_batches[0]["points"], dev_batches[0]["features"], dev_batches[0]["labels"], batches[0]["lengths"], inputs_ready=True)
        trainer.prepare(prefetch.get())
        torch.cuda.synchronize()

    def net_step(batch, allreduce=True):
        return trainer.step(batch) if allreduce else eager._body(batch)

    loss_pinned = torch.zeros(4, dtype=torch.float32).pin_memory()
    stamps = []  # (host time at which a step was launched, its pyramid's build time, time get() waited) of the last run

    def run_steps(first, n, e2e, allreduce=True, clocks=None, on_batch=None):
        """n steps over batches first, first+1, ...: the pyramid of step t+1 is built on the prefetcher's side stream
        (one native call, the counterpart of the reference's DataLoader workers) while step t trains. The pipeline
        starts and ends empty, so exactly n pyramids and n training steps happen inside the call."""
        src = pin_batches if e2e else dev_batches  # e2e: the worker copies the batch from pinned host memory

        def submit(it):
            b = it % N_BATCHES
            prefetch.submit(src[b]["points"], src[b]["features"], src[b]["labels"], batches[b]["lengths"], inputs_ready=True)

        pts = 0
        stamps.clear()
        done = []  # one event per launched step: the host stays at most two steps ahead of the GPU
        ahead = os.environ.get("WEASAL_BENCH_PREFETCH", "1") != "0"  # 0: build each pyramid when its step starts (A/B)
        if n > 0 and ahead:
            submit(first)
        for it in range(first, first + n):
            flush.zero_()  # evict L2 between steps (256 MB > 126 MB L2)
            if not ahead:
                submit(it)
            batch = prefetch.get()
            if len(done) >= 2:
                ev0, slot0 = done.pop(0)
                ev0.synchronize()
                if e2e:
                    loss_host = float(loss_pinned[slot0])  # the device -> host read of that step's result
            t_l0 = time.perf_counter()
            stamps.append((t_l0, prefetch.stats[-1][0], prefetch.stats[-1][1]))
            loss = net_step(batch, allreduce)
            if e2e:  # the step's result travels to pinned host memory behind the step; it is read two steps later, so
                loss_pinned[it % 4].copy_(loss, non_blocking=True)  # the host never drains the GPU inside the loop
            ev = torch.cuda.Event()
            ev.record()
            done.append((ev, it % 4))
            if ahead and it + 1 < first + n:
                submit(it + 1)  # after this step's launch: the GPU starts on step t while the host prepares batch t+1
            if os.environ.get("WEASAL_DEBUG") and rank == 0:
                print(f"[bench] {'e2e' if e2e else 'dev'} step {it}: build {prefetch.stats[-1][0] * 1e3:.2f} ms, get() waited "
                      f"{prefetch.stats[-1][1] * 1e3:.2f} ms, net launches {(time.perf_counter() - t_l0) * 1e3:.2f} ms",
                      file=sys.stderr)
            if clocks is not None and (it - first) % max(n // 8, 1) == 0:
                clocks.sample()  # the step's kernels are still in flight here: a reading under load
            if on_batch is not None:
                on_batch(batch)
            pts += batches[it % N_BATCHES]["points"].shape[0]
        for ev0, slot0 in done:  # the last steps' results are read before the call returns (inside the timed region)
            ev0.synchronize()
            if e2e:
                loss_host = float(loss_pinned[slot0])
        return pts

    def timed(n_warm, n_steps, e2e, clocks=None):
        import gc
        gc.collect()
        gc.disable()  # the step is host-launch bound: a generational GC pause inside a step shows up as a 10 % outlier
        run_steps(0, n_warm, e2e)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pts = run_steps(n_warm, n_steps, e2e, clocks=clocks)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        gc.enable()
        total_ms = e0.elapsed_time(e1) if n_steps > 0 else 0.0
        t = torch.tensor([total_ms, float(pts)], dtype=torch.float64, device=dev)
        if world > 1:
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(tmax[0]), float(t[1])
        return float(t[0]), float(t[1])

    W, K = max(args.warmup, 3), args.steps
    L = _lib.lib()
    # set-up, before any warm-up or timed step: every distinct batch once through the real pipeline, so that the
    # worker thread's scratch arena and the allocator pools have seen the largest batch (a first-time cudaMalloc
    # inside the timed region stalls the whole device for milliseconds)
    run_steps(0, N_BATCHES, False)
    run_steps(0, N_BATCHES, True)
    torch.cuda.synchronize()
    clocks = ClockSampler(local)
    for _ in range(3):  # the first NVML reads of a process take 10-35 ms (seen as a one-off stall in the first timed step)
        clocks.sample()
    clocks.rows.clear()
    launches0 = _lib.launch_count()
    timed(W, 0, False)
    launches0 = _lib.launch_count()
    g0 = trainer.n_graphed
    prof_range = os.environ.get("WEASAL_BENCH_PROFILE_RANGE") == "1"  # ncu --profile-from-start off: the timed steps only
    if prof_range:
        torch.cuda.profiler.start()
    ms, pts = timed(0, K, False, clocks if rank == 0 else None)
    if prof_range:
        torch.cuda.profiler.stop()
    iv = np.diff([a[0] for a in stamps]) * 1e3 if len(stamps) > 2 else np.zeros(1)
    pacing = {"launch_interval_ms": {"min": float(iv.min()), "median": float(np.median(iv)), "max": float(iv.max())},
              "pyramid_build_ms_median": float(np.median([a[1] for a in stamps]) * 1e3) if stamps else None,
              "get_wait_ms_median": float(np.median([a[2] for a in stamps]) * 1e3) if stamps else None,
              "launch_intervals_ms": [round(float(v), 2) for v in iv]}
    # library kernels launched in the timed region: the pyramid's (counted live) + those inside the replayed graphs
    gpu_launches = _lib.launch_count() - launches0 + (trainer.n_graphed - g0) * trainer.launches_per_replay
    graphed_steps = trainer.n_graphed - g0
    e2e_ms, e2e_pts = timed(W, K, True, clocks if rank == 0 else None)
    clk = clocks.summary() if rank == 0 else None

    # per-kernel device times (CUDA events on the launching stream, recorded inside the library)
    roof, kernels = None, {}
    if rank == 0:
        L.kp_profile_enable(1)
        psteps = min(K, 8)
        shapes, sbytes = None, 0
        seen = []

        def first_shapes(batch):
            if not seen:
                seen.append((conv_shapes(batch, net), search_bytes(batch)))

        run_steps(0, psteps, False, allreduce=False, on_batch=first_shapes)  # rank 0 only: no collective in this leg
        shapes, sbytes = seen[0]
        torch.cuda.synchronize()
        L.kp_profile_enable(0)
        buf = C.create_string_buffer(1 << 16)
        n = L.kp_profile_read(buf, len(buf))
        for line in buf.value.decode().splitlines() if n > 0 else []:
            tag, cnt, tot = line.split()
            kernels[tag] = {"launches_per_step": int(cnt) / psteps, "ms_per_step": float(tot) / psteps}
        alg = algorithmic(shapes)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        alg_bytes = {"kp_fwd": alg["kp_fwd"], "kp_fwd_dx": alg["kp_fwd_dx"], "kp_dw": alg["kp_dw"],
                     "rs_search": sbytes}
        for tag, b in alg_bytes.items():
            if tag in kernels and kernels[tag]["ms_per_step"] > 0:
                kernels[tag]["algorithmic_mb_per_step"] = b / 1e6
                kernels[tag]["achieved_gbs"] = b / 1e9 / (kernels[tag]["ms_per_step"] / 1e3)
        if "kp_fwd" in kernels:
            kernels["kp_fwd"]["contraction_tflops"] = alg["mma_flops"] / 1e12 / (kernels["kp_fwd"]["ms_per_step"] / 1e3)
            kernels["kp_fwd"]["tensor_frac_of_bf16_sustained"] = kernels["kp_fwd"]["contraction_tflops"] / tf_peak
        cand = [t for t in alg_bytes if t in kernels]
        if cand:
            dom = max(cand, key=lambda t: kernels[t]["ms_per_step"])
            a = kernels[dom]["achieved_gbs"]
            # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (profiles/traffic.json,
            # written by tools/ncu_summarise.py); null when no capture of it has been committed
            traffic = None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                kname = {"kp_fwd": "kp_fwd_kernel", "kp_fwd_dx": "kp_fwd_kernel", "kp_dw": "kp_dw_kernel",
                         "rs_search": "rs_search_kernel"}[dom]
                traffic = tj[kname]["dram_bytes_per_launch"]
            except Exception:
                pass
            roof = {"kernel": dom, "bound": "hbm", "achieved": a, "peak": hbm_peak, "unit": "GB/s", "frac": a / hbm_peak,
                    "traffic": traffic, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                    "bytes_per_launch": alg_bytes[dom] / max(kernels[dom]["launches_per_step"], 1),
                    "launch_ms": kernels[dom]["ms_per_step"] / max(kernels[dom]["launches_per_step"], 1)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = run_reference(args, budget_s=25.0, as_leg=True)

    if rank == 0:
        n0 = float(np.mean([b["points"].shape[0] for b in batches]))
        h2d = int(np.mean([sum(v.nbytes for k, v in b.items()) for b in batches]))
        line = {
            "metric": METRIC, "value": pts / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32 (fp32 gather, TF32 tensor-core contraction, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": f"{args.config}: pyramid precompute + KPFCNN fwd+bwd+SGD on {cfg['batch_num']} "
                                   f"synthetic ALS spheres of radius {cfg['in_radius']} m per step",
                       "points_per_step_per_gpu": n0, "global_points_per_step": n0 * world,
                       "first_subsampling_dl": cfg["dl"], "layers": 5, "kpconv_per_forward": 10,
                       "parallelism": f"dp{world}", "l2": "flushed between steps (256 MB write, inside the timed region)",
                       "pyramid": "built one step ahead on a side stream by one native call (kp_pyramid_build_dev)",
                       "random_grid_orient": True,
                       "neighborhood_limits": limits if limits is None else f"calibrated, no row cropped: {limits}",
                       "step": (f"one CUDA graph per step over static-shape batches (rows padded to {n_cap}); "
                                f"{graphed_steps} of {K} timed steps graphed") if use_graph else "eager launches",
                       "harness_linear_precision": "tf32"},
            "e2e": {"value": e2e_pts / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / K},
            "pacing": pacing, "gpu_launches": int(gpu_launches), "gpu_launches_per_step": gpu_launches / max(K, 1),
            "clocks": clk, "roofline": roof, "kernels": kernels, "cpu_baseline": cpu,
            "grad_allreduce_bytes": reducer.bytes() if world > 1 else 0,
        }
        emit(line)
    prefetch.close()
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------- reference arm
def run_reference(args, budget_s=150.0, as_leg=False):
    """The reference's CPU implementation of the path: C++ cores (oracle/_ref, or the C restatement when the
    reference sources were not available at build time) for the pyramid, prefetched by worker threads the way the
    reference's DataLoader workers do, and the PyTorch CPU operator chain for the network, all host threads."""
    import torch
    import torch.nn.functional as F
    from concurrent.futures import ThreadPoolExecutor

    import oracle
    from oracle.kpconv_torch import KPConvTorch
    from oracle.pyramid_ref import segmentation_inputs_cpu
    from weasal_b200.net import CfgView, KPFCNNHarness, net_config
    from weasal_b200.pyramid import DeviceBatch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind = "reference" if oracle.ref_available() else "port"

    def cpu_subsample(p, f, l, dl):
        fn = oracle.ref_subsample if kind == "reference" else oracle.grid_subsample
        return fn(p, features=f, classes=l, sampleDl=dl)

    cfg, batches = build_batches(args.config, args.seed, N_BATCHES, cpu_subsample)
    ncfg = net_config(args.config)
    view = CfgView(ncfg)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    net = KPFCNNHarness(ncfg, KPConvTorch)
    net.train()
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.98, weight_decay=1e-3)

    def precompute(b):
        li = segmentation_inputs_cpu(b["points"], b["features"], b["labels"], b["lengths"], view,
                                     use_ref=(kind == "reference"))
        return [torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a for a in li]

    def train(li):
        batch = DeviceBatch(li)
        loss = F.cross_entropy(net(batch), batch.labels)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_value_(net.parameters(), 100.0)
        opt.step()
        return float(loss)

    # bounded sample: sub-batches of `bn` spheres so that the whole run fits the budget
    def sub(b, bn):
        n = int(b["lengths"][:bn].sum())
        return dict(points=b["points"][:n], lengths=b["lengths"][:bn], features=b["features"][:n], labels=b["labels"][:n])

    t0 = time.time()
    train(precompute(sub(batches[0], 1)))
    t_one = time.time() - t0  # one sphere, cold
    W = 1 if as_leg else max(args.warmup, 1)
    K = 2 if as_leg else args.steps
    bn = cfg["batch_num"]
    while bn > 1 and (W + K) * t_one * bn * 0.6 > budget_s:
        bn -= 1
    pool = ThreadPoolExecutor(max_workers=min(4, cores))
    order = [sub(batches[i % N_BATCHES], bn) for i in range(W + K)]
    futs = [pool.submit(precompute, b) for b in order[:3]]
    pts, t_timed = 0, 0.0
    for i in range(W + K):
        if i == W:
            t_start = time.time()
        li = futs[i].result()
        if i + 3 < W + K:
            futs.append(pool.submit(precompute, order[i + 3]))
        train(li)
        if i >= W:
            pts += order[i]["points"].shape[0]
    t_timed = time.time() - t_start
    pool.shutdown()
    value = pts / t_timed
    sample = (f"{K} steps of {bn} of {cfg['batch_num']} spheres ({pts // K} points/step): pyramid on the "
              f"{'compiled reference C++ cores (oracle/_ref)' if kind == 'reference' else 'C restatement (oracle/)'} "
              f"prefetched by {min(4, cores)} worker threads + PyTorch CPU KPConv chain fwd+bwd+SGD")
    leg = {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
    if as_leg:
        return leg
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": K, "warmup": W, "ms_per_step": t_timed / K * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: pyramid precompute + KPFCNN fwd+bwd+SGD, reference CPU path",
                       "points_per_step": pts / K, "spheres_per_step": bn},
            "cpu_baseline": leg, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) was sent to stderr."""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # libraries that print to stdout (e.g. the NCCL version banner) must not break the one-line contract
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="vaihingen_pl", choices=["vaihingen_pl", "dales_pl"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if os.environ.get("WEASAL_BENCH_WATCHDOG"):  # debugging aid: dump every thread's stack and exit after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["WEASAL_BENCH_WATCHDOG"]), exit=True)
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return  # under torchrun only rank 0 runs the CPU reference
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
