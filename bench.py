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
N_BATCHES = 8           # distinct sphere batches cycled through the timed steps
N_CALIB = 8             # batches of ANOTHER tile the static capacities are calibrated on (never timed)
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
def measure_config(args, cfg_name, ctx, K, W, primary):
    """One workload (``vaihingen_pl`` / ``dales_pl``): K timed steps device-resident, K timed steps end to end; for the
    primary workload also the per-kernel profile leg. Returns the fields of the JSON line that depend on the workload."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from weasal_b200 import _lib, grid_subsampling, pyramid
    from weasal_b200.distributed import GradAllReducer
    from weasal_b200.engine import GraphedTrainStep, calibrate_conv_plans, calibrate_static_caps
    from weasal_b200.kpconv import KPConv
    from weasal_b200.net import CfgView, KPFCNNHarness, fused_linear_weights, net_config
    from weasal_b200.plan import WeightPacker

    rank, world, dev, flush = ctx["rank"], ctx["world"], ctx["dev"], ctx["flush"]

    def gpu_subsample(p, f, l, dl):
        return grid_subsampling.subsample(p, features=f, classes=l, sampleDl=dl)

    # timed batches and calibration batches come from DIFFERENT synthetic tiles: the capacities never saw the timed spheres
    cfg, batches = build_batches(cfg_name, args.seed + 17 * rank, N_BATCHES, gpu_subsample)
    _, calib = build_batches(cfg_name, args.seed + 17 * rank + 7919, N_CALIB, gpu_subsample)
    ncfg = net_config(cfg_name)
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
    cal_pts = [torch.from_numpy(b["points"]).to(dev) for b in calib]
    cal_lens = [b["lengths"] for b in calib]

    # Static shapes + one CUDA graph per step (weasal_b200/engine.py). Capacities come from a calibration pass, like the
    # reference's sampler calibration (batch_limit / neighborhood_limits, datasets/Vaihingen3D_PseudoLabel.py:1098-1404):
    # rows padded to a per-layer capacity, neighbourhood limits chosen so that no calibration row is cropped. A timed
    # batch that outgrows a capacity, or has a cropped row, takes the eager step (counted below).
    use_graph = os.environ.get("WEASAL_BENCH_GRAPH", "1") != "0"  # "eager": static batches, eager launches (ncu lists)
    n_cap = limits = plans = None
    # weight-only work (TF32 operand images of every KPConv / unary block): one launch per step
    packer = WeightPacker(net, fused_linear_weights(net)) if os.environ.get("WEASAL_BENCH_PACKER", "1") == "1" else None
    if use_graph:
        n_cap, limits = calibrate_static_caps(view, cal_pts, cal_lens, row_margin=1.10, width_margin=0.2)
        # geometry-only work (influence lists, transposed tables of all 10 KPConv): built by the prefetch stage
        if os.environ.get("WEASAL_BENCH_PLANS", "1") == "1":
            plans = calibrate_conv_plans(net, view, cal_pts, cal_lens, n_cap, limits)
    del cal_pts
    n_workers = int(os.environ.get("WEASAL_BENCH_WORKERS", "2"))  # builds in flight (= batches submitted ahead)
    prefetch = pyramid.PyramidPrefetcher(view, dev, neighborhood_limits=limits, n_cap=n_cap, plans=plans, workers=n_workers)
    trainer = GraphedTrainStep(net, opt, F.cross_entropy, reducer=reducer if world > 1 else None, clip_value=100.0,
                               use_graph=os.environ.get("WEASAL_BENCH_GRAPH", "1") == "1", plans=plans, packer=packer)
    eager = GraphedTrainStep(net, opt, F.cross_entropy, reducer=None, clip_value=100.0, packer=packer)  # profile leg: no collective
    if use_graph:  # capture before the prefetch pipeline runs (nothing else issues CUDA work meanwhile)
        for b in range(N_BATCHES):  # (the first timed batch that fits the capacities)
            prefetch.submit(dev_batches[b]["points"], dev_batches[b]["features"], dev_batches[b]["labels"],
                            batches[b]["lengths"], inputs_ready=True)
            first = prefetch.get()
            trainer.prepare(first)
            if trainer.graph is not None:
                break
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
        depth = n_workers if ahead else 0
        for k in range(min(depth, n)):
            submit(first + k)
        for it in range(first, first + n):
            flush.zero_()  # evict L2 between steps (256 MB > 126 MB L2)
            if not ahead:
                submit(it)
            batch = prefetch.get()
            t_w = 0.0
            if len(done) >= 2:
                ev0, slot0 = done.pop(0)
                t_w0 = time.perf_counter()
                ev0.synchronize()
                t_w = time.perf_counter() - t_w0   # host idle, waiting for the GPU: ~0 would mean the HOST paces the loop
                if e2e:
                    loss_host = float(loss_pinned[slot0])  # the device -> host read of that step's result
            t_l0 = time.perf_counter()
            stamps.append((t_l0, prefetch.stats[-1][0], prefetch.stats[-1][1], t_w))
            loss = net_step(batch, allreduce)
            if e2e:  # the step's result travels to pinned host memory behind the step; it is read two steps later, so
                loss_pinned[it % 4].copy_(loss, non_blocking=True)  # the host never drains the GPU inside the loop
            ev = torch.cuda.Event()
            ev.record()
            done.append((ev, it % 4))
            if ahead and it + depth < first + n:
                submit(it + depth)  # after this step's launch: the GPU starts on step t while the next batches are built
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
        gc.disable()  # a generational GC pause inside a step shows up as a 10 % outlier
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

    L = _lib.lib()
    # set-up, before any warm-up or timed step: every distinct batch once through the real pipeline, so that the
    # worker thread's scratch arena and the allocator pools have seen the largest batch (a first-time cudaMalloc
    # inside the timed region stalls the whole device for milliseconds)
    run_steps(0, N_BATCHES, False)
    run_steps(0, N_BATCHES, True)
    torch.cuda.synchronize()
    clocks = ctx["clocks"] if primary else None
    timed(W, 0, False)
    launches0 = _lib.launch_count()
    g0, e0_ = trainer.n_graphed, trainer.n_eager
    prof_range = primary and os.environ.get("WEASAL_BENCH_PROFILE_RANGE") == "1"  # ncu --profile-from-start off
    if prof_range:
        torch.cuda.profiler.start()
    ms, pts = timed(0, K, False, clocks if rank == 0 else None)
    if prof_range:
        torch.cuda.profiler.stop()
    iv = np.diff([a[0] for a in stamps]) * 1e3 if len(stamps) > 2 else np.zeros(1)
    pacing = {"launch_interval_ms": {"min": float(iv.min()), "median": float(np.median(iv)), "max": float(iv.max())},
              "pyramid_build_ms_median": float(np.median([a[1] for a in stamps]) * 1e3) if stamps else None,
              "get_wait_ms_median": float(np.median([a[2] for a in stamps]) * 1e3) if stamps else None,
              "host_waits_for_gpu_ms_median": float(np.median([a[3] for a in stamps]) * 1e3) if stamps else None}
    # library kernels launched in the timed region: the pyramid's (counted live) + those inside the replayed graphs
    gpu_launches = _lib.launch_count() - launches0 + (trainer.n_graphed - g0) * trainer.launches_per_replay
    graphed_steps, eager_steps = trainer.n_graphed - g0, trainer.n_eager - e0_
    e2e_ms, e2e_pts = timed(W, K, True, clocks if rank == 0 else None)

    # diagnostic, outside every timed region: the two halves of the pipeline each on an otherwise idle GPU. Their sum
    # against ms_per_step says how much of the side stream's work the step hides.
    if rank == 0 and primary and world == 1 and use_graph and trainer.graph is not None:
        def build_only(n):
            last = None
            for it in range(n):
                b = it % N_BATCHES
                prefetch.submit(dev_batches[b]["points"], dev_batches[b]["features"], dev_batches[b]["labels"],
                                batches[b]["lengths"], inputs_ready=True)
                last = prefetch.get()
            torch.cuda.synchronize()
            return last
        last = build_only(2)
        t0 = time.perf_counter()
        last = build_only(10)
        t_pyr = (time.perf_counter() - t0) / 10 * 1e3
        for _ in range(3):
            trainer.step(last)
        torch.cuda.synchronize()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for _ in range(10):
            trainer.step(last)
        eb.record()
        torch.cuda.synchronize()
        t_graph = ea.elapsed_time(eb) / 10
        ea.record()
        for _ in range(10):
            flush.zero_()
            trainer.step(last)
        eb.record()
        torch.cuda.synchronize()
        pacing["alone_ms"] = {"pyramid_and_lists_one_at_a_time": t_pyr, "training_graph": t_graph,
                              "training_graph_after_l2_flush": ea.elapsed_time(eb) / 10}

    # per-kernel device times (CUDA events on the launching stream, recorded inside the library)
    roof, kernels = None, {}
    if rank == 0 and primary:
        L.kp_profile_enable(1)
        psteps = min(K, 8)
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
        peaks = ctx["peaks"]
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        alg_bytes = {"kp_fwd": alg["kp_fwd"], "kp_fwd_dx": alg["kp_fwd_dx"], "kp_dw": alg["kp_dw"],
                     "rs_search": sbytes}
        for tag, b in alg_bytes.items():
            if tag in kernels and kernels[tag]["ms_per_step"] > 0:
                kernels[tag]["algorithmic_mb_per_step"] = b / 1e6
                kernels[tag]["achieved_gbs"] = b / 1e9 / (kernels[tag]["ms_per_step"] / 1e3)
                kernels[tag]["frac_of_hbm_peak"] = kernels[tag]["achieved_gbs"] / hbm_peak
        if "kp_fwd" in kernels:
            kernels["kp_fwd"]["contraction_tflops"] = alg["mma_flops"] / 1e12 / (kernels["kp_fwd"]["ms_per_step"] / 1e3)
            kernels["kp_fwd"]["tensor_frac_of_bf16_sustained"] = kernels["kp_fwd"]["contraction_tflops"] / tf_peak
        # forward and dX are the SAME kernel (kp_fwd_kernel: dX runs it over the transposed table with W^T), and ncu lists
        # them under one name: the roofline candidate is the kernel, not the call site
        if "kp_fwd" in kernels and "kp_fwd_dx" in kernels:
            both = [kernels["kp_fwd"], kernels["kp_fwd_dx"]]
            kernels["kp_fwd_kernel"] = {
                "launches_per_step": sum(k["launches_per_step"] for k in both), "ms_per_step": sum(k["ms_per_step"] for k in both),
                "algorithmic_mb_per_step": sum(k["algorithmic_mb_per_step"] for k in both), "tags": ["kp_fwd", "kp_fwd_dx"]}
            kk = kernels["kp_fwd_kernel"]
            kk["achieved_gbs"] = kk["algorithmic_mb_per_step"] / 1e3 / (kk["ms_per_step"] / 1e3)
            kk["frac_of_hbm_peak"] = kk["achieved_gbs"] / hbm_peak
            alg_bytes["kp_fwd_kernel"] = alg_bytes["kp_fwd"] + alg_bytes["kp_fwd_dx"]
        cand = [t for t in alg_bytes if t in kernels and not ("kp_fwd_kernel" in kernels and t in ("kp_fwd", "kp_fwd_dx"))]
        if cand:
            dom = max(cand, key=lambda t: kernels[t]["ms_per_step"])
            a = kernels[dom]["achieved_gbs"]
            # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (profiles/traffic.json,
            # written by tools/ncu_summarise.py); null when no capture of it has been committed
            traffic = None
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
                kname = {"kp_fwd": "kp_fwd_kernel", "kp_fwd_dx": "kp_fwd_kernel", "kp_fwd_kernel": "kp_fwd_kernel",
                         "kp_dw": "kp_dw_kernel", "rs_search": "rs_search_kernel"}[dom]
                traffic = tj[kname]["dram_bytes_per_launch"]
            except Exception:
                pass
            roof = {"kernel": dom, "bound": "hbm", "achieved": a, "peak": hbm_peak, "unit": "GB/s", "frac": a / hbm_peak,
                    "traffic": traffic, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                    "bytes_per_launch": alg_bytes[dom] / max(kernels[dom]["launches_per_step"], 1),
                    "launch_ms": kernels[dom]["ms_per_step"] / max(kernels[dom]["launches_per_step"], 1)}
    prefetch.close()
    n0 = float(np.mean([b["points"].shape[0] for b in batches]))
    h2d = int(np.mean([sum(v.nbytes for k, v in b.items()) for b in batches]))
    res = {
        "value": pts / (ms / 1e3), "ms_per_step": ms / K,
        "config": {"workload": f"{cfg_name}: pyramid precompute + KPFCNN fwd+bwd+SGD on {cfg['batch_num']} "
                               f"synthetic ALS spheres of radius {cfg['in_radius']} m per step",
                   "points_per_step_per_gpu": n0, "global_points_per_step": n0 * world,
                   "first_subsampling_dl": cfg["dl"], "first_features_dim": ncfg["first_features_dim"], "layers": 5,
                   "kpconv_per_forward": 10, "parallelism": f"dp{world}",
                   "l2": "flushed between steps (256 MB write, inside the timed region)",
                   "pyramid": f"built {n_workers} step(s) ahead by {n_workers} prefetch thread(s), each on its own side stream: one "
                              "native call (kp_pyramid_build_static_dev), then the influence lists of all KPConv "
                              "(kp_kpconv_prepare_dev)",
                   "random_grid_orient": True,
                   "capacities": f"calibrated on {N_CALIB} batches of another synthetic tile; timed on {N_BATCHES} unseen batches",
                   "neighborhood_limits": limits,
                   "step": (f"CUDA graph(s) per step over static-shape batches (rows padded to {n_cap}); "
                            f"{graphed_steps} of {K} timed steps graphed, {eager_steps} eager (capacity / crop fall-back)")
                           if use_graph else "eager launches",
                   "n_graphed": graphed_steps, "n_eager": eager_steps,
                   "harness_linear_precision": "tf32"},
        "e2e": {"value": e2e_pts / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms / K},
        "pacing": pacing, "gpu_launches": int(gpu_launches), "gpu_launches_per_step": gpu_launches / max(K, 1),
        "roofline": roof, "kernels": kernels, "grad_allreduce_bytes": reducer.bytes() if world > 1 else 0,
    }
    del trainer, eager, prefetch, net, opt, dev_batches, pin_batches
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

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
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    clocks = ClockSampler(local)
    for _ in range(3):  # the first NVML reads of a process take 10-35 ms (seen as a one-off stall in the first timed step)
        clocks.sample()
    clocks.rows.clear()
    ctx = {"rank": rank, "world": world, "dev": dev, "peaks": peaks, "clocks": clocks,
           "flush": torch.empty(L2_FLUSH_BYTES // 4, dtype=torch.float32, device=dev)}
    W, K = max(args.warmup, 3), args.steps
    main = measure_config(args, args.config, ctx, K, W, primary=True)
    clk = clocks.summary() if rank == 0 else None

    # the other workload north_star names, in the same line (DALES-shaped PseudoLabel: dl 0.4, first_features_dim 128,
    # KPConv up to 512 -> 512), at every N, so that the scaling run carries it too
    extra = {}
    other = "dales_pl" if args.config == "vaihingen_pl" else "vaihingen_pl"
    if not args.no_extra_configs:
        r = measure_config(args, other, ctx, min(K, 10), 3, primary=False)
        extra[other] = {"metric": METRIC, "unit": UNIT, "value": r["value"], "ms_per_step": r["ms_per_step"],
                        "steps": min(K, 10), "e2e": r["e2e"], "config": r["config"],
                        "gpu_launches_per_step": r["gpu_launches_per_step"]}
    sweep = None
    if rank == 0 and world == 1 and not args.no_sweep:
        try:
            sweep = operator_sweep(dev, peaks)
        except Exception as e:  # the sweep must never cost the headline line
            sweep = {"error": repr(e)[:300]}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = run_reference(args, budget_s=25.0, as_leg=True)

    if rank == 0:
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32 (fp32 gather, TF32 tensor-core contraction, fp32 accumulate)", "data": "synthetic",
            "config": main["config"], "e2e": main["e2e"], "pacing": main["pacing"], "gpu_launches": main["gpu_launches"],
            "gpu_launches_per_step": main["gpu_launches_per_step"], "clocks": clk, "roofline": main["roofline"],
            "kernels": main["kernels"], "cpu_baseline": cpu, "grad_allreduce_bytes": main["grad_allreduce_bytes"],
            "configs": {**extra, **({"operator_sweep": sweep} if sweep is not None else {})},
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# -------------------------------------------------------------------------------------------------- operator sweep
def operator_sweep(dev, peaks, budget_s=75.0):
    """BASELINE.json configs[4]: batch radius search + grid subsampling on 1e5 .. 1e7 raw points (B = 1 and 8) and KPConv
    K = 15 at Cin = Cout in {64, 128, 256, 512}, H in {16, 32, 64}, each row next to the reference's CPU implementation
    (compiled reference cores oracle/_ref; the reference's aten chain on PyTorch CPU) timed on a bounded sample, with the
    achieved GB/s against the SURVEY.md section-8d algorithmic bytes. Rank 0, one GPU."""
    import torch
    import oracle
    from oracle.kpconv_torch import kpconv_reference_ops
    from weasal_b200 import ops
    from weasal_b200.synthetic import make_als_tile
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    tf = float(peaks.get("bf16_tflops", 1590.0))
    have_ref = oracle.ref_available()
    t_start = time.time()
    rows = []

    def timeit(fn, warm=2, reps=5):
        for _ in range(warm):
            out = fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts)), out

    def cpu_rate(fn, units):
        t0 = time.time()
        fn()
        return units / max(time.time() - t0, 1e-9)

    geo = {}
    for n_raw in (100_000, 1_000_000, 10_000_000):
        if time.time() - t_start > budget_s * 0.5 and n_raw > 1_000_000:
            rows.append({"op": "precompute", "N_raw": n_raw, "skipped": "sweep time budget"})
            continue
        extent = float(np.sqrt(n_raw / 40.0))  # DALES-like raw density, 40 pts/m^2
        pts, _, _ = make_als_tile(1, extent, 40.0)
        order = np.argsort(pts[:, 0], kind="stable")
        for B in (1, 8):
            if B == 8 and n_raw != 1_000_000:
                continue
            P_np = pts if B == 1 else pts[order]   # B = 8: eight x-slabs as batch elements
            L = np.array([len(pts)], np.int32) if B == 1 else np.diff(np.linspace(0, len(pts), B + 1).astype(np.int64)).astype(np.int32)
            P = torch.from_numpy(np.ascontiguousarray(P_np)).to(dev)
            for mode in ("reference", "first"):
                ms, (sp, sl) = timeit(lambda: ops.grid_subsample(P, L, sampleDl=0.4, order=mode), warm=1, reps=3)
                nbytes = 12 * len(pts) + 12 * len(sp) + 8 * B
                row = {"op": "grid_subsample", "order": mode, "N": len(pts), "B": B, "M": int(len(sp)), "ms": ms,
                       "Mpts_per_s": len(pts) / ms / 1e3, "algorithmic_GBs": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / hbm}
                if mode == "reference" and have_ref and B == 1:
                    ns = min(len(pts), 1_000_000)
                    row["cpu_Mpts_per_s"] = cpu_rate(lambda: oracle.ref_subsample(P_np[:ns], sampleDl=0.4), ns) / 1e6
                    row["cpu_sample"] = f"reference grid_subsampling on the first {ns} points, 1 thread"
                rows.append(row)
            S = sp.contiguous()
            Ls = np.asarray(sl, np.int32)
            ms, nb = timeit(lambda: ops.batch_query(S, S, Ls, Ls, 1.0, dtype=torch.int32, cap_hint=64), warm=1, reps=3)
            nbytes = 24 * len(S) + 8 * B + 4 * len(S) * nb.shape[1]
            row = {"op": "batch_query", "N": int(len(S)), "B": B, "radius": 1.0, "Hmax": int(nb.shape[1]), "ms": ms,
                   "Mqueries_per_s": len(S) / ms / 1e3, "algorithmic_GBs": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / hbm}
            if have_ref and B == 1:
                S_np = S.cpu().numpy()
                nq = min(len(S_np), 100_000)
                row["cpu_Mqueries_per_s"] = cpu_rate(lambda: oracle.ref_batch_neighbors(S_np[:nq], S_np, np.array([nq], np.int32), Ls, 1.0), nq) / 1e6
                row["cpu_sample"] = f"reference batch_nanoflann_neighbors, first {nq} queries against all supports, 1 thread"
            rows.append(row)
            if B == 1:
                geo[n_raw] = (S, Ls)
            del P
    # KPConv: geometry = the subsampled clouds above (dl 0.4); H = the `limit` closest neighbours of a search of radius
    # 1.0 = conv_radius 2.5 x dl; kernel points as the reference lays them out (one at the centre, 14 on a shell of 0.66 x
    # that radius, kernel_points.py:484) with influence extent KP_extent 1.0 x dl (blocks.py:533-535)
    R_CONV, EXTENT = 1.0, 0.4
    for n_raw, Cs, Hs in ((100_000, (64, 128, 256, 512), (16, 32, 64)), (1_000_000, (64, 128, 256, 512), (32,)),
                          (10_000_000, (64,), (32,))):
        if n_raw not in geo:
            continue
        S, Ls = geo[n_raw]
        n = len(S)
        for H in Hs:
            nb = ops.batch_query(S, S, Ls, Ls, R_CONV, limit=H, dtype=torch.int32).contiguous()
            shadow = float((nb == n).float().mean())
            for Cc in Cs:
                if time.time() - t_start > budget_s:
                    rows.append({"op": "kpconv", "N": n, "H": H, "C": Cc, "skipped": "sweep time budget"})
                    continue
                x = torch.randn(n, Cc, device=dev, requires_grad=True)
                w = (torch.randn(15, Cc, Cc, device=dev) / Cc ** 0.5).requires_grad_(True)
                v = torch.randn(15, 3, device=dev)
                kp = v / v.norm(dim=1, keepdim=True) * 0.66 * R_CONV
                kp[0] = 0
                ms_f, y = timeit(lambda: ops.kpconv(S, S, nb, x, w, kp, EXTENT), warm=2, reps=3)
                g = torch.randn_like(y)

                def fb():
                    yy = ops.kpconv(S, S, nb, x, w, kp, EXTENT)
                    yy.backward(g)
                    return yy
                ms_fb, _ = timeit(fb, warm=2, reps=3)
                bytes_f = 4 * n * H + 24 * n + 4 * n * Cc * 2 + 4 * 15 * Cc * Cc
                flops = 2 * n * 15 * Cc * Cc
                row = {"op": "kpconv", "N": n, "H": H, "shadow_frac": shadow, "C": Cc, "fwd_ms": ms_f, "fwd_bwd_ms": ms_fb,
                       "fwd_algorithmic_GBs": bytes_f / ms_f / 1e6, "fwd_frac_hbm": bytes_f / ms_f / 1e6 / hbm,
                       "fwd_contraction_TFLOPs": flops / ms_f / 1e9, "fwd_frac_bf16_peak": flops / ms_f / 1e9 / tf,
                       "fwd_Mpts_per_s": n / ms_f / 1e3, "fwd_bwd_Mpts_per_s": n / ms_fb / 1e3}
                if n_raw == 100_000:
                    ns = 2048
                    q_c, nb_c = S[:ns].cpu(), nb[:ns].long().cpu()
                    s_c, x_c, w_c, kp_c = S.cpu(), x.detach().cpu().requires_grad_(True), w.detach().cpu().requires_grad_(True), kp.cpu()

                    def cpu_fb():
                        yy = kpconv_reference_ops(q_c, s_c, nb_c, x_c, w_c, kp_c, EXTENT)
                        yy.backward(torch.ones_like(yy))
                    cpu_fb()
                    row["cpu_fwd_bwd_Mpts_per_s"] = cpu_rate(cpu_fb, ns) / 1e6
                    row["cpu_sample"] = f"the reference's aten chain (models/blocks.py:277-374) fwd+bwd on {ns} query points, PyTorch CPU, all host threads"
                rows.append(row)
                del x, w, y, g
            del nb
    return {"rows": rows, "seconds": time.time() - t_start,
            "peaks": {"hbm_gbs": hbm, "bf16_tflops_burst": tf, "source": "MEASURED_PEAKS.json" if peaks else "fallback"}}


# ---------------------------------------------------------------------------------------------------- reference arm
def run_reference(args, budget_s=150.0, as_leg=False):
    """The reference's CPU implementation of the path on the box's host cores, same config as our arm (4 of 4 spheres per
    step). When a copy of the reference is on the machine (baseline/_ref, tools/install_reference.py) this is the
    reference ITSELF: its compiled C++ cores (oracle/_ref) behind its own ``segmentation_inputs``
    (datasets/common.py:461-577), its ``<DS>CustomBatch``, its ``models.architectures.KPFCNN`` built by its own config
    class, and the training step of utils/trainer_PseudoLabel.py:195-222 (forward, loss, backward, clip_grad_value_, SGD),
    the pyramid prefetched by ``input_threads`` worker threads like its DataLoader workers. Otherwise the restatements
    under oracle/ (``kind: "port"``)."""
    import importlib
    import torch
    import torch.nn.functional as F
    from concurrent.futures import ThreadPoolExecutor

    import oracle
    from oracle import ref_harness

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    threads = min(10, cores)   # the reference's input_threads = 10 (train_*_PseudoLabel.py)
    root = ref_harness.find_root() if oracle.ref_available() else None
    kind = "reference" if root is not None else "port"
    cfg_name = args.config

    def cpu_subsample(p, f, l, dl):
        fn = oracle.ref_subsample if oracle.ref_available() else oracle.grid_subsample
        return fn(p, features=f, classes=l, sampleDl=dl)

    cfg, batches = build_batches(cfg_name, args.seed, N_BATCHES, cpu_subsample)
    n_cls = int(cfg["num_classes"])
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    cwd = os.getcwd()
    if kind == "reference":
        ref_harness.install(root, backend="oracle_ref")
        script, cls, ds_mod, batch_cls = {
            "vaihingen_pl": ("train_Vaihingen3D_PseudoLabel", "Vaihingen3DPLConfig", "datasets.Vaihingen3D_PseudoLabel",
                             "Vaihingen3DPLCustomBatch"),
            "dales_pl": ("train_DALES_PseudoLabel", "DALESPLConfig", "datasets.DALES_PseudoLabel", "DALESPLCustomBatch"),
        }[cfg_name]
        rcfg = getattr(importlib.import_module(script), cls)()
        Batch = getattr(importlib.import_module(ds_mod), batch_cls)
        from datasets.common import PointCloudDataset
        from models.architectures import KPFCNN
        rcfg.num_classes = n_cls
        rcfg.class_w = [1.0] * n_cls
        ds = PointCloudDataset("x")
        ds.config = rcfg
        ds.neighborhood_limits = []
        net = KPFCNN(rcfg, list(range(n_cls)), [])
        net.train()
        opt = torch.optim.SGD(net.parameters(), lr=rcfg.learning_rate, momentum=rcfg.momentum, weight_decay=rcfg.weight_decay)

        def precompute(b):
            nb = len(b["lengths"])
            li = ds.segmentation_inputs(b["points"], b["features"], (b["labels"] % n_cls).astype(np.int64), b["lengths"])
            li += [np.ones((nb, 3), np.float32), np.tile(np.eye(3, dtype=np.float32), (nb, 1, 1)), np.zeros(nb, np.int32),
                   np.zeros(nb, np.int32), np.arange(len(b["points"]), dtype=np.int64)]
            return Batch([li])

        def train(batch):
            opt.zero_grad()
            out = net(batch, rcfg)
            loss = net.loss(out, batch.labels)
            loss.backward()
            if rcfg.grad_clip_norm > 0:
                torch.nn.utils.clip_grad_value_(net.parameters(), rcfg.grad_clip_norm)
            opt.step()
            return float(loss)
        what = ("the reference itself (baseline/_ref): segmentation_inputs on its compiled C++ cores, CustomBatch, "
                "models.architectures.KPFCNN, training step of utils/trainer_PseudoLabel.py:195-222")
    else:
        from oracle.kpconv_torch import KPConvTorch
        from oracle.pyramid_ref import segmentation_inputs_cpu
        from weasal_b200.net import CfgView, KPFCNNHarness, net_config
        from weasal_b200.pyramid import DeviceBatch
        ncfg = net_config(cfg_name)
        view = CfgView(ncfg)
        net = KPFCNNHarness(ncfg, KPConvTorch)
        net.train()
        opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.98, weight_decay=1e-3)

        def precompute(b):
            li = segmentation_inputs_cpu(b["points"], b["features"], b["labels"] % n_cls, b["lengths"], view,
                                         use_ref=oracle.ref_available())
            return DeviceBatch([torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a for a in li])

        def train(batch):
            loss = F.cross_entropy(net(batch), batch.labels)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_value_(net.parameters(), 100.0)
            opt.step()
            return float(loss)
        what = "restatement under oracle/ (no copy of the reference on this machine): C cores + aten chain, harness network"

    # always the full workload (4 of 4 spheres per step); the budget bounds the NUMBER of steps
    t0 = time.time()
    train(precompute(batches[0]))
    t_one = time.time() - t0  # one step, cold
    W = 1
    K = 2 if as_leg else max(2, min(args.steps, int(budget_s / max(t_one * 0.6, 1e-3)) - W))
    pool = ThreadPoolExecutor(max_workers=threads)
    order = [batches[i % N_BATCHES] for i in range(W + K)]
    depth = min(3, W + K)
    futs = [pool.submit(precompute, b) for b in order[:depth]]
    pts = 0
    for i in range(W + K):
        if i == W:
            t_start = time.time()
        batch = futs[i].result()
        if i + depth < W + K:
            futs.append(pool.submit(precompute, order[i + depth]))
        train(batch)
        if i >= W:
            pts += order[i]["points"].shape[0]
    t_timed = time.time() - t_start
    pool.shutdown()
    os.chdir(cwd)
    value = pts / t_timed
    sample = (f"{K} steps of {cfg['batch_num']} of {cfg['batch_num']} spheres ({pts // K} points/step), pyramid prefetched by "
              f"{threads} worker threads, network on {cores} host threads: {what}")
    leg = {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "input_threads": threads}
    if as_leg:
        return leg
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": K, "warmup": W, "ms_per_step": t_timed / K * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{cfg_name}: pyramid precompute + KPFCNN fwd+bwd+SGD on {cfg['batch_num']} synthetic ALS "
                                   f"spheres of radius {cfg['in_radius']} m per step, reference CPU path",
                       "points_per_step_per_gpu": pts / K, "spheres_per_step": int(cfg["batch_num"]),
                       "first_subsampling_dl": cfg["dl"], "steps_requested": args.steps},
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
    ap.add_argument("--no-sweep", action="store_true", help="skip the operator sweep (configs[4]) of the N = 1 line")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the second workload (DALES-shaped) of the line")
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
