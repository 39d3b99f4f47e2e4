"""Training-step engine for the hot path: static-shape batches + one CUDA graph per step.

Why: one Vaihingen batch is ~40k points; the network's forward + backward + SGD is ~700 kernel launches of a few
microseconds each, so a Python-driven step is bound by host launch work, not by the GPU. The reference bounds its
batches the same way this engine needs them bounded: a calibrated ``batch_limit`` caps the points per batch and
calibrated ``neighborhood_limits`` fix the width of every neighbour matrix (datasets/Vaihingen3D_PseudoLabel.py
Sampler.calibration; datasets/common.py:336-346). With those caps the pyramid builder emits every batch in ONE fixed
layout (kp_pyramid_build_static_dev: rows padded to a per-layer capacity, padded query rows have no neighbours, padded
labels are ``ignore_index``), the whole step is captured once and replayed per batch with a single host call.

Padded rows change nothing on the real rows: KPConv / max_pool / closest_pool of a row without neighbours is zero and
no real row refers to a padded one; the reference's BatchNorm is the identity on these 2-D features (blocks.py:453-463),
so there are no batch statistics to pollute; the loss ignores padded labels. (tests/test_boundary_cpu.py checks this
on the reference's own operator chain in float64: loss, real-row logits and every parameter gradient are identical.)
One reference quirk is visible through the fixed WIDTHS: ``max_pool`` pads with a zero row (blocks.py:104), so a shadow
entry contributes 0 to the maximum; a row that fills the batch's widest matrix has no shadow entry in the reference's
layout (width = Hmax of that batch) but has one here (width = the limit), i.e. its maximum is clamped at 0 — exactly
what the reference computes for the same sphere in a batch with a wider Hmax. KPConv and closest_pool are unaffected
(shadow terms are zeros in a sum / not the first column). The static layout therefore carries every matrix' true width
as a device scalar (``batch.pool_widths``) and max_pool ignores the columns beyond it (kp_max_pool_forward_width_dev),
which removes the difference.

A batch that does not fit the capacities, or whose rows were cropped by a limit (the symmetric-table shortcut of the
conv matrices then does not hold), takes the ordinary eager step with the same kernels.
"""
import numpy as np
import os

import torch

from . import _lib
from .pyramid import DeviceBatch, NativeBuild


def calibrate_static_caps(config, point_sets, length_sets, row_margin=1.05, row_quantum=256, width_margin=0.12,
                          random_grid_orient=True, passes=2):
    """Capacities that fit every given batch without cropping: per-layer row capacities ``n_cap`` and neighbourhood
    limits ``limits`` (the reference's calibrated ``neighborhood_limits``, here chosen so that no row is cropped:
    results equal the unlimited pyramid). ``point_sets``: CUDA [N,3] tensors, ``length_sets``: their batch lengths."""
    n_max = conv_w = pool_w = up_w = None
    for pts, lens in list(zip(point_sets, length_sets)) * (passes if random_grid_orient else 1):
        nb = NativeBuild(pts, lens, config, random_grid_orient=random_grid_orient)
        while True:
            slab = torch.empty(nb.slab_bytes(), dtype=torch.uint8, device=pts.device)
            if nb.run(slab, torch.cuda.current_stream(pts.device).cuda_stream):
                break
        L = nb.L
        w = nb.widths.reshape(3, L)
        if n_max is None:
            n_max, conv_w, pool_w, up_w = nb.n_out.copy(), w[0].copy(), w[1].copy(), w[2].copy()
        else:
            n_max = np.maximum(n_max, nb.n_out)
            conv_w, pool_w, up_w = np.maximum(conv_w, w[0]), np.maximum(pool_w, w[1]), np.maximum(up_w, w[2])
    L = len(n_max)
    n_cap = [int(-(-int(n * row_margin + 1) // row_quantum) * row_quantum) for n in n_max]
    limits = []
    for l in range(L):
        need = max(int(conv_w[l]), int(pool_w[l]), int(up_w[l - 1]) if l > 0 else 0)
        limits.append(need + max(4, int(np.ceil(need * width_margin))))
    return n_cap, limits


def calibrate_conv_plans(net, config, point_sets, length_sets, n_cap, limits, random_grid_orient=True, margin=1.3,
                         forward_only=False):
    """The fixed layout of the prefetched KPConv lists (weasal_b200.plan.ConvPlans) for ``net``: capacities of the entry
    buffers from a calibration pass over the given batches (records actually needed, plus a margin; a batch that
    outgrows them takes the eager step)."""
    from .plan import ConvPlans, conv_specs, measure_entries
    from .pyramid import DeviceBatch, segmentation_inputs
    specs = conv_specs(net)
    need = None
    for pts, lens in zip(point_sets, length_sets):
        li = segmentation_inputs(pts, None, None, lens, config, neighborhood_limits=limits,
                                 random_grid_orient=random_grid_orient, native=True)
        used = measure_entries(specs, DeviceBatch(li))
        need = used if need is None else [max(a, b) for a, b in zip(need, used)]
    caps = [int(v * margin) + 4096 for v in need]
    return ConvPlans(specs, n_cap, limits, limits, caps, forward_only=forward_only)


class GraphedTrainStep:
    """forward -> loss -> backward (-> gradient all-reduce) -> clip -> optimizer step of a static-shape batch, captured
    once into CUDA graphs and replayed per batch.

        trainer = GraphedTrainStep(net, optimizer, loss_fn, reducer=reducer, clip_value=100.0, plans=plans, packer=packer)
        loss = trainer.step(batch)          # batch from PyramidPrefetcher(..., n_cap=..., neighborhood_limits=..., plans=plans)

    ``loss_fn(logits, labels)``; ``net(batch)`` consumes a :class:`DeviceBatch`. The returned loss is a device scalar
    (the graph's static output for graphed steps).

    One GPU: ONE graph (weight packing, forward, loss, backward, clip, optimizer). With a ``reducer`` (data parallel): graph A
    = packing, forward, loss, backward and one multi-tensor copy of the gradients into a flat buffer; then ONE eager NCCL
    all-reduce (average) of that buffer; then graph B = clip + optimizer step reading gradients that are views of the flat
    buffer. Both graphs address the flat buffer and the captured gradient tensors directly, so an eager fall-back step in
    between (which re-creates ``p.grad``) cannot leave a later replay with stale gradients."""

    def __init__(self, net, optimizer, loss_fn, reducer=None, clip_value=None, warmup=3, use_graph=True, plans=None,
                 packer=None):
        self.use_graph = use_graph  # False: every batch takes the eager step (profiling under ncu)
        self.net, self.opt, self.loss_fn, self.reducer, self.clip, self.warmup = net, optimizer, loss_fn, reducer, clip_value, warmup
        self.plans, self.packer = plans, packer
        self.graph = self.graph_tail = None
        self.slab = None          # the graph's input: one static slab, every tensor of the batch is a view of it
        self.plan_buf = None      # ... and the prefetched KPConv lists of the batch
        self.layout = None        # (offsets, n_cap, strides) of the captured layout
        self.loss = None
        self.flat = None          # flat gradient buffer (data parallel)
        self.launches_per_replay = 0   # library kernels inside one replay (bench.py's gpu_launches)
        self.n_graphed = self.n_eager = 0
        self.tail_in_graph = True

    # ------------------------------------------------------------------------------------------------------ the step
    def _head(self, batch):
        if self.packer is not None:
            self.packer.pack()   # operand images of every layer's weights: one launch
        logits = self.net(batch)
        loss = self.loss_fn(logits, batch.labels)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        return loss.detach()  # (a live loss would keep the autograd graph, and with it every activation, alive)

    def _tail(self):
        if self.reducer is not None:
            self.reducer.step()
        if self.clip is not None:
            torch.nn.utils.clip_grad_value_(self.net.parameters(), self.clip)
        self.opt.step()

    def _body(self, batch):
        loss = self._head(batch)
        self._tail()
        return loss

    def _static_batch(self, nbld):
        """Fresh views of the graph's slab (fresh tensor objects: per-tensor caches such as KPConv's transposed tables
        must not survive from one batch's data to the next)."""
        P, Nn, Po, Up, Le = nbld.views(self.slab, mark_symmetric=True)
        f, lb = nbld.static_extras(self.slab)
        batch = DeviceBatch(P + Nn + Po + Up + Le + [f, lb])
        batch.pool_widths = nbld.static_pool_widths(self.slab)  # max_pool ignores the columns beyond the true width
        if self.plan_buf is not None:
            self.plans.attach(self.plan_buf, batch.neighbors, batch.pools)
        return batch

    def _capture(self, batch):
        nbld, dev = batch.build, batch.static_slab.device
        self.slab = torch.empty_like(batch.static_slab)
        self.plan_buf = torch.empty_like(batch.plan_buf) if (self.plans is not None and batch.plan_buf is not None) else None
        self.layout = (nbld.offs.copy(), nbld.n_cap.copy(), nbld.strides.copy())
        params = [p for p in self.net.parameters()]
        saved = [p.detach().clone() for p in params]
        had_state = {id(p) for p in params if self.opt.state.get(p)}
        cur = torch.cuda.current_stream(dev)
        # capture stream of high priority: kernel nodes keep the priority of the stream they were captured on, and the
        # step must win SM slots against the prefetch streams' builds (WEASAL_TRAIN_PRIORITY=0: no preference)
        s = torch.cuda.Stream(dev, priority=int(os.environ.get("WEASAL_TRAIN_PRIORITY", "-1")))
        self._capture_stream = s  # kept for the trainer's lifetime: the library's scratch arena is keyed by (thread,
        s.wait_stream(cur)        # stream) and the captured graph holds pointers into this stream's arena
        with torch.cuda.stream(s):
            self.slab.copy_(batch.static_slab)
            if self.plan_buf is not None:
                self.plan_buf.copy_(batch.plan_buf)
            from . import ops
            ops.CAPTURE_WARMUP[0] = True   # (side-stream forks that exist only in a captured step run here too)
            try:
                for _ in range(self.warmup):  # sizes the allocator pools, the library's scratch arenas (per thread and
                    self._body(self._static_batch(nbld))  # stream) and the optimizer's momentum buffers before recording
            finally:
                ops.CAPTURE_WARMUP[0] = False
        cur.wait_stream(s)
        torch.cuda.synchronize(dev)
        self.opt.zero_grad(set_to_none=True)
        g = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        self.tail_in_graph = self.reducer is None
        if self.tail_in_graph:
            with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
                self.loss = self._body(self._static_batch(nbld))
        else:
            # graph A ends with the gradients gathered into one flat buffer; graph B (clip + optimizer) reads them there
            with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
                self.loss = self._head(self._static_batch(nbld))
                with_grad = [p for p in params if p.grad is not None]
                self.flat = torch.empty(sum(p.numel() for p in with_grad), dtype=torch.float32, device=dev)
                views, o = [], 0
                for p in with_grad:
                    views.append(self.flat[o:o + p.numel()].view_as(p))
                    o += p.numel()
                torch._foreach_copy_(views, [p.grad for p in with_grad])
            self._grad_params, self._grad_views = with_grad, views
            for p, v in zip(with_grad, views):
                p.grad = v
            gt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gt, stream=s, pool=g.pool(), capture_error_mode="thread_local"):
                if self.clip is not None:
                    torch.nn.utils.clip_grad_value_(with_grad, self.clip)
                self.opt.step()
            self.graph_tail = gt
        self.launches_per_replay = _lib.launch_count() - n0
        self.graph = g
        # the warm-up steps must not count as training: parameters back to their values, momentum buffers that did
        # not exist before back to zero (a zero buffer reproduces the optimizer's first-step rule buf = grad)
        with torch.no_grad():
            for p, v in zip(params, saved):
                p.copy_(v)
            for p in params:
                st = self.opt.state.get(p)
                if st and id(p) not in had_state:
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
        torch.cuda.synchronize(dev)

    def _fits(self, batch):
        if not self.use_graph or getattr(batch, "static_slab", None) is None or not batch.no_crop:
            return False
        if self.plans is not None and getattr(batch, "plan_buf", None) is None:
            return False   # a list outgrew its calibrated capacity: this batch builds its lists on the fly (eager step)
        if self.layout is None:
            return True
        nbld = batch.build
        return (np.array_equal(nbld.offs, self.layout[0]) and np.array_equal(nbld.n_cap, self.layout[1])
                and np.array_equal(nbld.strides, self.layout[2]) and batch.static_slab.numel() == self.slab.numel())

    def prepare(self, batch):
        """Capture the graph on a first static batch (call while no other thread is issuing CUDA work)."""
        if self.graph is None and self._fits(batch):
            self._capture(batch)

    def step(self, batch):
        if not self._fits(batch):
            self.n_eager += 1
            if self.graph_tail is not None:   # the eager step uses its own gradient tensors
                self.opt.zero_grad(set_to_none=True)
            loss = self._body(batch)
            return loss
        if self.graph is None:
            self._capture(batch)
        self.slab.copy_(batch.static_slab, non_blocking=True)
        if self.plan_buf is not None:
            self.plan_buf.copy_(batch.plan_buf, non_blocking=True)
        self.graph.replay()
        if not self.tail_in_graph:
            self.reducer.allreduce_flat(self.flat)
            self.graph_tail.replay()
        self.n_graphed += 1
        return self.loss


class GraphedForward:
    """Inference counterpart of :class:`GraphedTrainStep` (sphere voting, utils/tester_PseudoLabel.py:149-195):
    ``post(net(batch))`` under ``no_grad`` captured once on a static-shape batch and replayed per batch; a batch that does
    not fit the captured layout runs eagerly. The weights must not change after construction (their operand images are
    packed once). ``run(batch)`` returns the rows of the batch's real points ``[batch.n_points, C]`` — for graphed batches a
    view of the graph's static output, valid until the next ``run``."""

    def __init__(self, net, post=None, plans=None, packer=None, warmup=2):
        self.net, self.post, self.plans, self.packer, self.warmup = net, post, plans, packer, warmup
        self.graph = self.slab = self.plan_buf = self.layout = self.out = None
        self.n_graphed = self.n_eager = 0
        if packer is not None:
            packer.pack()

    _static_batch = GraphedTrainStep._static_batch

    def _forward(self, batch):
        with torch.no_grad():
            y = self.net(batch)
            return self.post(y) if self.post is not None else y

    def _fits(self, batch):
        if getattr(batch, "static_slab", None) is None or not batch.no_crop:
            return False
        if self.plans is not None and getattr(batch, "plan_buf", None) is None:
            return False
        if self.layout is None:
            return True
        nbld = batch.build
        return (np.array_equal(nbld.offs, self.layout[0]) and np.array_equal(nbld.n_cap, self.layout[1])
                and np.array_equal(nbld.strides, self.layout[2]) and batch.static_slab.numel() == self.slab.numel())

    def _capture(self, batch):
        nbld, dev = batch.build, batch.static_slab.device
        self.slab = torch.empty_like(batch.static_slab)
        self.plan_buf = torch.empty_like(batch.plan_buf) if (self.plans is not None and batch.plan_buf is not None) else None
        self.layout = (nbld.offs.copy(), nbld.n_cap.copy(), nbld.strides.copy())
        cur = torch.cuda.current_stream(dev)
        s = torch.cuda.Stream(dev, priority=int(os.environ.get("WEASAL_TRAIN_PRIORITY", "-1")))
        self._capture_stream = s
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            self.slab.copy_(batch.static_slab)
            if self.plan_buf is not None:
                self.plan_buf.copy_(batch.plan_buf)
            from . import ops
            ops.CAPTURE_WARMUP[0] = True
            try:
                for _ in range(self.warmup):
                    self._forward(self._static_batch(nbld))
            finally:
                ops.CAPTURE_WARMUP[0] = False
        cur.wait_stream(s)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
            self.out = self._forward(self._static_batch(nbld))
        self.graph = g
        torch.cuda.synchronize(dev)

    def run(self, batch):
        if not self._fits(batch):
            self.n_eager += 1
            return self._forward(batch)[:batch.n_points]
        if self.graph is None:
            self._capture(batch)
        self.slab.copy_(batch.static_slab, non_blocking=True)
        if self.plan_buf is not None:
            self.plan_buf.copy_(batch.plan_buf, non_blocking=True)
        self.graph.replay()
        self.n_graphed += 1
        return self.out[:batch.n_points]
