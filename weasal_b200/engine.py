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


class GraphedTrainStep:
    """forward -> loss -> backward (-> gradient all-reduce) -> clip -> optimizer step of a static-shape batch, captured
    once into a CUDA graph and replayed per batch.

        trainer = GraphedTrainStep(net, optimizer, loss_fn, reducer=reducer, clip_value=100.0)
        loss = trainer.step(batch)          # batch from PyramidPrefetcher(..., n_cap=..., neighborhood_limits=...)

    ``loss_fn(logits, labels)``; ``net(batch)`` consumes a :class:`DeviceBatch`. The returned loss is a device scalar
    (the graph's static output for graphed steps)."""

    def __init__(self, net, optimizer, loss_fn, reducer=None, clip_value=None, warmup=3, use_graph=True):
        self.use_graph = use_graph  # False: every batch takes the eager step (profiling under ncu)
        self.net, self.opt, self.loss_fn, self.reducer, self.clip, self.warmup = net, optimizer, loss_fn, reducer, clip_value, warmup
        self.graph = None
        self.slab = None          # the graph's input: one static slab, every tensor of the batch is a view of it
        self.layout = None        # (offsets, n_cap, strides) of the captured layout
        self.loss = None
        self.launches_per_replay = 0   # library kernels inside one replay (bench.py's gpu_launches)
        self.n_graphed = self.n_eager = 0
        self.tail_in_graph = True

    # ------------------------------------------------------------------------------------------------------ the step
    def _head(self, batch):
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
        return batch

    def _capture(self, batch):
        nbld, dev = batch.build, batch.static_slab.device
        self.slab = torch.empty_like(batch.static_slab)
        self.layout = (nbld.offs.copy(), nbld.n_cap.copy(), nbld.strides.copy())
        params = [p for p in self.net.parameters()]
        saved = [p.detach().clone() for p in params]
        had_state = {id(p) for p in params if self.opt.state.get(p)}
        cur = torch.cuda.current_stream(dev)
        s = torch.cuda.Stream(dev)
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            self.slab.copy_(batch.static_slab)
            for _ in range(self.warmup):  # sizes the allocator pools, the library's scratch arena (per stream) and the
                self._body(self._static_batch(nbld))  # optimizer's momentum buffers before anything is recorded
        cur.wait_stream(s)
        torch.cuda.synchronize(dev)
        self.opt.zero_grad(set_to_none=True)
        g = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        # With a gradient all-reduce the graph ends after backward and the collective + clip + optimizer (a handful
        # of multi-tensor launches) stay eager: the collective keeps its place in the process group's own stream order.
        self.tail_in_graph = self.reducer is None
        with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
            self.loss = self._body(self._static_batch(nbld)) if self.tail_in_graph else self._head(self._static_batch(nbld))
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
            return self._body(batch)
        if self.graph is None:
            self._capture(batch)
        self.slab.copy_(batch.static_slab, non_blocking=True)
        self.graph.replay()
        if not self.tail_in_graph:
            self._tail()
        self.n_graphed += 1
        return self.loss
