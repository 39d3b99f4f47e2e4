"""Device-side pyramid builder: the GPU counterpart of ``PointCloudDataset.segmentation_inputs``
(datasets/common.py:461-577) with ``batch_neighbors`` (common.py:185-196), ``batch_grid_subsampling`` incl. its random
grid orientation (common.py:77-182) and ``big_neighborhood_filter`` (common.py:336-346).

The reference runs this inside forked DataLoader workers on the CPU; CUDA cannot be used there, so this version runs
in the main process on the training stream and returns the same flat list
``points*L + neighbors*L + pools*L + upsamples*L + lengths*L + [features, labels]`` with device tensors
(indices int64 as collated by the reference, common.py:551-553). :class:`DeviceBatch` exposes that list with the
field names of ``<DS>CustomBatch`` (datasets/Vaihingen3D_PseudoLabel.py:1407-1481) so the unchanged networks consume it.
"""
import os
import time

import numpy as np
import torch

from . import ops


def axis_angle_rotations(axis, angle):
    """Rotation matrices about unit ``axis`` [B,3] by ``angle`` [B] (Rodrigues); float64 [B,3,3].
    Same matrices as the reference's create_3D_rotations (kernels/kernel_points.py:43-74)."""
    axis = np.asarray(axis, np.float64)
    angle = np.asarray(angle, np.float64)
    c, s = np.cos(angle), np.sin(angle)
    v = 1.0 - c
    x, y, z = axis[:, 0], axis[:, 1], axis[:, 2]
    R = np.empty((len(angle), 3, 3), np.float64)
    vx = v * x
    R[:, 0, 0] = c + v * (x * x)
    R[:, 0, 1] = vx * y - s * z
    R[:, 0, 2] = vx * z + s * y
    R[:, 1, 0] = vx * y + s * z
    R[:, 1, 1] = c + v * (y * y)
    R[:, 1, 2] = v * y * z - s * x
    R[:, 2, 0] = vx * z - s * y
    R[:, 2, 1] = v * y * z + s * x
    R[:, 2, 2] = c + v * (z * z)
    return R


def random_grid_rotations(B):
    """The per-batch-element grid orientation of datasets/common.py:98-111, drawing from ``np.random`` in the same
    order (theta, phi, alpha) so a seeded run reproduces the reference's pyramid."""
    theta = np.random.rand(B) * 2 * np.pi
    phi = (np.random.rand(B) - 0.5) * np.pi
    u = np.vstack([np.cos(theta) * np.cos(phi), np.sin(theta) * np.cos(phi), np.sin(phi)])
    alpha = np.random.rand(B) * 2 * np.pi
    return axis_angle_rotations(u.T, alpha).astype(np.float32)


def batch_neighbors(queries, supports, q_batches, s_batches, radius, limit=None, dtype=torch.int64):
    return ops.batch_query(queries, supports, q_batches, s_batches, radius, limit=limit, dtype=dtype)


def batch_grid_subsampling(points, batches_len, sampleDl=0.1, max_p=0, random_grid_orient=True, order="reference"):
    rot = random_grid_rotations(len(batches_len)) if random_grid_orient else None
    s_points, s_len = ops.grid_subsample(points, batches_len, sampleDl=sampleDl, max_p=max_p, order=order, rot=rot)
    return s_points, s_len


class PyramidBuilder:
    """The walk over ``config.architecture`` of datasets/common.py:461-577, one pyramid layer per :meth:`step`.

    Every radius search is issued without a host sync (``ops.PendingSearches``); grids are shared between the searches
    that use the same supports and radius. ``resolve()`` reads the widths of the searches issued so far with one
    device->host copy and swaps the placeholders for the column-sliced index matrices.
    """

    def __init__(self, stacked_points, stack_lengths, config, neighborhood_limits=None, random_grid_orient=True,
                 order="reference", index_dtype=torch.int64, device="cuda"):
        self.dev = torch.device(device)
        self.cfg, self.order, self.dtype, self.orient = config, order, index_dtype, random_grid_orient
        t = stacked_points if torch.is_tensor(stacked_points) else torch.from_numpy(np.ascontiguousarray(stacked_points))
        self.pts = t.to(self.dev, non_blocking=True).to(torch.float32)
        self.lens = np.ascontiguousarray(stack_lengths.cpu().numpy() if torch.is_tensor(stack_lengths)
                                         else stack_lengths, dtype=np.int32)
        self.limits = list(neighborhood_limits) if neighborhood_limits is not None and len(neighborhood_limits) else None
        self.r_normal = config.first_subsampling_dl * config.conv_radius
        self.points, self.neighbors, self.pools, self.upsamples, self.lengths = [], [], [], [], []
        self.pending = ops.PendingSearches(self.dev)
        self.resolved = 0
        self.grids = {}
        self.blocks = list(config.architecture)
        self.pos = 0
        self.done = False
        # number of layers = pooling / strided blocks before the first 'global' / 'upsample' block, plus one
        self.n_layers = 0
        for b in self.blocks:
            if 'pool' in b or 'strided' in b or 'global' in b or 'upsample' in b:
                self.n_layers += 1
                if 'global' in b or 'upsample' in b:
                    break

    def _lim(self, layer):
        return int(self.limits[layer]) if self.limits is not None and layer < len(self.limits) else None

    def _grid(self, p, b, r):
        key = (p.data_ptr(), p.shape[0], float(r))
        if key not in self.grids:
            self.grids[key] = ops.SearchGrid(p, b, r)
        return self.grids[key]

    def _empty(self):
        return torch.zeros((0, 1), dtype=self.dtype, device=self.dev)

    def step(self):
        """Issue the work of the next layer (conv search, subsampling, pool and upsample searches)."""
        if self.done:
            return False
        cfg, layer_blocks = self.cfg, []
        while self.pos < len(self.blocks):
            block = self.blocks[self.pos]
            self.pos += 1
            if not ('pool' in block or 'strided' in block or 'global' in block or 'upsample' in block):
                layer_blocks.append(block)
                continue
            layer, pts, lens, r_normal = len(self.points), self.pts, self.lens, self.r_normal
            if layer_blocks:
                deform = any('deformable' in b for b in layer_blocks)
                r = r_normal * cfg.deform_radius / cfg.conv_radius if deform else r_normal
                conv_i = self.pending.add(pts, pts, lens, lens, r, limit=self._lim(layer), dtype=self.dtype,
                                          grid=self._grid(pts, lens, r))
            else:
                conv_i = self._empty()
            if 'pool' in block or 'strided' in block:
                dl = 2 * r_normal / cfg.conv_radius
                pool_p, pool_b = batch_grid_subsampling(pts, lens, sampleDl=dl, random_grid_orient=self.orient,
                                                        order=self.order)
                r = r_normal * cfg.deform_radius / cfg.conv_radius if 'deformable' in block else r_normal
                pool_i = self.pending.add(pool_p, pts, pool_b, lens, r, limit=self._lim(layer), dtype=self.dtype,
                                          grid=self._grid(pts, lens, r))
                up_i = self.pending.add(pts, pool_p, lens, pool_b, 2 * r, limit=self._lim(layer + 1), dtype=self.dtype,
                                        grid=self._grid(pool_p, pool_b, 2 * r))
            else:
                pool_i, up_i = self._empty(), self._empty()
                pool_p = torch.zeros((0, 3), dtype=torch.float32, device=self.dev)
                pool_b = np.zeros((0,), np.int32)
            self.points.append(pts)
            self.neighbors.append(conv_i)
            self.pools.append(pool_i)
            self.upsamples.append(up_i)
            self.lengths.append(torch.from_numpy(lens.copy()).to(self.dev, non_blocking=True))
            self.pts, self.lens = pool_p, pool_b
            self.r_normal *= 2
            if 'global' in block or 'upsample' in block:
                self.done = True
            return True
        self.done = True
        return False

    def resolve(self):
        found = self.pending.resolve(start=self.resolved)
        for lst in (self.neighbors, self.pools, self.upsamples):
            for i, v in enumerate(lst):
                if isinstance(v, int):
                    lst[i] = found[v - self.resolved]
        self.resolved += len(found)
        for key in [k for k in self.grids if k[0] != self.pts.data_ptr()]:
            del self.grids[key]  # only the grid over the newest layer can be needed again


def layer_plan(config):
    """Per-layer radii of the walk over ``config.architecture`` (datasets/common.py:468-567): conv search radius
    (0 = the layer has no conv block), pool search radius, upsample radius (2r) and the subsampling cell ``dl`` towards
    the next layer; float64 lists of length L."""
    r_normal = config.first_subsampling_dl * config.conv_radius
    conv_r, pool_r, up_r, dls = [], [], [], []
    layer_blocks = []
    for block in config.architecture:
        if not ('pool' in block or 'strided' in block or 'global' in block or 'upsample' in block):
            layer_blocks.append(block)
            continue
        if layer_blocks:
            deform = any('deformable' in b for b in layer_blocks)
            conv_r.append(r_normal * config.deform_radius / config.conv_radius if deform else r_normal)
        else:
            conv_r.append(0.0)
        if 'pool' in block or 'strided' in block:
            dls.append(2 * r_normal / config.conv_radius)
            pool_r.append(r_normal * config.deform_radius / config.conv_radius if 'deformable' in block else r_normal)
            up_r.append(2 * r_normal)
        else:
            dls.append(0.0)
            pool_r.append(0.0)
            up_r.append(0.0)
        r_normal *= 2
        layer_blocks = []
        if 'global' in block or 'upsample' in block:
            break
    return conv_r, pool_r, up_r, dls


def draw_grid_rotations(config, nb, random_grid_orient=True):
    """All grid orientations of one batch, drawn layer by layer in the reference's order (common.py:98-105 runs inside
    each batch_grid_subsampling call): float32 [pooled layers, nb, 3, 3], or None."""
    if not random_grid_orient:
        return None
    dls = layer_plan(config)[3]
    rots = [random_grid_rotations(nb) for d in dls if d > 0]
    return np.ascontiguousarray(np.stack(rots), dtype=np.float32) if rots else None


_SLAB_HINT = {}  # (n_layers, index bytes, cap) -> slab bytes per input point that sufficed last time


class StaticCapacityExceeded(RuntimeError):
    """A batch does not fit the static capacities (kp_pyramid_build_static_dev returned KP_ERR_CAPACITY)."""


class NativeBuild:
    """One kp_pyramid_build_dev call split into the three phases a prefetching caller runs on different threads:
    ``__init__`` (argument arrays; caller thread), ``run`` (the native call, GIL released; worker thread) and
    ``views`` (tensor views of the slab; consumer thread)."""

    def __init__(self, points, stack_lengths, config, neighborhood_limits=None, random_grid_orient=True,
                 order="reference", index_dtype=torch.int64, cap=80, rot=None, n_cap=None, features=None, labels=None,
                 label_pad=-100):
        """``n_cap`` (per-layer row capacities) selects the static-shape layout of kp_pyramid_build_static_dev;
        ``features`` / ``labels`` (CUDA tensors) are then padded into the slab as well."""
        if not points.is_cuda:
            raise RuntimeError("weasal_b200: tensors must be CUDA tensors (there is no CPU fallback)")
        self.dev = points.device
        self.pts = points.contiguous() if points.dtype == torch.float32 else points.float().contiguous()
        self.lens = np.ascontiguousarray(stack_lengths.cpu().numpy() if torch.is_tensor(stack_lengths) else stack_lengths,
                                         dtype=np.int32).reshape(-1)
        self.nb, self.n0 = len(self.lens), self.pts.shape[0]
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        self.conv_r, self.pool_r, self.up_r, self.dls = (f32(a) for a in layer_plan(config))
        self.L = L = len(self.conv_r)
        # (a prefetching caller draws the orientations itself, in submission order)
        self.rot = rot if rot is not None else draw_grid_rotations(config, self.nb, random_grid_orient)
        self.limits = None
        if neighborhood_limits is not None and len(neighborhood_limits):
            self.limits = np.zeros(L, np.int32)
            self.limits[:min(L, len(neighborhood_limits))] = np.asarray(neighborhood_limits, np.int64)[:L]
        self.order, self.dtype, self.cap = order, index_dtype, int(cap)
        self.isz = 8 if index_dtype == torch.int64 else 4
        self.offs = np.zeros(5 * L + 3, np.int64)
        self.n_cap = np.ascontiguousarray(n_cap, dtype=np.int32) if n_cap is not None else None
        self.feats = self.labs = None
        self.label_pad = int(label_pad)
        if self.n_cap is not None:
            if len(self.n_cap) != L:
                raise ValueError("n_cap needs one entry per layer")
            if features is not None:
                self.feats = features.contiguous() if features.dtype == torch.float32 else features.float().contiguous()
            if labels is not None:
                self.labs = labels.contiguous() if labels.dtype == torch.int64 else labels.long().contiguous()
        self.n_out, self.lens_out = np.zeros(L, np.int32), np.zeros(L * self.nb, np.int32)
        self.widths, self.strides = np.zeros(3 * L, np.int32), np.zeros(3 * L, np.int32)
        self.need_bytes, self.need_cap = np.zeros(1, np.int64), np.zeros(1, np.int32)

    def static_slab_bytes(self):
        """Exact size of the static layout (256-byte aligned ranges in the builder's order)."""
        al = lambda b: (int(b) + 255) & ~255
        nc, isz, L = self.n_cap, self.isz, self.L
        w = lambda l: int(self.limits[l]) if self.limits is not None and l < L and self.limits[l] > 0 else self.cap
        tot = al(int(nc[0]) * 12)
        if self.feats is not None:
            tot += al(int(nc[0]) * self.feats.shape[1] * 4)
        if self.labs is not None:
            tot += al(int(nc[0]) * 8)
        tot += al(3 * L * 4)
        for l in range(L):
            tot += al(self.nb * 4)
            if self.conv_r[l] > 0:
                tot += al(int(nc[l]) * w(l) * isz)
            if l + 1 < L and self.dls[l] > 0:
                tot += al(int(nc[l + 1]) * 12) + al(int(nc[l + 1]) * w(l) * isz) + al(int(nc[l]) * w(l + 1) * isz)
        return tot + 256

    def slab_bytes(self):
        if self.n_cap is not None:
            return self.static_slab_bytes()
        per_point = _SLAB_HINT.get((self.L, self.isz, self.cap), 2.6 * (2.2 * self.cap * self.isz + 12))
        return int(per_point * self.n0) + (1 << 16)

    def run(self, slab, stream_handle):
        """Returns True when done, False when it has to be repeated with a larger ``cap`` / slab (see slab_bytes)."""
        from . import _lib
        p = lambda a: a.ctypes.data if a is not None else None
        if self.n_cap is not None:
            rc = _lib.lib().kp_pyramid_build_static_dev(
                self.pts.data_ptr(), self.n0, p(self.lens), self.nb, self.L, p(self.conv_r), p(self.pool_r),
                p(self.up_r), p(self.dls), p(self.rot), p(self.limits), 1 if self.order == "reference" else 0,
                1 if self.isz == 8 else 0, self.cap, p(self.n_cap),
                self.feats.data_ptr() if self.feats is not None else None,
                self.feats.shape[1] if self.feats is not None else 0,
                self.labs.data_ptr() if self.labs is not None else None, self.label_pad, slab.data_ptr(), slab.numel(),
                p(self.offs), p(self.n_out), p(self.lens_out), p(self.widths), p(self.strides), p(self.need_bytes),
                p(self.need_cap), stream_handle)
            if rc == _lib.KP_ERR_CAPACITY:
                raise StaticCapacityExceeded(_lib.last_error())
            _lib.check(rc, "pyramid_build_static")
            return True
        rc = _lib.lib().kp_pyramid_build_dev(
            self.pts.data_ptr(), self.n0, p(self.lens), self.nb, self.L, p(self.conv_r), p(self.pool_r), p(self.up_r),
            p(self.dls), p(self.rot), p(self.limits), 1 if self.order == "reference" else 0, 1 if self.isz == 8 else 0,
            self.cap, slab.data_ptr(), slab.numel(), p(self.offs), p(self.n_out), p(self.lens_out), p(self.widths),
            p(self.strides), p(self.need_bytes), p(self.need_cap), stream_handle)
        if rc == _lib.KP_ERR_CAPACITY:
            need = int(self.need_bytes[0])
            if int(self.need_cap[0]) > self.cap:
                need = max(need, int(slab.numel() * int(self.need_cap[0]) / max(int(self.strides.max()), 1) * 1.1))
                self.cap = int(self.need_cap[0])
            _SLAB_HINT[(self.L, self.isz, self.cap)] = 1.05 * max(need, slab.numel()) / self.n0
            return False
        _lib.check(rc, "pyramid_build")
        _SLAB_HINT[(self.L, self.isz, self.cap)] = max(1.15 * int(self.need_bytes[0]) / self.n0,
                                                       _SLAB_HINT.get((self.L, self.isz, self.cap), 0.0) * 0.98)
        return True

    def no_crop(self):
        """True when no row of any matrix lost a neighbour to its width (limits / cap never bit)."""
        return bool((self.widths <= self.strides).all())

    def views(self, slab, mark_symmetric=None):
        """Tensor views of a built slab. Static layout: rows = capacities and full-stride widths (shapes never change
        from batch to batch); ``mark_symmetric`` overrides the per-batch no-crop test for the conv matrices."""
        L, nb, isz, dev = self.L, self.nb, self.isz, self.dev
        static = self.n_cap is not None

        def view(off, rows, cols, dtype, esz):
            return slab[off:off + rows * cols * esz].view(dtype).view(rows, cols)

        rows_of = (lambda l: int(self.n_cap[l])) if static else (lambda l: int(self.n_out[l]))
        P, Nn, Po, Up, Le = [], [], [], [], []
        for l in range(L):
            n = rows_of(l)
            P.append(self.pts if (l == 0 and not static) else view(int(self.offs[l]), n, 3, torch.float32, 4))
            for kind, lst in ((0, Nn), (1, Po), (2, Up)):
                o, w, sd = int(self.offs[(1 + kind) * L + l]), int(self.widths[kind * L + l]), int(self.strides[kind * L + l])
                if o < 0:
                    lst.append(torch.zeros((0, 1), dtype=self.dtype, device=dev))
                else:
                    rows = rows_of(l + 1) if kind == 1 else n
                    m = view(o, rows, sd, self.dtype, isz)
                    lst.append(m if static else m[:, :min(w, sd)])
            o = int(self.offs[4 * L + l])
            Le.append(slab[o:o + 4 * nb].view(torch.int32))
        # conv matrices of a layer searched against itself without a crop are symmetric (j in row i <=> i in row j, the
        # f32 distance is exactly symmetric): KPConv's backward can use the matrix itself as its transposed table
        for l in range(L):
            sym = (int(self.widths[l]) <= int(self.strides[l])) if mark_symmetric is None else mark_symmetric
            if Nn[l].shape[0] and sym:
                Nn[l]._kp_symmetric = True
        return P, Nn, Po, Up, Le

    def static_pool_widths(self, slab):
        """Per layer, the true width of the pool matrix as an int32 device scalar (view of the slab), for max_pool."""
        o = int(self.offs[5 * self.L + 2])
        w = slab[o:o + 12 * self.L].view(torch.int32)
        return [w[self.L + l:self.L + l + 1] for l in range(self.L)]

    def static_extras(self, slab):
        """(features [n_cap0, fdim], labels [n_cap0]) views of a static slab (None where not supplied)."""
        n = int(self.n_cap[0])
        fo, lo = int(self.offs[5 * self.L]), int(self.offs[5 * self.L + 1])
        f = slab[fo:fo + n * self.feats.shape[1] * 4].view(torch.float32).view(n, self.feats.shape[1]) if fo >= 0 else None
        lb = slab[lo:lo + n * 8].view(torch.int64) if lo >= 0 else None
        return f, lb


def build_native(points, stack_lengths, config, neighborhood_limits=None, random_grid_orient=True, order="reference",
                 index_dtype=torch.int64, cap=80, stream=None, rot=None):
    """The whole pyramid of one batch through ONE native call (kp_pyramid_build_dev): no Python between the ~150
    launches, the GIL released for its duration. ``points`` is a CUDA float32 [N,3] tensor. Returns the lists
    (points, neighbors, pools, upsamples, lengths) and the slab tensor all of them but ``points[0]`` are views of.
    ``stream``: the CUDA stream to run on (default: torch's current stream); the call returns after synchronising it.
    ``rot``: grid orientations from :func:`draw_grid_rotations` (default: drawn here)."""
    import ctypes as C
    nbld = NativeBuild(points, stack_lengths, config, neighborhood_limits, random_grid_orient, order, index_dtype, cap, rot)
    st = stream if stream is not None else torch.cuda.current_stream(nbld.dev)
    while True:
        with torch.cuda.stream(st):
            slab = torch.empty(nbld.slab_bytes(), dtype=torch.uint8, device=nbld.dev)
        if nbld.run(slab, C.c_void_p(st.cuda_stream)):
            break
    return (*nbld.views(slab), slab)


def segmentation_inputs(stacked_points, stacked_features, labels, stack_lengths, config, neighborhood_limits=None,
                        random_grid_orient=True, order="reference", index_dtype=torch.int64, device="cuda",
                        native=True):
    """Device counterpart of datasets/common.py:461-577: the flat list ``points*L + neighbors*L + pools*L +
    upsamples*L + lengths*L + [features, labels]`` of device tensors. ``native`` (default) walks the layers inside
    one C call (kp_pyramid_build_dev); ``native=False`` drives the per-operator entry points from Python, issuing all
    searches of the batch before the single synchronisation that reads their widths back. Same kernels, same results."""
    dev = torch.device(device)
    if native:
        def to_dev0(a, dtype=None):
            t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
            t = t.to(dev, non_blocking=True)
            return t.to(dtype) if dtype is not None else t
        P, Nn, Po, Up, Le, _ = build_native(to_dev0(stacked_points, torch.float32), stack_lengths, config,
                                         neighborhood_limits, random_grid_orient, order, index_dtype)
        feats = to_dev0(stacked_features, torch.float32) if stacked_features is not None else None
        labs = to_dev0(labels) if labels is not None else None
        return P + Nn + Po + Up + Le + [feats, labs]
    pb = PyramidBuilder(stacked_points, stack_lengths, config, neighborhood_limits, random_grid_orient, order,
                        index_dtype, dev)
    while pb.step():
        pass
    pb.resolve()

    def to_dev(a, dtype=None):
        t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
        t = t.to(dev, non_blocking=True)
        return t.to(dtype) if dtype is not None else t

    feats = to_dev(stacked_features, torch.float32) if stacked_features is not None else None
    labs = to_dev(labels) if labels is not None else None
    return pb.points + pb.neighbors + pb.pools + pb.upsamples + pb.lengths + [feats, labs]


class PyramidPrefetcher:
    """Builds the pyramids of upcoming batches on a side stream from a worker thread while the caller trains on the
    current one — the counterpart of the reference's DataLoader workers (train_*.py: ``num_workers=input_threads``),
    which run ``segmentation_inputs`` ahead of the training loop.

        pf.submit(points, features, labels, lengths)   # host (pinned) or device tensors / numpy arrays
        batch = pf.get()                               # DeviceBatch, complete on the device

    The worker runs nothing but ONE native call per batch (kp_pyramid_build_dev, GIL released): argument arrays,
    host->device copies and the output slab are prepared by ``submit`` and the tensor views by ``get``, both on the
    caller's thread, so the two threads hardly ever compete for the interpreter. Output slabs live in a ring of
    ``slots`` buffers; a slot is reused only after the consumer's stream has passed the step that read it: ``get()``
    marks the PREVIOUS batch as consumed (everything launched on the current stream so far), so launch a batch's work
    before asking for the next one and keep at most ``slots - 1`` batches alive.
    Grid orientations are drawn from ``np.random`` at submit time, in submission order."""

    def __init__(self, config, device="cuda", neighborhood_limits=None, random_grid_orient=True, order="reference",
                 index_dtype=torch.int64, slots=None, n_cap=None, plans=None, workers=1, sm_partition=0):
        """``n_cap``: per-layer row capacities -> batches come in the static layout of kp_pyramid_build_static_dev
        (features and labels inside the slab; ``batch.static_slab`` set) for :class:`weasal_b200.engine.GraphedTrainStep`;
        a batch that does not fit falls back to the ordinary layout."""
        import queue
        import sys
        import threading
        self.dev = torch.device(device)
        if self.dev.index is None:
            self.dev = torch.device("cuda", torch.cuda.current_device())
        self.cfg, self.limits, self.orient, self.order, self.dtype = config, neighborhood_limits, random_grid_orient, order, index_dtype
        self.n_cap = list(n_cap) if n_cap is not None else None
        # plans (weasal_b200.plan.ConvPlans, static layout only): the influence lists / transposed tables of every KPConv
        # of the network are built right behind the pyramid, by the same worker on the same stream
        self.plans = plans if n_cap is not None else None
        # ``workers`` build threads, each with its own side stream: batch t goes to worker t % workers, get() returns the
        # batches in submission order. One build is a chain of ~200 small kernels and a few host synchronisations (layer
        # sizes, search widths), ~3 ms of latency for ~2 ms of kernels: two builds in flight double the throughput of the
        # stage when the consumer submits two batches ahead.
        self.workers = max(1, int(workers))
        slots = slots if slots is not None else self.workers + 2
        self.plan_bufs = [None] * slots
        self.plan_jobs = [None] * slots      # (slab address, buffer address, ctypes job table) per ring slot
        # Stream priority (WEASAL_PREFETCH_PRIORITY, default 0 = low): with batches submitted two ahead nothing waits
        # for a build, so the builds run underneath the training step, whose graph is captured on a high-priority
        # stream (engine.GraphedTrainStep): its chain of ~400 small dependent kernels then gets SM slots as soon as a
        # node becomes ready instead of queueing behind the builds' CTAs. (With a single build in flight and the
        # consumer waiting for it, -1 is the better choice.)
        prio = int(os.environ.get("WEASAL_PREFETCH_PRIORITY", "0"))
        # WEASAL_PREFETCH_SMS=n (n > 0): the build streams are confined to a partition of >= n SMs (CUDA green context,
        # kp_sm_partition_streams), so that their CTAs never hold slots on the other SMs, where the training step runs
        n_sms = int(os.environ.get("WEASAL_PREFETCH_SMS", str(sm_partition)))
        self.sm_partition = 0
        self.sides = None
        if n_sms > 0:
            import ctypes as C
            from . import _lib
            handles, granted = (C.c_void_p * self.workers)(), C.c_int(0)
            with torch.cuda.device(self.dev):
                rc = _lib.lib().kp_sm_partition_streams(n_sms, self.workers, prio, handles, C.byref(granted))
            if rc == 0:
                self.sides = [torch.cuda.ExternalStream(int(h), device=self.dev) for h in handles]
                self.sm_partition = int(granted.value)
        if self.sides is None:
            self.sides = [torch.cuda.Stream(self.dev, priority=prio) for _ in range(self.workers)]
        self.side = self.sides[0]
        self.slabs = [None] * slots          # ring of output slabs (uint8 tensors allocated on the side stream)
        self.free_ev = [None] * slots        # recorded on the consumer's stream when a slot's batch has been consumed
        self.n_sub, self.n_got, self.last_slot = 0, 0, None
        self.stats = []  # per batch: (seconds the native call took in the worker, seconds get() waited for it)
        self.q_in = [queue.Queue() for _ in range(self.workers)]
        self.q_out = [queue.Queue() for _ in range(self.workers)]
        # the worker holds the GIL for microseconds per batch but must get it promptly when its call returns: with the
        # default 5 ms switch interval every hand-over from the launch-bound training thread would stall that long
        if sys.getswitchinterval() > 2e-4:
            sys.setswitchinterval(2e-4)
        self.threads = [threading.Thread(target=self._run, args=(w,), name=f"weasal-pyramid-{w}", daemon=True)
                        for w in range(self.workers)]
        for t in self.threads:
            t.start()

    def submit(self, points, features, labels, lengths, extras=None, inputs_ready=False):
        """``inputs_ready``: the caller guarantees that CUDA input tensors were completed long ago (e.g. a resident data
        set); otherwise the side stream first waits for everything queued so far on the caller's current stream, which
        may still be writing them (e.g. spheres just cut out of a cloud on the device)."""
        import ctypes as C
        lens = np.ascontiguousarray(lengths.cpu().numpy() if torch.is_tensor(lengths) else lengths, dtype=np.int32).reshape(-1)
        rot = draw_grid_rotations(self.cfg, len(lens), self.orient)

        def to_dev(a, dtype=None):
            if a is None:
                return None
            t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
            t = t.to(self.dev, non_blocking=True)
            return t.to(dtype) if dtype is not None and t.dtype != dtype else t

        slot = self.n_sub % len(self.slabs)
        w = self.n_sub % self.workers
        side = self.sides[w]
        self.n_sub += 1
        cuda_inputs = [t for t in (points, features, labels) if torch.is_tensor(t) and t.is_cuda]
        if cuda_inputs and not inputs_ready:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(self.dev))
            side.wait_event(ready)
        for t in cuda_inputs:  # allocated on the caller's stream, read on the side stream: tell the caching allocator
            t.record_stream(side)
        with torch.cuda.stream(side):
            if self.free_ev[slot] is not None:
                side.wait_event(self.free_ev[slot])  # the step that read this slot's previous batch has finished
            pts, feats, labs = to_dev(points, torch.float32), to_dev(features, torch.float32), to_dev(labels)
            nbld = NativeBuild(pts, lens, self.cfg, self.limits, self.orient, self.order, self.dtype, rot=rot,
                               n_cap=self.n_cap, features=feats, labels=labs)
            if self.slabs[slot] is None or self.slabs[slot].numel() < nbld.slab_bytes():
                self.slabs[slot] = torch.empty(int(nbld.slab_bytes() * 1.25), dtype=torch.uint8, device=self.dev)
            if self.plans is not None and self.plan_bufs[slot] is None:
                self.plan_bufs[slot] = torch.zeros(self.plans.nbytes, dtype=torch.uint8, device=self.dev)
        self.q_in[w].put((nbld, slot, (pts, feats, labs), extras, C.c_void_p(side.cuda_stream), w))

    def _run(self, w):
        torch.cuda.set_device(self.dev)
        while True:
            item = self.q_in[w].get()
            if item is None:
                return
            try:
                nbld, slot, owned, extras, sh, _ = item
                t0 = time.perf_counter()
                while True:
                    try:
                        if nbld.run(self.slabs[slot], sh):
                            break
                    except StaticCapacityExceeded:  # this batch goes out in the ordinary (dynamic) layout
                        nbld = NativeBuild(nbld.pts, nbld.lens, self.cfg, self.limits, self.orient, self.order,
                                           self.dtype, rot=nbld.rot)
                    # rare: grow the slab / neighbour capacity and repeat
                    if self.slabs[slot].numel() < nbld.slab_bytes():
                        with torch.cuda.stream(self.sides[w]):
                            self.slabs[slot] = torch.empty(int(nbld.slab_bytes() * 1.25), dtype=torch.uint8, device=self.dev)
                nbld.plans_ok = False
                if self.plans is not None and nbld.n_cap is not None and nbld.no_crop():
                    nbld.plans_ok = self._run_plans(nbld, slot, sh, w)
                nbld.build_s = time.perf_counter() - t0
                self.q_out[w].put((nbld, slot, owned, extras))
            except BaseException as e:  # surfaced by get()
                self.q_out[w].put((e, None, None, None))

    def _run_plans(self, nbld, slot, sh, w=0):
        """Worker thread: the lists of every KPConv for the batch just built into ring slot ``slot``. The job table holds
        device addresses only, all fixed for a (slab, buffer) pair, so it is built once per slot."""
        slab, buf = self.slabs[slot], self.plan_bufs[slot]
        key = (slab.data_ptr(), buf.data_ptr())
        if self.plan_jobs[slot] is None or self.plan_jobs[slot][0] != key:
            P, Nn, Po, Up, Le = nbld.views(slab)
            self.plan_jobs[slot] = (key, self.plans.jobs(P, Nn, Po, self.dtype == torch.int64, buf))
        self.plans.run(self.plan_jobs[slot][1], buf, sh)
        with torch.cuda.stream(self.sides[w]):
            overflow = int(buf[self.plans.flag_off:self.plans.flag_off + 4].view(torch.int32).item())  # syncs the stream
        return overflow == 0

    def get(self):
        t0 = time.perf_counter()
        nbld, slot, owned, extras = self.q_out[self.n_got % self.workers].get()
        self.n_got += 1
        if slot is None:
            raise nbld
        self.stats.append((nbld.build_s, time.perf_counter() - t0))
        cur = torch.cuda.current_stream(self.dev)
        if self.last_slot is not None:  # everything the caller launched for the previous batch precedes this event
            ev = torch.cuda.Event()
            ev.record(cur)
            self.free_ev[self.last_slot] = ev
        self.last_slot = slot
        slab = self.slabs[slot]
        for t in owned:  # allocated on the side stream, consumed on the caller's: tell the caching allocator
            if t is not None and t.is_cuda:
                t.record_stream(cur)
        slab.record_stream(cur)
        P, Nn, Po, Up, Le = nbld.views(slab)
        feats, labs = owned[1], owned[2]
        if nbld.n_cap is not None:
            feats, labs = nbld.static_extras(slab)
        batch = DeviceBatch(P + Nn + Po + Up + Le + [feats, labs], extras)
        batch.build, batch.no_crop = nbld, nbld.no_crop()
        batch.static_slab = slab[:nbld.static_slab_bytes()] if nbld.n_cap is not None else None
        if nbld.n_cap is not None:
            batch.pool_widths = nbld.static_pool_widths(slab)
        batch.n_points = int(nbld.n_out[0])
        batch.plan_buf = None
        if getattr(nbld, "plans_ok", False):
            batch.plan_buf = self.plan_bufs[slot]
            batch.plan_buf.record_stream(cur)
            self.plans.attach(batch.plan_buf, batch.neighbors, batch.pools)
        return batch

    def close(self):
        for q in self.q_in:
            q.put(None)
        for t in self.threads:
            t.join(timeout=10)


class DeviceBatch:
    """The fields of the reference's ``<DS>CustomBatch`` (Vaihingen3D_PseudoLabel.py:1407-1481) over device tensors."""

    def __init__(self, input_list, extras=None):
        L = (len(input_list) - 2) // 5
        ind = 0
        self.points = list(input_list[ind:ind + L]); ind += L
        self.neighbors = list(input_list[ind:ind + L]); ind += L
        self.pools = list(input_list[ind:ind + L]); ind += L
        self.upsamples = list(input_list[ind:ind + L]); ind += L
        self.lengths = list(input_list[ind:ind + L]); ind += L
        self.features = input_list[ind]; ind += 1
        self.labels = input_list[ind]
        for k, v in (extras or {}).items():
            setattr(self, k, v)

    def pin_memory(self):
        return self

    def to(self, device):
        return self
