"""Geometry-only and weight-only halves of the KPConv operator, taken off the training stream.

``KPConv.forward`` (models/blocks.py:238-374) recomputes, at every call, things that do not depend on the features:
  * the kernel-point influence weights of every (query, neighbour) pair (blocks.py:281-338) — a function of the batch
    geometry and of ``kernel_points``, which the reference freezes (``requires_grad=False``, blocks.py:235-236);
  * for the backward pass, the transposed neighbour relation (autograd's ``scatter_add_`` over the index matrix);
and the B200 path adds a third feature-independent step, the TF32 operand images of the weights.

A training loop knows the geometry one step ahead (the pyramid of batch t+1 is built while batch t trains), so
:class:`ConvPlans` builds the influence lists of EVERY KPConv of a network for a batch — forward lists, dX lists, the
transposed tables of the strided layers — with one native call (``kp_kpconv_prepare_dev``) issued by the prefetch thread
on its side stream, into a fixed-layout buffer next to the static pyramid slab. The lists are attached to the batch's index
tensors (``_kp_plans``), where ``ops.KPConvFunction`` finds them: forward, dW and dX then start at the tensor-core
kernels. :class:`WeightPacker` packs the operand images of all layers (KPConv W and W^T, unary blocks) with ONE launch
per step instead of one per operator call.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

TILE = 128
HDR_INTS_PER_TILE = 272
CTL_INTS = 4


class ConvSpec:
    """One KPConv call site of a network: which pyramid matrices it reads and with which kernel points."""

    def __init__(self, layer, strided, module):
        self.layer, self.strided, self.module = int(layer), bool(strided), module
        self.kp = module.kernel_points          # frozen parameter: its storage is the plan's key
        self.extent = float(module.KP_extent)
        self.K = int(module.K)


def conv_specs(net):
    """KPConv call sites of a network, in module order: the harness blocks (``net.ConvBlock``: ``.conv``, ``.layer``,
    ``.strided``) and the reference's blocks after ``dropin.install()`` (``.KPConv``, ``.layer_ind``, ``.block_name``,
    models/blocks.py:510-565, 625-709)."""
    from .kpconv import KPConv
    specs = []
    for m in net.modules():
        conv = getattr(m, "conv", None) if not isinstance(m, KPConv) else None
        if isinstance(conv, KPConv) and hasattr(m, "layer"):
            specs.append(ConvSpec(m.layer, getattr(m, "strided", False), conv))
            continue
        conv = getattr(m, "KPConv", None) if not isinstance(m, KPConv) else None
        if isinstance(conv, KPConv) and hasattr(m, "layer_ind"):
            specs.append(ConvSpec(m.layer_ind, "strided" in getattr(m, "block_name", ""), conv))
    return specs


def _al(b, a=256):
    return (int(b) + a - 1) // a * a


class Plan:
    """Lists of one KPConv for one batch (views of a plan buffer)."""
    __slots__ = ("f_hdr", "f_ent", "d_hdr", "d_ent")

    def __init__(self, f_hdr, f_ent, d_hdr, d_ent):
        self.f_hdr, self.f_ent, self.d_hdr, self.d_ent = f_hdr, f_ent, d_hdr, d_ent


class ConvPlans:
    """Fixed layout of all lists of a network over the STATIC pyramid layout (rows = per-layer capacities, matrix widths =
    the calibrated limits), the job table that builds them, and the views that hand them to the operators.

    ``entry_caps[i]``: capacity of conv i's lists in records (forward and dX lists hold exactly the same number of
    records: one per (query, neighbour, kernel point) triple with a non-zero weight)."""

    def __init__(self, specs, n_cap, conv_widths, pool_widths, entry_caps, forward_only=False):
        self.forward_only = bool(forward_only)  # inference: no dX lists, no transposed tables
        self.specs, self.n_cap = list(specs), [int(v) for v in n_cap]
        self.conv_w, self.pool_w = [int(v) for v in conv_widths], [int(v) for v in pool_widths]
        self.caps = [int(v) for v in entry_caps]
        off = 0
        self.flag_off = off
        off += 256
        self.items = []
        self.tr = {}          # layer -> (rowptr offset, col offset) of the transposed pool table
        for sp, cap in zip(self.specs, self.caps):
            l = sp.layer
            nq = self.n_cap[l + 1] if sp.strided else self.n_cap[l]
            ns = self.n_cap[l]
            H = self.pool_w[l] if sp.strided else self.conv_w[l]
            it = {"nq": nq, "ns": ns, "H": H, "cap": cap}
            for side, nc in (("f", nq), ("d", ns)):
                tiles = -(-nc // TILE)
                it[side + "_hdr"] = off
                it[side + "_hdr_bytes"] = (CTL_INTS + tiles * HDR_INTS_PER_TILE) * 4
                off += _al(it[side + "_hdr_bytes"])
                it[side + "_ent"] = off
                it[side + "_ent_bytes"] = (cap + 2) * 8
                off += _al(it[side + "_ent_bytes"])
            if sp.strided and l not in self.tr:
                rp = off
                off += _al((ns + 2) * 4)
                col = off
                off += _al(nq * H * 4)
                self.tr[l] = (rp, col)
            self.items.append(it)
        self.nbytes = off

    # ------------------------------------------------------------------------------------------------ job table
    def jobs(self, points, neighbors, pools, index_is_i64, buf):
        """ctypes job array for one (pyramid slab, plan buffer) pair; addresses are fixed for a ring slot, so the table
        is built once per slot. ``points`` / ``neighbors`` / ``pools``: the static views of the pyramid slab."""
        base = buf.data_ptr()
        jobs = []
        done_tr = set()
        for sp, it in zip(self.specs, self.items):
            l = sp.layer
            q = points[l + 1] if sp.strided else points[l]
            s = points[l]
            idx = pools[l] if sp.strided else neighbors[l]
            stride = idx.stride(0)
            kp = sp.kp
            common = dict(kernel_points=kp.data_ptr(), K=sp.K, KP_extent=sp.extent, idx_is_i64=1 if index_is_i64 else 0,
                          H=it["H"], idx_stride=stride, entries_cap=it["cap"])
            jobs.append(_lib.ListJob(kind=0, centres=q.data_ptr(), nc=it["nq"], others=s.data_ptr(), no=it["ns"],
                                     neighb_inds=idx.data_ptr(), kp_sign=1.0, hdr=base + it["f_hdr"],
                                     entries=base + it["f_ent"], **common))
            if self.forward_only:
                continue
            if not sp.strided:
                # conv matrix of a layer against itself, no crop: its own transpose (the f32 distance is symmetric)
                jobs.append(_lib.ListJob(kind=0, centres=s.data_ptr(), nc=it["ns"], others=q.data_ptr(), no=it["nq"],
                                         neighb_inds=idx.data_ptr(), kp_sign=-1.0, hdr=base + it["d_hdr"],
                                         entries=base + it["d_ent"], **common))
            else:
                rp, col = self.tr[l]
                if l not in done_tr:
                    done_tr.add(l)
                    jobs.append(_lib.ListJob(kind=1, nc=it["nq"], no=it["ns"], neighb_inds=idx.data_ptr(),
                                             idx_is_i64=1 if index_is_i64 else 0, H=it["H"], idx_stride=stride,
                                             rowptr=base + rp, col=base + col))
                jobs.append(_lib.ListJob(kind=2, centres=s.data_ptr(), nc=it["ns"], others=q.data_ptr(), no=it["nq"],
                                         rowptr=base + rp, col=base + col, n_pairs=it["nq"] * it["H"],
                                         kernel_points=kp.data_ptr(), K=sp.K, kp_sign=-1.0, KP_extent=sp.extent,
                                         hdr=base + it["d_hdr"], entries=base + it["d_ent"], entries_cap=it["cap"]))
        arr = (_lib.ListJob * len(jobs))(*jobs)
        return arr

    def run(self, jobs, buf, stream_handle):
        """Issue the whole table on ``stream_handle`` (no synchronisation; the overflow flag is the first int of buf)."""
        _lib.check(_lib.lib().kp_kpconv_prepare_dev(C.cast(jobs, C.c_void_p), len(jobs), buf.data_ptr() + self.flag_off,
                                                    stream_handle), "kpconv_prepare")

    # ---------------------------------------------------------------------------------------------------- views
    def attach(self, buf, neighbors, pools):
        """Hang the plans on the index tensors the KPConv calls will receive (``_kp_plans``: kernel-point storage ->
        Plan), as views of ``buf``."""
        for sp, it in zip(self.specs, self.items):
            idx = pools[sp.layer] if sp.strided else neighbors[sp.layer]
            v = lambda o, n: buf[o:o + n]
            plan = Plan(v(it["f_hdr"], it["f_hdr_bytes"]), v(it["f_ent"], it["f_ent_bytes"]),
                        v(it["d_hdr"], it["d_hdr_bytes"]), v(it["d_ent"], it["d_ent_bytes"]))
            d = getattr(idx, "_kp_plans", None)
            if d is None:
                d = {}
                idx._kp_plans = d
            d[sp.kp.data_ptr()] = plan


def measure_entries(specs, batch):
    """Records each conv's forward lists need for ``batch`` (a DeviceBatch of any layout), by building them once into a
    worst-case buffer: the calibration pass of the entry capacities."""
    L = _lib.lib()
    out = []
    cache = {}
    for sp in specs:
        l = sp.layer
        q = batch.points[l + 1] if sp.strided else batch.points[l]
        s = batch.points[l]
        idx = batch.pools[l] if sp.strided else batch.neighbors[l]
        key = (l, sp.strided, sp.kp.data_ptr())
        if key in cache:
            out.append(cache[key])
            continue
        nq, H = idx.shape[0], idx.shape[1]
        if nq == 0 or H == 0:
            out.append(0)
            continue
        kb, eb = C.c_longlong(0), C.c_longlong(0)
        L.kp_kpconv_lists_bytes(nq, H, C.byref(kb), C.byref(eb))
        hdr = torch.empty(kb.value, dtype=torch.uint8, device=q.device)
        ent = torch.empty(eb.value, dtype=torch.uint8, device=q.device)
        _lib.check(L.kp_kpconv_lists_build_dev(q.data_ptr(), nq, s.data_ptr(), s.shape[0], idx.data_ptr(),
                                               1 if idx.dtype == torch.int64 else 0, H, idx.stride(0), None, None, 0,
                                               sp.kp.data_ptr(), sp.K, 1.0, sp.extent, hdr.data_ptr(), ent.data_ptr(),
                                               (eb.value - 16) // 8, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "kpconv_lists_build")
        used = int(hdr[:4].view(torch.int32)[0].item())
        cache[key] = used
        out.append(used)
    return out


class WeightPacker:
    """Operand images of every tensor-core contraction of a network, packed by ONE launch (``kp_pack_weights_dev``).

        packer = WeightPacker(net)      # finds KPConv modules and the unary blocks that run on the fused linear kernels
        packer.pack()                   # after every optimizer step (the captured training step starts with it)

    The images are views of one flat buffer, published on the parameters as ``param._kp_packed = {"fwd": ..., "dx": ...}``;
    ``ops.KPConvFunction`` / ``ops.LinearActFunction`` use them when present. A caller that changes weights without
    calling :meth:`pack` (e.g. ``load_state_dict``) must call it, or :meth:`release` the images."""

    def __init__(self, net, linear_weights=()):
        from .kpconv import KPConv
        L = _lib.lib()
        jobs = []   # (kind, param, K, cin, cout)
        for m in net.modules():
            if isinstance(m, KPConv):
                K, cin, cout = m.weights.shape
                jobs.append((0, m.weights, K, cin, cout))
                jobs.append((1, m.weights, K, cin, cout))
        for w in linear_weights:
            cout, cin = w.shape
            jobs.append((2, w, 1, cin, cout))
            jobs.append((3, w, 1, cin, cout))
        self.jobs = jobs
        sizes = [int(L.kp_pack_image_floats(k, K, cin, cout)) for k, _, K, cin, cout in jobs]
        if any(sz < 0 for sz in sizes):
            raise RuntimeError("WeightPacker: a layer is too wide for the tensor-core path")
        offs = np.concatenate([[0], np.cumsum([_al(4 * sz) // 4 for sz in sizes])]).astype(np.int64)
        dev = jobs[0][1].device if jobs else torch.device("cuda")
        self.buf = torch.zeros(int(offs[-1]) if len(jobs) else 1, dtype=torch.float32, device=dev)
        self.views = [self.buf[int(offs[i]):int(offs[i]) + sizes[i]] for i in range(len(jobs))]
        n = len(jobs)
        self._kinds = (C.c_int * n)(*[j[0] for j in jobs])
        self._Ks = (C.c_int * n)(*[j[2] for j in jobs])
        self._cins = (C.c_int * n)(*[j[3] for j in jobs])
        self._couts = (C.c_int * n)(*[j[4] for j in jobs])
        self._w = (C.c_void_p * n)(*[j[1].data_ptr() for j in jobs])
        self._img = (C.c_void_p * n)(*[v.data_ptr() for v in self.views])
        for (kind, p, *_), v in zip(jobs, self.views):
            d = getattr(p, "_kp_packed", None)
            if d is None:
                d = {}
                p._kp_packed = d
            d["fwd" if kind in (0, 2) else "dx"] = v

    def pack(self):
        if not self.jobs:
            return
        for i, j in enumerate(self.jobs):   # (parameters keep their storage; checked because a stale pointer is silent)
            if j[1].data_ptr() != self._w[i]:
                self._w[i] = j[1].data_ptr()
        _lib.check(_lib.lib().kp_pack_weights_dev(len(self.jobs), self._kinds, self._w, self._Ks, self._cins, self._couts,
                                                  self._img, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "pack_weights")

    def release(self):
        for _, p, *_ in self.jobs:
            if hasattr(p, "_kp_packed"):
                del p._kp_packed
