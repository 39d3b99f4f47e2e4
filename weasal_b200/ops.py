"""Torch-tensor front ends of the device C-ABI (``*_dev`` entry points of include/weasal_b200.h).

PyTorch is plumbing here: it owns device memory and the current stream; all arithmetic happens in the hand-written
kernels of libweasal_b200.so. Tensors must live on a CUDA device; there is no CPU path.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("weasal_b200: tensors must be CUDA tensors (there is no CPU fallback)")


def _f32c(t):
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


def _lens(a):
    if torch.is_tensor(a):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.int32).reshape(-1)


# ------------------------------------------------------------------------------------------------------ radius search
def batch_query(queries, supports, q_batches, s_batches, radius, limit=None, dtype=torch.int64, cap_hint=96):
    """GPU ``batch_neighbors`` (datasets/common.py:185-196) on device tensors.

    Returns ``[Nq, min(Hmax, limit)]`` indices (``dtype`` int64 like the collated batch, or int32), sorted by
    (d2, index), padded with ``Ns``. ``limit`` is the ``neighborhood_limits`` crop of common.py:336-346.
    The result is a column-sliced view of the row buffer when Hmax < capacity (row stride != width).
    """
    _need_cuda(queries, supports)
    q, s = _f32c(queries), _f32c(supports)
    qb, sb = _lens(q_batches), _lens(s_batches)
    nq, ns = q.shape[0], s.shape[0]
    L = _lib.lib()
    cap = int(limit) if limit is not None else int(cap_hint)
    i64 = dtype == torch.int64
    while True:
        out = torch.empty((nq, max(cap, 1)), dtype=dtype, device=q.device)
        hmax = C.c_int(0)
        rc = L.kp_batch_query_dev(q.data_ptr(), nq, s.data_ptr(), ns, qb.ctypes.data, sb.ctypes.data, len(qb),
                                  float(radius), out.data_ptr(), 1 if i64 else 0, cap, C.byref(hmax), _stream())
        _lib.check(rc, "batch_query")
        if limit is not None or hmax.value <= cap:
            break
        cap = hmax.value
    width = hmax.value if limit is None else min(hmax.value, cap)
    return out[:, :width]


class SearchGrid:
    """Hash grid over a set of supports at one radius (kp_search_grid_build_dev), reusable by several searches."""

    def __init__(self, supports, s_batches, radius):
        _need_cuda(supports)
        self.s = _f32c(supports)
        self.sb = _lens(s_batches)
        self.radius = float(radius)
        L = _lib.lib()
        self.buf = torch.empty(int(L.kp_search_grid_bytes(self.s.shape[0], len(self.sb))), dtype=torch.uint8,
                               device=self.s.device)
        _lib.check(L.kp_search_grid_build_dev(self.s.data_ptr(), self.s.shape[0], self.sb.ctypes.data, len(self.sb),
                                              self.radius, self.buf.data_ptr(), _stream()), "search_grid_build")


class PendingSearches:
    """Radius searches issued without a host sync (kp_batch_query_dev_async). ``resolve()`` reads all their
    {Hmax, error} results with one device->host copy, re-runs the rare search whose rows outgrew its buffer, and
    returns the column-sliced index matrices in issue order."""

    def __init__(self, device, capacity=64):
        self.results = torch.zeros((capacity, 2), dtype=torch.int32, device=device)
        self.items = []

    def add(self, queries, supports, q_batches, s_batches, radius, limit=None, dtype=torch.int64, cap_hint=80,
            grid=None):
        """``grid``: a :class:`SearchGrid` built over ``supports`` at ``radius`` (skips the grid build)."""
        _need_cuda(queries, supports)
        q, s = _f32c(queries), _f32c(supports)
        qb, sb = _lens(q_batches), _lens(s_batches)
        nq, ns = q.shape[0], s.shape[0]
        cap = int(limit) if limit is not None else int(cap_hint)
        out = torch.empty((nq, max(cap, 1)), dtype=dtype, device=q.device)
        slot = len(self.items)
        if slot >= self.results.shape[0]:
            raise RuntimeError("PendingSearches: capacity exceeded")
        if grid is not None:
            rc = _lib.lib().kp_search_grid_query_dev(grid.buf.data_ptr(), ns, len(sb), grid.radius, q.data_ptr(), nq,
                                                     qb.ctypes.data, out.data_ptr(), 1 if dtype == torch.int64 else 0,
                                                     cap, None, self.results[slot].data_ptr(), _stream())
        else:
            rc = _lib.lib().kp_batch_query_dev_async(q.data_ptr(), nq, s.data_ptr(), ns, qb.ctypes.data, sb.ctypes.data,
                                                     len(qb), float(radius), out.data_ptr(),
                                                     1 if dtype == torch.int64 else 0, cap,
                                                     self.results[slot].data_ptr(), _stream())
        _lib.check(rc, "batch_query")
        self.items.append((out, cap, limit, (q, s, qb, sb, radius, dtype)))
        return slot

    def resolve(self, start=0):
        """Index matrices of the searches issued since slot ``start``."""
        if len(self.items) <= start:
            return []
        res = self.results[start:len(self.items)].cpu().numpy()  # the one synchronisation
        outs = []
        for k, (out, cap, limit, args) in enumerate(self.items[start:]):
            hmax, err = int(res[k, 0]), int(res[k, 1])
            if err & 1:
                raise RuntimeError("batch_query: cloud extent / radius exceeds 2^18 cells per axis")
            if err & 2:
                raise RuntimeError("batch_query: more than 1024 neighbours for one query")
            if (err & 4) or (limit is None and hmax > cap):
                # rare: a row outgrew the buffer (or the kernel's 256-hit staging): redo this one synchronously
                q, s, qb, sb, radius, dtype = args
                outs.append(batch_query(q, s, qb, sb, radius, limit=limit, dtype=dtype, cap_hint=max(hmax, cap)))
            else:
                outs.append(out[:, :min(hmax, cap)])
        return outs


# -------------------------------------------------------------------------------------------------- grid subsampling
def grid_subsample(points, batches, features=None, classes=None, sampleDl=0.1, max_p=0, order="reference", rot=None):
    """GPU ``batch_grid_subsampling`` core (datasets/common.py:77-182 minus the numpy rotation, which ``rot`` folds
    in). Returns (s_points, s_batches(np.int32)[, s_features][, s_classes]) as device tensors."""
    _need_cuda(points, features, classes)
    p = _f32c(points)
    b = _lens(batches)
    n = p.shape[0]
    f = _f32c(features) if features is not None else None
    c = classes.to(torch.int32).contiguous() if classes is not None else None
    fdim = f.shape[1] if f is not None else 0
    ldim = (c.shape[1] if c.dim() == 2 else 1) if c is not None else 0
    op = torch.empty((max(n, 1), 3), dtype=torch.float32, device=p.device)
    of = torch.empty((max(n, 1), fdim), dtype=torch.float32, device=p.device) if f is not None else None
    oc = torch.empty((max(n, 1), ldim), dtype=torch.int32, device=p.device) if c is not None else None
    ob = np.zeros(len(b), np.int32)
    m = C.c_int(0)
    r = np.ascontiguousarray(rot, np.float32) if rot is not None else None
    rc = _lib.lib().kp_grid_subsample_dev(p.data_ptr(), n, b.ctypes.data, len(b), f.data_ptr() if f is not None else None,
                                          fdim, c.data_ptr() if c is not None else None, ldim, float(sampleDl),
                                          int(max_p), 1 if order == "reference" else 0,
                                          r.ctypes.data if r is not None else None, op.data_ptr(), ob.ctypes.data,
                                          of.data_ptr() if of is not None else None,
                                          oc.data_ptr() if oc is not None else None, C.byref(m), _stream())
    _lib.check(rc, "grid_subsample")
    M = m.value
    res = [op[:M], ob]
    if f is not None:
        res.append(of[:M])
    if c is not None:
        res.append(oc[:M])
    return tuple(res)


# ------------------------------------------------------------------------------------------------------------ KPConv
def _idx_args(idx):
    if idx.dtype not in (torch.int32, torch.int64):
        idx = idx.long()
    if idx.dim() != 2:
        raise RuntimeError("neighb_inds must be [Nq, H]")
    if idx.shape[1] > 0 and idx.stride(1) != 1:
        idx = idx.contiguous()
    stride = idx.stride(0) if idx.shape[0] > 1 else max(idx.shape[1], 1)
    if stride < idx.shape[1]:
        idx = idx.contiguous()
        stride = idx.shape[1]
    return idx, 1 if idx.dtype == torch.int64 else 0, idx.shape[1], stride


def kpconv_impl():
    return os.environ.get("WEASAL_KPCONV_IMPL", "tc")


class KPConvFunction(torch.autograd.Function):
    """Autograd node around kp_kpconv_forward_dev / kp_kpconv_backward_dev (gradients w.r.t. x and weights only,
    like the reference: kernel_points has requires_grad=False and points carry no grad, blocks.py:235-236)."""

    @staticmethod
    def forward(ctx, q_pts, s_pts, neighb_inds, x, weights, kernel_points, KP_extent):
        _need_cuda(q_pts, s_pts, neighb_inds, x, weights, kernel_points)
        q, s, xx, w, kp = _f32c(q_pts), _f32c(s_pts), _f32c(x), _f32c(weights), _f32c(kernel_points)
        idx, i64, H, stride = _idx_args(neighb_inds)
        K, cin, cout = w.shape
        nq, ns = q.shape[0], s.shape[0]
        L = _lib.lib()
        lists = None
        plan = _find_plan(neighb_inds, kernel_points)
        if plan is not None and kpconv_impl() != "simt":
            # the geometry-only half of the operator (influence lists, transposed table) was built ahead of time by the
            # prefetch stage (weasal_b200.plan); the weights may come as ready-made operand images (WeightPacker)
            img = getattr(weights, "_kp_packed", None)
            out = torch.empty((nq, cout), dtype=torch.float32, device=q.device)
            wsrc = img["fwd"] if img is not None else w
            _lib.check(L.kp_kpconv_apply_lists_dev(nq, xx.data_ptr(), ns, cin, wsrc.data_ptr(), 1 if img is not None else 0, 0,
                                                   cout, K, plan.f_hdr.data_ptr(), plan.f_ent.data_ptr(), out.data_ptr(), 1.0,
                                                   _stream()), "kpconv_apply_lists")
            ctx.plan, ctx.img = plan, img
            ctx.save_for_backward(xx, w)
            ctx.shape = (nq, ns)
            return out
        ctx.plan = None
        if kpconv_impl() == "simt":  # fp32 cross-check path (CUDA-core gather + library GEMM), not the product path
            wf = torch.empty((nq, K * cin), dtype=torch.float32, device=q.device)
            _lib.check(L.kp_kpconv_wf_dev(q.data_ptr(), nq, s.data_ptr(), ns, idx.data_ptr(), i64, H, stride,
                                          xx.data_ptr(), cin, kp.data_ptr(), K, float(KP_extent), wf.data_ptr(),
                                          _stream()), "kpconv_wf")
            out = wf @ w.reshape(K * cin, cout)
        else:
            out = torch.empty((nq, cout), dtype=torch.float32, device=q.device)
            need_grad = ctx.needs_input_grad[3] or ctx.needs_input_grad[4]
            if need_grad and nq > 0 and ns > 0 and H > 0:
                # keep the influence entry lists for the backward pass (they depend on geometry only)
                kb, eb = C.c_longlong(0), C.c_longlong(0)
                L.kp_kpconv_lists_bytes(nq, H, C.byref(kb), C.byref(eb))
                lk = torch.empty(kb.value, dtype=torch.uint8, device=q.device)
                le = torch.empty(eb.value, dtype=torch.uint8, device=q.device)
                _lib.check(L.kp_kpconv_forward_keep_dev(q.data_ptr(), nq, s.data_ptr(), ns, idx.data_ptr(), i64, H,
                                                        stride, xx.data_ptr(), cin, w.data_ptr(), cout, kp.data_ptr(),
                                                        K, float(KP_extent), out.data_ptr(), lk.data_ptr(),
                                                        le.data_ptr(), _stream()), "kpconv_forward")
                lists = (lk, le)
            else:
                _lib.check(L.kp_kpconv_forward_dev(q.data_ptr(), nq, s.data_ptr(), ns, idx.data_ptr(), i64, H, stride,
                                                   xx.data_ptr(), cin, w.data_ptr(), cout, kp.data_ptr(), K,
                                                   float(KP_extent), out.data_ptr(), _stream()), "kpconv_forward")
        ctx.lists = lists
        ctx.idx_obj = neighb_inds  # the caller's tensor object: the transposed table is cached on it (see backward)
        ctx.save_for_backward(q, s, idx, xx, w, kp)
        ctx.meta = (i64, H, stride, float(KP_extent))
        return out

    @staticmethod
    def backward(ctx, d_out):
        if ctx.plan is not None:
            return KPConvFunction._backward_planned(ctx, d_out)
        q, s, idx, xx, w, kp = ctx.saved_tensors
        i64, H, stride, ext = ctx.meta
        K, cin, cout = w.shape
        nq, ns = q.shape[0], s.shape[0]
        do = _f32c(d_out)
        L = _lib.lib()
        if kpconv_impl() == "simt":
            wf = torch.empty((nq, K * cin), dtype=torch.float32, device=q.device)
            _lib.check(L.kp_kpconv_wf_dev(q.data_ptr(), nq, s.data_ptr(), ns, idx.data_ptr(), i64, H, stride,
                                          xx.data_ptr(), cin, kp.data_ptr(), K, ext, wf.data_ptr(), _stream()),
                       "kpconv_wf")
            dw = (wf.t() @ do).reshape(K, cin, cout)
            dwf = (do @ w.reshape(K * cin, cout).t()).contiguous()
            dx = torch.zeros((ns, cin), dtype=torch.float32, device=q.device)
            _lib.check(L.kp_kpconv_dx_atomic_dev(q.data_ptr(), nq, s.data_ptr(), ns, idx.data_ptr(), i64, H, stride,
                                                 dwf.data_ptr(), cin, kp.data_ptr(), K, ext, dx.data_ptr(), _stream()),
                       "kpconv_dx")
        else:
            dx = torch.empty((ns, cin), dtype=torch.float32, device=q.device)
            dw = torch.empty((K, cin, cout), dtype=torch.float32, device=q.device)
            lk, le = ctx.lists if ctx.lists is not None else (None, None)
            if getattr(ctx.idx_obj, "_kp_symmetric", False) and q.data_ptr() == s.data_ptr() and nq == ns:
                # the pyramid builder marked this conv matrix as its own transpose (queries == supports, no crop)
                _lib.check(L.kp_kpconv_backward_sym_dev(q.data_ptr(), nq, idx.data_ptr(), i64, H, stride, xx.data_ptr(),
                                                        cin, w.data_ptr(), cout, kp.data_ptr(), K, ext, do.data_ptr(),
                                                        dx.data_ptr(), dw.data_ptr(),
                                                        lk.data_ptr() if lk is not None else None,
                                                        le.data_ptr() if le is not None else None, _stream()),
                           "kpconv_backward")
                return None, None, None, dx, dw, None, None
            # The transposed neighbour table depends on the index matrix only: build it once per matrix and let every
            # KPConv that shares the matrix (the two blocks of a layer) reuse it.
            tr = getattr(ctx.idx_obj, "_kp_transposed", None)
            if tr is None or tr[2] != (idx.data_ptr(), nq, H, stride, ns):
                rowptr = torch.empty(ns + 2, dtype=torch.int32, device=q.device)
                col = torch.empty(max(nq * H, 1), dtype=torch.int32, device=q.device)
                _lib.check(L.kp_transpose_table_dev(idx.data_ptr(), i64, nq, H, stride, ns, rowptr.data_ptr(),
                                                    col.data_ptr(), _stream()), "transpose_table")
                tr = (rowptr, col, (idx.data_ptr(), nq, H, stride, ns))
                try:
                    ctx.idx_obj._kp_transposed = tr
                except Exception:
                    pass
            _lib.check(L.kp_kpconv_backward_kept_dev(q.data_ptr(), nq, s.data_ptr(), ns, idx.data_ptr(), i64, H, stride,
                                                     xx.data_ptr(), cin, w.data_ptr(), cout, kp.data_ptr(), K, ext,
                                                     do.data_ptr(), dx.data_ptr(), dw.data_ptr(),
                                                     lk.data_ptr() if lk is not None else None,
                                                     le.data_ptr() if le is not None else None, tr[0].data_ptr(),
                                                     tr[1].data_ptr(), _stream()),
                       "kpconv_backward")
        return None, None, None, dx, dw, None, None


def _backward_planned(ctx, d_out):
    """dW and dX over the prefetched lists: two independent kernels, forked onto two streams (capturable: the fork and
    the join are event dependencies) so that each one's tail overlaps the other's head."""
    xx, w = ctx.saved_tensors
    plan, img = ctx.plan, ctx.img
    nq, ns = ctx.shape
    K, cin, cout = w.shape
    do = _f32c(d_out)
    L = _lib.lib()
    dev = do.device
    dx = dw = None
    cur = torch.cuda.current_stream(dev)
    side = _side_stream(dev) if (ctx.needs_input_grad[3] and ctx.needs_input_grad[4] and _FORK_BACKWARD) else None
    if ctx.needs_input_grad[4]:
        dw = torch.empty((K, cin, cout), dtype=torch.float32, device=dev)
        if side is not None:
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                _lib.check(L.kp_kpconv_dw_lists_dev(nq, xx.data_ptr(), ns, cin, do.data_ptr(), cout, K, plan.f_hdr.data_ptr(),
                                                    plan.f_ent.data_ptr(), dw.data_ptr(), _stream()), "kpconv_dw_lists")
        else:
            _lib.check(L.kp_kpconv_dw_lists_dev(nq, xx.data_ptr(), ns, cin, do.data_ptr(), cout, K, plan.f_hdr.data_ptr(),
                                                plan.f_ent.data_ptr(), dw.data_ptr(), _stream()), "kpconv_dw_lists")
    if ctx.needs_input_grad[3]:
        dx = torch.empty((ns, cin), dtype=torch.float32, device=dev)
        wsrc = img["dx"] if img is not None else w
        _lib.check(L.kp_kpconv_apply_lists_dev(ns, do.data_ptr(), nq, cout, wsrc.data_ptr(), 1 if img is not None else 0, 1, cin,
                                               K, plan.d_hdr.data_ptr(), plan.d_ent.data_ptr(), dx.data_ptr(), 1.0, _stream()),
                   "kpconv_apply_lists")
    if side is not None:
        cur.wait_stream(side)
        for t in (xx, do, dw):
            t.record_stream(side)
    return None, None, None, dx, dw, None, None


KPConvFunction._backward_planned = staticmethod(_backward_planned)

_FORK_BACKWARD = os.environ.get("WEASAL_FORK_BACKWARD", "1") != "0"
_SIDE = {}
# True while a step engine runs its warm-up passes before a capture: code that forks onto side streams only inside a
# captured step must take the same streams during the warm-up, so that every (thread, stream) scratch arena of the
# library has its final size before recording starts (cudaMalloc is illegal during capture).
CAPTURE_WARMUP = [False]


def forking_for_capture():
    return CAPTURE_WARMUP[0] or torch.cuda.is_current_stream_capturing()


def _side_stream(dev, which=0):
    """``which``: 0 = the stream dW runs on beside dX, 1 = the shortcut branch of a residual block (net.ConvBlock)"""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device(), which)
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(dev, priority=int(os.environ.get("WEASAL_TRAIN_PRIORITY", "-1")))
    return _SIDE[key]


def _find_plan(neighb_inds, kernel_points):
    plans = getattr(neighb_inds, "_kp_plans", None)
    return plans.get(kernel_points.data_ptr()) if plans else None


def kpconv(q_pts, s_pts, neighb_inds, x, weights, kernel_points, KP_extent):
    return KPConvFunction.apply(q_pts, s_pts, neighb_inds, x, weights, kernel_points, KP_extent)


# ----------------------------------------------------------------------------------------------------------- pooling
class MaxPoolFunction(torch.autograd.Function):
    """models/blocks.py:93-112 as one gather-max kernel; backward routes each gradient to the winning support row."""

    @staticmethod
    def forward(ctx, x, inds, width=None):
        """``width``: optional int32 CUDA scalar tensor holding the matrix' true width (fixed-width matrices)."""
        _need_cuda(x, inds, width)
        xx = _f32c(x)
        idx, i64, H, stride = _idx_args(inds)
        ns, C_ = xx.shape
        nq = idx.shape[0]
        out = torch.empty((nq, C_), dtype=torch.float32, device=xx.device)
        arg = torch.empty((nq, C_), dtype=torch.int32, device=xx.device)
        if width is not None:
            if width.dtype != torch.int32:
                raise RuntimeError("max_pool: width must be an int32 tensor")
            _lib.check(_lib.lib().kp_max_pool_forward_width_dev(xx.data_ptr(), ns, C_, idx.data_ptr(), i64, nq, H, stride,
                                                                width.data_ptr(), out.data_ptr(), arg.data_ptr(),
                                                                _stream()), "max_pool")
        else:
            _lib.check(_lib.lib().kp_max_pool_forward_dev(xx.data_ptr(), ns, C_, idx.data_ptr(), i64, nq, H, stride,
                                                          out.data_ptr(), arg.data_ptr(), _stream()), "max_pool")
        ctx.save_for_backward(arg)
        ctx.ns = ns
        return out

    @staticmethod
    def backward(ctx, d_out):
        (arg,) = ctx.saved_tensors
        do = _f32c(d_out)
        nq, C_ = arg.shape
        dx = torch.empty((ctx.ns, C_), dtype=torch.float32, device=do.device)
        _lib.check(_lib.lib().kp_max_pool_backward_dev(do.data_ptr(), arg.data_ptr(), nq, C_, dx.data_ptr(), ctx.ns,
                                                       _stream()), "max_pool_backward")
        return dx, None, None


class ClosestPoolFunction(torch.autograd.Function):
    """models/blocks.py:77-90: nearest upsampling through the first neighbour column."""

    @staticmethod
    def forward(ctx, x, inds):
        _need_cuda(x, inds)
        xx = _f32c(x)
        idx, i64, H, stride = _idx_args(inds)
        if H == 0:  # the reference indexes inds[:, 0] (blocks.py:87): an index error there, an error here
            raise IndexError("closest_pool: the index matrix has no columns")
        ns, C_ = xx.shape
        nq = idx.shape[0]
        out = torch.empty((nq, C_), dtype=torch.float32, device=xx.device)
        _lib.check(_lib.lib().kp_closest_pool_dev(xx.data_ptr(), ns, C_, idx.data_ptr(), i64, nq, stride,
                                                  out.data_ptr(), 0, _stream()), "closest_pool")
        ctx.save_for_backward(idx)
        ctx.meta = (ns, i64, stride)
        return out

    @staticmethod
    def backward(ctx, d_out):
        (idx,) = ctx.saved_tensors
        ns, i64, stride = ctx.meta
        # the gradient of a torch.cat input is a column slice of a wider matrix: read it in place (row stride) instead
        # of copying it (the copy was 21 us per decoder level)
        if d_out.dtype == torch.float32 and d_out.dim() == 2 and d_out.stride(1) == 1 and d_out.stride(0) >= d_out.shape[1]:
            do, ld = d_out, d_out.stride(0)
        else:
            do = _f32c(d_out)
            ld = do.shape[1]
        nq, C_ = do.shape
        dx = torch.empty((ns, C_), dtype=torch.float32, device=do.device)
        _lib.check(_lib.lib().kp_closest_pool_strided_dev(do.data_ptr(), ld, ns, C_, idx.data_ptr(), i64, nq, stride,
                                                          dx.data_ptr(), 1, _stream()), "closest_pool_backward")
        return dx, None


def max_pool(x, inds, width=None):
    return MaxPoolFunction.apply(x, inds, width)


def closest_pool(x, inds):
    return ClosestPoolFunction.apply(x, inds)


# ------------------------------------------------------------------------------------------------------ unary blocks
class LinearActFunction(torch.autograd.Function):
    """``leaky_relu(x @ weight.T + bias, slope)`` (models/blocks.py:467-507 UnaryBlock on 2-D features) as one tcgen05
    kernel; backward = one kernel for dX and one for dW, the LeakyReLU derivative applied while loading dY."""

    @staticmethod
    def forward(ctx, x, weight, bias, slope):
        _need_cuda(x, weight, bias)
        xx, w = _f32c(x), _f32c(weight)
        b = _f32c(bias) if bias is not None else None
        n, cin = xx.shape
        cout = w.shape[0]
        y = torch.empty((n, cout), dtype=torch.float32, device=xx.device)
        img = getattr(weight, "_kp_packed", None)
        if img is not None:
            _lib.check(_lib.lib().kp_linear_forward_packed_dev(xx.data_ptr(), n, cin, img["fwd"].data_ptr(),
                                                               b.data_ptr() if b is not None else None, cout, float(slope),
                                                               y.data_ptr(), _stream()), "linear_forward")
        else:
            _lib.check(_lib.lib().kp_linear_forward_dev(xx.data_ptr(), n, cin, w.data_ptr(), b.data_ptr() if b is not None else None,
                                                        cout, float(slope), y.data_ptr(), _stream()), "linear_forward")
        ctx.img = img
        ctx.save_for_backward(xx, w, y)
        ctx.slope, ctx.has_bias = float(slope), bias is not None
        return y

    @staticmethod
    def backward(ctx, d_y):
        xx, w, y = ctx.saved_tensors
        n, cin = xx.shape
        cout = w.shape[0]
        g = _f32c(d_y)
        act = ctx.slope != 1.0
        db = None
        if ctx.has_bias:  # head layers only: materialise g once, its column sums are the bias gradient
            if act:
                g = torch.where(y > 0, g, g * ctx.slope)
                act = False
            db = g.sum(0)
        dx = torch.empty((n, cin), dtype=torch.float32, device=xx.device) if ctx.needs_input_grad[0] else None
        dw = torch.empty((cout, cin), dtype=torch.float32, device=xx.device)
        if ctx.img is not None:
            # dW and dX as two kernels on two streams (see _backward_planned), dX from the ready-made transposed images
            L = _lib.lib()
            yp = y.data_ptr() if act else None
            sl = ctx.slope if act else 1.0
            dev = g.device
            cur = torch.cuda.current_stream(dev)
            side = _side_stream(dev) if (dx is not None and _FORK_BACKWARD) else None
            if side is not None:
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    _lib.check(L.kp_linear_dw_dev(xx.data_ptr(), n, cin, cout, yp, sl, g.data_ptr(), dw.data_ptr(), _stream()),
                               "linear_dw")
            else:
                _lib.check(L.kp_linear_dw_dev(xx.data_ptr(), n, cin, cout, yp, sl, g.data_ptr(), dw.data_ptr(), _stream()),
                           "linear_dw")
            if dx is not None:
                _lib.check(L.kp_linear_dx_packed_dev(n, cin, ctx.img["dx"].data_ptr(), cout, yp, sl, g.data_ptr(), dx.data_ptr(),
                                                     _stream()), "linear_dx")
            if side is not None:
                cur.wait_stream(side)
                for t in (xx, y, g, dw):
                    t.record_stream(side)
            return dx, dw, db, None
        _lib.check(_lib.lib().kp_linear_backward_dev(xx.data_ptr(), n, cin, w.data_ptr(), cout,
                                                     y.data_ptr() if act else None, ctx.slope if act else 1.0,
                                                     g.data_ptr(), dx.data_ptr() if dx is not None else None,
                                                     dw.data_ptr(), _stream()), "linear_backward")
        return dx, dw, db, None


def linear_act(x, weight, bias=None, negative_slope=1.0):
    return LinearActFunction.apply(x, weight, bias, negative_slope)
