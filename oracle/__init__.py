"""TEST INFRASTRUCTURE — not product code.

ctypes bindings to the CPU oracle:
  * ``liboracle.so``  — plain-C restatement of the hot path (oracle/kp_oracle.c, each function cites the
    reference file:line it follows) plus the live-``std::unordered_map`` probe (oracle/umap_probe.cpp);
  * ``_ref/libweasal_ref.so`` — the UNMODIFIED reference C++ cores compiled from /root/reference behind
    oracle/ref_shim.cpp (built in the build container, travels to the GPU box as a prebuilt file).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs may import
this package. Nothing under ``weasal_b200/`` does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
_ref = None
_sched_n = 0

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int)
_i64p = C.POINTER(C.c_int64)
_u64p = C.POINTER(C.c_uint64)


def build():
    """Compile the oracle (and oracle/_ref when the reference sources are present)."""
    subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _lib = C.CDLL(path)
        _lib.orc_batch_neighbors.restype = C.c_int
        _lib.orc_grid_subsample.restype = C.c_int
        _lib.orc_grid_subsample_batch.restype = C.c_int
        _lib.umap_rehash_schedule.restype = C.c_int
        _ensure_schedule(1 << 16)
    return _lib


def _ensure_schedule(max_n):
    """Probe the live unordered_map up to ``max_n`` elements and hand the schedule to the C model."""
    global _sched_n
    if max_n <= _sched_n:
        return
    n = 1 << 16
    while n < max_n:
        n *= 2
    elt = np.zeros(64, np.int64)
    bkt = np.zeros(64, np.int64)
    cnt = _lib.umap_rehash_schedule(C.c_int64(n), _ptr(elt, _i64p), _ptr(bkt, _i64p), 64)
    _lib.orc_set_rehash_schedule(_ptr(elt, _i64p), _ptr(bkt, _i64p), cnt)
    _sched_n = n


def rehash_schedule(max_n):
    lib()
    elt = np.zeros(64, np.int64)
    bkt = np.zeros(64, np.int64)
    cnt = _lib.umap_rehash_schedule(C.c_int64(max_n), _ptr(elt, _i64p), _ptr(bkt, _i64p), 64)
    return elt[:cnt].copy(), bkt[:cnt].copy()


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libweasal_ref.so"))


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(os.path.join(_HERE, "_ref", "libweasal_ref.so"))
    return _ref


def _take(ptr, shape, dtype, free):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return np.zeros(shape, dtype)
    arr = np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True).reshape(shape)
    free(ptr)
    return arr


# ---------------------------------------------------------------------------------------------- restatement
def batch_neighbors(queries, supports, q_batches, s_batches, radius):
    """(d2, index)-ordered radius search; int32 [Nq, Hmax] padded with Ns (neighbors.cpp:125-208)."""
    L = lib()
    q, s, qb, sb = _f32(queries), _f32(supports), _i32(q_batches), _i32(s_batches)
    out = _i32p()
    hmax = L.orc_batch_neighbors(_ptr(q, _f32p), len(q), _ptr(s, _f32p), len(s), _ptr(qb, _i32p), _ptr(sb, _i32p),
                                 len(qb), C.c_float(radius), C.byref(out))
    return _take(out, (len(q), hmax), np.int32, L.orc_free)


def umap_order(keys):
    """C model of the libstdc++ unordered_map iteration order after inserting distinct ``keys`` in order."""
    L = lib()
    k = np.ascontiguousarray(keys, dtype=np.uint64)
    _ensure_schedule(len(k) + 1)
    order = np.zeros(len(k), np.int32)
    L.orc_umap_order(_ptr(k, _u64p), len(k), _ptr(order, _i32p))
    return order


def umap_live_order(keys):
    """Iteration order of a real std::unordered_map<size_t,int> (pins :func:`umap_order`)."""
    L = lib()
    k = np.ascontiguousarray(keys, dtype=np.uint64)
    order = np.zeros(len(k), np.int32)
    L.umap_live_order(_ptr(k, _u64p), len(k), _ptr(order, _i32p))
    return order


def grid_subsample(points, features=None, classes=None, sampleDl=0.1, order="reference", return_keys=False):
    """grid_subsampling.cpp:5-106. Returns (points[, features][, classes][, keys, first_idx])."""
    L = lib()
    p = _f32(points)
    _ensure_schedule(len(p) + 1)
    f = _f32(features) if features is not None else None
    c = _i32(classes) if classes is not None else None
    fdim = f.shape[1] if f is not None else 0
    ldim = (c.shape[1] if c.ndim == 2 else 1) if c is not None else 0
    op, of, oc, ok, ofi = _f32p(), _f32p(), _i32p(), _u64p(), _i32p()
    m = L.orc_grid_subsample(_ptr(p, _f32p), len(p), _ptr(f, _f32p), fdim, _ptr(c, _i32p), ldim, C.c_float(sampleDl),
                             1 if order == "reference" else 0, C.byref(op), C.byref(of) if f is not None else None,
                             C.byref(oc) if c is not None else None, C.byref(ok), C.byref(ofi))
    res = [_take(op, (m, 3), np.float32, L.orc_free)]
    if f is not None:
        res.append(_take(of, (m, fdim), np.float32, L.orc_free))
    if c is not None:
        res.append(_take(oc, (m, ldim), np.int32, L.orc_free))
    keys = _take(ok, (m,), np.uint64, L.orc_free)
    first = _take(ofi, (m,), np.int32, L.orc_free)
    if return_keys:
        res += [keys, first]
    return res[0] if len(res) == 1 else tuple(res)


def grid_subsample_batch(points, batches, features=None, classes=None, sampleDl=0.1, max_p=0, order="reference"):
    """grid_subsampling.cpp:109-211. Returns (points, batches[, features][, classes])."""
    L = lib()
    p, b = _f32(points), _i32(batches)
    _ensure_schedule(len(p) + 1)
    f = _f32(features) if features is not None else None
    c = _i32(classes) if classes is not None else None
    fdim = f.shape[1] if f is not None else 0
    ldim = (c.shape[1] if c.ndim == 2 else 1) if c is not None else 0
    op, of, oc = _f32p(), _f32p(), _i32p()
    ob = np.zeros(len(b), np.int32)
    m = L.orc_grid_subsample_batch(_ptr(p, _f32p), len(p), _ptr(b, _i32p), len(b), _ptr(f, _f32p), fdim,
                                   _ptr(c, _i32p), ldim, C.c_float(sampleDl), int(max_p),
                                   1 if order == "reference" else 0, C.byref(op), _ptr(ob, _i32p),
                                   C.byref(of) if f is not None else None, C.byref(oc) if c is not None else None)
    res = [_take(op, (m, 3), np.float32, L.orc_free), ob]
    if f is not None:
        res.append(_take(of, (m, fdim), np.float32, L.orc_free))
    if c is not None:
        res.append(_take(oc, (m, ldim), np.int32, L.orc_free))
    return tuple(res)


def rotate(points, R, transpose=False):
    """datasets/common.py:118 / :134 (f32, products summed left to right)."""
    L = lib()
    p, r = _f32(points), _f32(R)
    out = np.empty_like(p)
    L.orc_rotate(_ptr(p, _f32p), len(p), _ptr(r, _f32p), 1 if transpose else 0, _ptr(out, _f32p))
    return out


def kpconv_forward(q_pts, s_pts, neighb_inds, x, weights, kernel_points, KP_extent):
    """models/blocks.py:277-374 (rigid, linear, sum); f64 accumulation."""
    L = lib()
    q, s, xx, w, kp = _f32(q_pts), _f32(s_pts), _f32(x), _f32(weights), _f32(kernel_points)
    idx = np.ascontiguousarray(neighb_inds, dtype=np.int64)
    K, cin, cout = w.shape
    out = np.zeros((len(q), cout), np.float32)
    L.orc_kpconv_forward(_ptr(q, _f32p), len(q), _ptr(s, _f32p), len(s), _ptr(idx, _i64p), idx.shape[1],
                         _ptr(xx, _f32p), cin, _ptr(w, _f32p), cout, _ptr(kp, _f32p), K, C.c_float(KP_extent),
                         _ptr(out, _f32p))
    return out


def kpconv_backward(q_pts, s_pts, neighb_inds, x, weights, kernel_points, KP_extent, d_out):
    """Analytic gradient of :func:`kpconv_forward` w.r.t. ``x`` and ``weights``; f64 accumulation."""
    L = lib()
    q, s, xx, w, kp = _f32(q_pts), _f32(s_pts), _f32(x), _f32(weights), _f32(kernel_points)
    do = _f32(d_out)
    idx = np.ascontiguousarray(neighb_inds, dtype=np.int64)
    K, cin, cout = w.shape
    dx = np.zeros_like(xx)
    dw = np.zeros_like(w)
    L.orc_kpconv_backward(_ptr(q, _f32p), len(q), _ptr(s, _f32p), len(s), _ptr(idx, _i64p), idx.shape[1],
                          _ptr(xx, _f32p), cin, _ptr(w, _f32p), cout, _ptr(kp, _f32p), K, C.c_float(KP_extent),
                          _ptr(do, _f32p), _ptr(dx, _f32p), _ptr(dw, _f32p))
    return dx, dw


# ---------------------------------------------------------------------------------------------- oracle/_ref
def ref_batch_neighbors(queries, supports, q_batches, s_batches, radius, ordered=False):
    """The unmodified reference search (nanoflann path by default; ``ordered`` = batch_ordered_neighbors)."""
    R = ref()
    q, s, qb, sb = _f32(queries), _f32(supports), _i32(q_batches), _i32(s_batches)
    out, hmax = _i32p(), C.c_int(0)
    fn = R.ref_batch_ordered if ordered else R.ref_batch_neighbors
    rc = fn(_ptr(q, _f32p), len(q), _ptr(s, _f32p), len(s), _ptr(qb, _i32p), _ptr(sb, _i32p), len(qb),
            C.c_float(radius), C.byref(out), C.byref(hmax))
    if rc != 0:
        raise RuntimeError("Error")
    return _take(out, (len(q), hmax.value), np.int32, R.ref_free)


def ref_subsample(points, features=None, classes=None, sampleDl=0.1):
    R = ref()
    p = _f32(points)
    f = _f32(features) if features is not None else None
    c = _i32(classes) if classes is not None else None
    fdim = f.shape[1] if f is not None else 0
    ldim = (c.shape[1] if c.ndim == 2 else 1) if c is not None else 0
    op, of, oc, n = _f32p(), _f32p(), _i32p(), C.c_int(0)
    rc = R.ref_subsample(_ptr(p, _f32p), len(p), _ptr(f, _f32p), fdim, _ptr(c, _i32p), ldim, C.c_float(sampleDl),
                         C.byref(op), C.byref(of), C.byref(oc), C.byref(n))
    if rc != 0:
        raise RuntimeError("Error")
    res = [_take(op, (n.value, 3), np.float32, R.ref_free)]
    if f is not None:
        res.append(_take(of, (n.value, fdim), np.float32, R.ref_free))
    if c is not None:
        res.append(_take(oc, (n.value, ldim), np.int32, R.ref_free))
    return res[0] if len(res) == 1 else tuple(res)


def ref_subsample_batch(points, batches, features=None, classes=None, sampleDl=0.1, max_p=0):
    R = ref()
    p, b = _f32(points), _i32(batches)
    f = _f32(features) if features is not None else None
    c = _i32(classes) if classes is not None else None
    fdim = f.shape[1] if f is not None else 0
    ldim = (c.shape[1] if c.ndim == 2 else 1) if c is not None else 0
    op, of, oc, n = _f32p(), _f32p(), _i32p(), C.c_int(0)
    ob = np.zeros(len(b), np.int32)
    rc = R.ref_subsample_batch(_ptr(p, _f32p), len(p), _ptr(b, _i32p), len(b), _ptr(f, _f32p), fdim, _ptr(c, _i32p),
                               ldim, C.c_float(sampleDl), int(max_p), C.byref(op), _ptr(ob, _i32p), C.byref(of),
                               C.byref(oc), C.byref(n))
    if rc != 0:
        raise RuntimeError("Error")
    res = [_take(op, (n.value, 3), np.float32, R.ref_free), ob]
    if f is not None:
        res.append(_take(of, (n.value, fdim), np.float32, R.ref_free))
    if c is not None:
        res.append(_take(oc, (n.value, ldim), np.int32, R.ref_free))
    return tuple(res)
