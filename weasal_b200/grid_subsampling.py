"""Drop-in for the reference extension module ``cpp_wrappers.cpp_subsampling.grid_subsampling``.

``subsample`` (cpp_wrappers/cpp_subsampling/wrapper.cpp:338-566) and ``subsample_batch`` (wrapper.cpp:62-333): same
names, keyword-only options, dtypes, return arity and error messages; the work runs on the GPU through
``kp_grid_subsample_host``. Voxels come out in the reference's own order (``order="reference"``).
"""
import ctypes as C

import numpy as np

from . import _lib

_METHOD_ERR = 'Error parsing method. Valid method names are "barycenters" and "voxelcenters" '


def _coerce(obj, dtype, msg):
    try:
        return np.ascontiguousarray(obj, dtype=dtype)
    except Exception:
        raise RuntimeError(msg)


def _run(p, b, f, c, ldim, sampleDl, max_p, order, rot):
    L = _lib.lib()
    n, nb = p.shape[0], b.shape[0]
    fdim = f.shape[1] if f is not None else 0
    op, of, oc = _lib.c_f32p(), _lib.c_f32p(), _lib.c_i32p()
    ob = np.zeros(nb, np.int32)
    m = C.c_int(0)
    r = np.ascontiguousarray(rot, np.float32) if rot is not None else None
    rc = L.kp_grid_subsample_host(p.ctypes.data, n, b.ctypes.data, nb, f.ctypes.data if f is not None else None, fdim,
                                  c.ctypes.data if c is not None else None, ldim, float(sampleDl), int(max_p),
                                  1 if order == "reference" else 0, r.ctypes.data if r is not None else None,
                                  C.byref(op), ob.ctypes.data, C.byref(of), C.byref(oc), C.byref(m))
    if rc == _lib.KP_ERR_EMPTY:
        raise RuntimeError("Error")  # wrapper.cpp:266-270
    _lib.check(rc, "grid_subsampling")
    M = m.value
    sp = np.ctypeslib.as_array(op, shape=(M * 3,)).copy().reshape(M, 3)
    L.kp_free_host(op)
    res = [sp, ob]
    if f is not None:
        res.append(np.ctypeslib.as_array(of, shape=(M * fdim,)).copy().reshape(M, fdim))
        L.kp_free_host(of)
    if c is not None:
        res.append(np.ctypeslib.as_array(oc, shape=(M * ldim,)).copy().reshape(M, ldim))
        L.kp_free_host(oc)
    return res


def _prepare(points, features, classes):
    p = _coerce(points, np.float32, "Error converting input points to numpy arrays of type float32")
    f = _coerce(features, np.float32, "Error converting input features to numpy arrays of type float32") \
        if features is not None else None
    c = _coerce(classes, np.int32, "Error converting input classes to numpy arrays of type int32") \
        if classes is not None else None
    if p.ndim != 2 or p.shape[1] != 3:
        raise RuntimeError("Wrong dimensions : points.shape is not (N, 3)")
    if f is not None and f.ndim != 2:
        raise RuntimeError("Wrong dimensions : features.shape is not (N, d)")
    if c is not None and c.ndim > 2:
        raise RuntimeError("Wrong dimensions : classes.shape is not (N,) or (N, d)")
    ldim = 1
    if c is not None and c.ndim == 2:
        ldim = c.shape[1]
    if f is not None and f.shape[0] != p.shape[0]:
        raise RuntimeError("Wrong dimensions : features.shape is not (N, d)")
    if c is not None and c.shape[0] != p.shape[0]:
        raise RuntimeError("Wrong dimensions : classes.shape is not (N,) or (N, d)")
    return p, f, c, ldim


def subsample(points, *, features=None, classes=None, sampleDl=0.1, method="barycenters", verbose=0,
              order="reference"):
    if method not in ("barycenters", "voxelcenters"):
        raise RuntimeError(_METHOD_ERR)
    p, f, c, ldim = _prepare(points, features, classes)
    res = _run(p, np.asarray([p.shape[0]], np.int32), f, c, ldim, sampleDl, 0, order, None)
    res.pop(1)  # no batch lengths in the single-cloud form
    return res[0] if len(res) == 1 else tuple(res)


def subsample_batch(points, batches, *, features=None, classes=None, sampleDl=0.1, method="barycenters", max_p=0,
                    verbose=0, order="reference", rot=None):
    if method not in ("barycenters", "voxelcenters"):
        raise RuntimeError(_METHOD_ERR)
    p, f, c, ldim = _prepare(points, features, classes)
    b = _coerce(batches, np.int32, "Error converting input batches to numpy arrays of type int32")
    if b.ndim > 1:
        raise RuntimeError("Wrong dimensions : batches.shape is not (B,) ")
    return tuple(_run(p, b.reshape(-1), f, c, ldim, sampleDl, max_p, order, rot))
