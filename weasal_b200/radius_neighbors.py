"""Drop-in for the reference extension module ``cpp_wrappers.cpp_neighbors.radius_neighbors``.

Same name, arguments, dtypes and error behaviour as cpp_wrappers/cpp_neighbors/wrapper.cpp:58-238
(``batch_query(queries, supports, q_batches, s_batches, radius=0.1)``, ``radius`` keyword-only, int32 result
``[Nq, Hmax]`` padded with ``Ns``); the work runs on the GPU through ``kp_batch_query_host``.
"""
import ctypes as C

import numpy as np

from . import _lib


def _coerce(obj, dtype, msg):
    try:
        return np.ascontiguousarray(obj, dtype=dtype)  # PyArray_FROM_OTF(..., NPY_IN_ARRAY), wrapper.cpp:83-86
    except Exception:
        raise RuntimeError(msg)


def batch_query(queries, supports, q_batches, s_batches, *, radius=0.1):
    q = _coerce(queries, np.float32, "Error converting query points to numpy arrays of type float32")
    s = _coerce(supports, np.float32, "Error converting support points to numpy arrays of type float32")
    qb = _coerce(q_batches, np.int32, "Error converting query batches to numpy arrays of type int32")
    sb = _coerce(s_batches, np.int32, "Error converting support batches to numpy arrays of type int32")
    if q.ndim != 2 or q.shape[1] != 3:
        raise RuntimeError("Wrong dimensions : query.shape is not (N, 3)")
    if s.ndim != 2 or s.shape[1] != 3:
        raise RuntimeError("Wrong dimensions : support.shape is not (N, 3)")
    if qb.ndim > 1:
        raise RuntimeError("Wrong dimensions : queries_batches.shape is not (B,) ")
    if sb.ndim > 1:
        raise RuntimeError("Wrong dimensions : supports_batches.shape is not (B,) ")
    qb, sb = qb.reshape(-1), sb.reshape(-1)
    if qb.shape[0] != sb.shape[0]:
        raise RuntimeError("Wrong number of batch elements: different for queries and supports ")
    L = _lib.lib()
    out = _lib.c_i32p()
    hmax = C.c_int(0)
    rc = L.kp_batch_query_host(q.ctypes.data, q.shape[0], s.ctypes.data, s.shape[0], qb.ctypes.data, sb.ctypes.data,
                               qb.shape[0], float(radius), C.byref(out), C.byref(hmax))
    if rc == _lib.KP_ERR_EMPTY:
        raise RuntimeError("Error")  # wrapper.cpp:201-205
    _lib.check(rc, "batch_query")
    res = np.ctypeslib.as_array(out, shape=(q.shape[0] * hmax.value,)).copy().reshape(q.shape[0], hmax.value)
    L.kp_free_host(out)
    return res
