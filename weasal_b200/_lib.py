"""ctypes loader for libweasal_b200.so (the C-ABI declared in include/weasal_b200.h).

There is no CPU fallback: if the library cannot be loaded (and cannot be built because nvcc is absent) importing any
compute entry point raises; if it loads but no CUDA device is present every compute call returns KP_ERR_CUDA, which
:func:`check` turns into a RuntimeError.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WEASAL_B200_LIB") or os.path.join(_HERE, "libweasal_b200.so")  # (env: experiment builds)
_lib = None

c_f32p = C.POINTER(C.c_float)
c_i32p = C.POINTER(C.c_int)
vp = C.c_void_p

KP_OK, KP_ERR_CUDA, KP_ERR_ARG, KP_ERR_CAPACITY, KP_ERR_TOO_DENSE, KP_ERR_EMPTY, KP_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6

# every symbol include/weasal_b200.h declares (tests check the library exports all of them)
SYMBOLS = ["kp_last_error", "kp_version", "kp_launch_count", "kp_free_host", "kp_batch_query_host",
           "kp_batch_query_dev", "kp_batch_query_dev_async", "kp_search_grid_bytes", "kp_search_grid_build_dev",
           "kp_search_grid_query_dev", "kp_grid_subsample_host", "kp_grid_subsample_dev", "kp_kpconv_forward_dev",
           "kp_kpconv_backward_dev", "kp_kpconv_lists_bytes", "kp_kpconv_forward_keep_dev",
           "kp_kpconv_backward_kept_dev", "kp_transpose_table_dev", "kp_kpconv_wf_dev", "kp_kpconv_dx_atomic_dev", "kp_profile_enable",
           "kp_profile_read", "kp_max_pool_forward_dev", "kp_max_pool_backward_dev", "kp_closest_pool_dev",
           "kp_pyramid_build_dev", "kp_pyramid_build_static_dev", "kp_kpconv_backward_sym_dev",
           "kp_linear_forward_dev", "kp_linear_backward_dev", "kp_closest_pool_strided_dev", "kp_max_pool_forward_width_dev", "kp_plan_ksplit", "kp_sm_partition_streams",
           "kp_kpconv_lists_build_dev", "kp_kpconv_apply_lists_dev", "kp_kpconv_dw_lists_dev", "kp_pack_image_floats",
           "kp_pack_weights_dev", "kp_linear_forward_packed_dev", "kp_linear_dx_packed_dev", "kp_linear_dw_dev",
           "kp_kpconv_prepare_dev", "kp_extract_spheres_dev", "kp_augment_spheres_dev", "kp_vote_update_dev",
           "kp_vote_reproject_dev"]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.environ.get("WEASAL_B200_LIB"):
        from . import build as _build
        # a no-op when the library is newer than every source; rebuilds after an edit of csrc/ or the header, so a stale
        # library is never loaded silently. Without nvcc (a box that only received the prebuilt .so) the check is skipped.
        if not os.path.exists(LIB_PATH) or (_build.needs_build() and os.path.exists(_build.nvcc_path())):
            _build.build()
    L = C.CDLL(LIB_PATH)
    L.kp_last_error.restype = C.c_char_p
    L.kp_launch_count.restype = C.c_longlong
    L.kp_free_host.argtypes = [vp]
    L.kp_batch_query_host.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, C.c_int, C.c_float, C.POINTER(c_i32p), c_i32p]
    L.kp_batch_query_dev.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, C.c_int, C.c_float, vp, C.c_int, C.c_int,
                                     c_i32p, vp]
    L.kp_batch_query_dev_async.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, C.c_int, C.c_float, vp, C.c_int, C.c_int,
                                           vp, vp]
    L.kp_search_grid_bytes.argtypes = [C.c_int, C.c_int]
    L.kp_search_grid_bytes.restype = C.c_longlong
    L.kp_search_grid_build_dev.argtypes = [vp, C.c_int, vp, C.c_int, C.c_float, vp, vp]
    L.kp_search_grid_query_dev.argtypes = [vp, C.c_int, C.c_int, C.c_float, vp, C.c_int, vp, vp, C.c_int, C.c_int, c_i32p,
                                           vp, vp]
    L.kp_grid_subsample_host.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_float, C.c_int,
                                         C.c_int, vp, C.POINTER(c_f32p), vp, C.POINTER(c_f32p), C.POINTER(c_i32p),
                                         c_i32p]
    L.kp_grid_subsample_dev.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_float, C.c_int,
                                        C.c_int, vp, vp, vp, vp, vp, c_i32p, vp]
    conv_common = [vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]
    L.kp_kpconv_forward_dev.argtypes = conv_common + [vp, C.c_int, vp, C.c_int, C.c_float, vp, vp]
    L.kp_kpconv_backward_dev.argtypes = conv_common + [vp, C.c_int, vp, C.c_int, C.c_float, vp, vp, vp, vp]
    L.kp_kpconv_lists_bytes.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.kp_kpconv_lists_bytes.restype = None
    L.kp_kpconv_forward_keep_dev.argtypes = conv_common + [vp, C.c_int, vp, C.c_int, C.c_float, vp, vp, vp, vp]
    L.kp_kpconv_backward_kept_dev.argtypes = conv_common + [vp, C.c_int, vp, C.c_int, C.c_float, vp, vp, vp, vp, vp, vp,
                                                            vp, vp]
    L.kp_transpose_table_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.kp_kpconv_wf_dev.argtypes = conv_common + [vp, C.c_int, C.c_float, vp, vp]
    L.kp_kpconv_dx_atomic_dev.argtypes = conv_common + [vp, C.c_int, C.c_float, vp, vp]
    L.kp_profile_enable.argtypes = [C.c_int]
    L.kp_profile_read.argtypes = [C.c_char_p, C.c_int]
    L.kp_max_pool_forward_dev.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]
    L.kp_max_pool_forward_width_dev.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
    L.kp_max_pool_backward_dev.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, vp]
    L.kp_closest_pool_dev.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]
    L.kp_kpconv_backward_sym_dev.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int, vp,
                                             C.c_int, C.c_float, vp, vp, vp, vp, vp, vp]
    L.kp_pyramid_build_dev.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int,
                                       vp, C.c_longlong, vp, vp, vp, vp, vp, vp, vp, vp]
    L.kp_pyramid_build_static_dev.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int,
                                              C.c_int, vp, vp, C.c_int, vp, C.c_longlong, vp, C.c_longlong, vp, vp, vp,
                                              vp, vp, vp, vp, vp]
    L.kp_closest_pool_strided_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]
    L.kp_plan_ksplit.argtypes = [C.c_int, C.c_int, C.c_int]
    L.kp_sm_partition_streams.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    L.kp_linear_forward_dev.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int, C.c_float, vp, vp]
    L.kp_linear_backward_dev.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp, C.c_float, vp, vp, vp, vp]
    L.kp_kpconv_lists_build_dev.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_longlong, vp,
                                            C.c_int, C.c_float, C.c_float, vp, vp, C.c_longlong, vp]
    L.kp_kpconv_apply_lists_dev.argtypes = [C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp,
                                            C.c_float, vp]
    L.kp_kpconv_dw_lists_dev.argtypes = [C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, vp]
    L.kp_pack_image_floats.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    L.kp_pack_image_floats.restype = C.c_longlong
    L.kp_pack_weights_dev.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, vp]
    L.kp_linear_forward_packed_dev.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int, C.c_float, vp, vp]
    L.kp_linear_dx_packed_dev.argtypes = [C.c_int, C.c_int, vp, C.c_int, vp, C.c_float, vp, vp, vp]
    L.kp_linear_dw_dev.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_float, vp, vp, vp]
    L.kp_kpconv_prepare_dev.argtypes = [vp, C.c_int, vp, vp]
    L.kp_extract_spheres_dev.argtypes = [vp, C.c_longlong, vp, C.c_int, C.c_double, vp, vp, C.c_longlong, vp, vp]
    L.kp_augment_spheres_dev.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, C.c_int, vp]
    L.kp_vote_update_dev.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, vp, vp, vp]
    L.kp_vote_reproject_dev.argtypes = [vp, vp, vp, C.c_longlong, C.c_int, vp, vp, vp, vp, vp]
    _lib = L
    return L


class ListJob(C.Structure):
    """struct kp_list_job of include/weasal_b200.h"""
    _fields_ = [("kind", C.c_int), ("centres", vp), ("nc", C.c_int), ("others", vp), ("no", C.c_int),
                ("neighb_inds", vp), ("idx_is_i64", C.c_int), ("H", C.c_int), ("idx_stride", C.c_int),
                ("rowptr", vp), ("col", vp), ("n_pairs", C.c_longlong),
                ("kernel_points", vp), ("K", C.c_int), ("kp_sign", C.c_float), ("KP_extent", C.c_float),
                ("hdr", vp), ("entries", vp), ("entries_cap", C.c_longlong)]


def last_error():
    return lib().kp_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != KP_OK:
        raise RuntimeError(f"{what}: {last_error()} (status {rc})")


def launch_count():
    return int(lib().kp_launch_count())
