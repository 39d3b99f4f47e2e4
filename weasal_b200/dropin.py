"""Installs the B200 path under the reference's own module names (SURVEY.md §8b / INTEGRATION.md).

After :func:`install`, the unchanged WeaSAL sources resolve
  * ``cpp_wrappers.cpp_neighbors.radius_neighbors``     -> :mod:`weasal_b200.radius_neighbors`
  * ``cpp_wrappers.cpp_subsampling.grid_subsampling``   -> :mod:`weasal_b200.grid_subsampling`
  * ``models.blocks.KPConv``                            -> :class:`weasal_b200.kpconv.KPConv` (when ``models.blocks`` is
    importable; call before ``models.architectures`` does ``from models.blocks import *``);
  * ``models.blocks.max_pool`` / ``closest_pool``        -> the gather kernels of :mod:`weasal_b200.ops` for CUDA tensors
    (blocks.py:77-112; the block classes look these functions up in the module at call time, blocks.py:704, 737), the
    reference's own functions for anything else.
"""
import sys
import types


def _patch_pools(ref_blocks):
    if getattr(ref_blocks.max_pool, "_weasal_b200", False):
        return
    ref_max, ref_closest = ref_blocks.max_pool, ref_blocks.closest_pool

    def max_pool(x, inds):
        if x.is_cuda:
            from . import ops
            return ops.max_pool(x, inds)
        return ref_max(x, inds)

    def closest_pool(x, inds):
        if x.is_cuda:
            from . import ops
            return ops.closest_pool(x, inds)
        return ref_closest(x, inds)

    max_pool._weasal_b200 = closest_pool._weasal_b200 = True
    max_pool.__doc__, closest_pool.__doc__ = ref_max.__doc__, ref_closest.__doc__
    ref_blocks.max_pool, ref_blocks.closest_pool = max_pool, closest_pool


def install(patch_kpconv=True, patch_pools=True):
    from . import grid_subsampling, radius_neighbors

    for name in ("cpp_wrappers", "cpp_wrappers.cpp_subsampling", "cpp_wrappers.cpp_neighbors"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["cpp_wrappers.cpp_subsampling.grid_subsampling"] = grid_subsampling
    sys.modules["cpp_wrappers.cpp_subsampling"].grid_subsampling = grid_subsampling
    sys.modules["cpp_wrappers.cpp_neighbors.radius_neighbors"] = radius_neighbors
    sys.modules["cpp_wrappers.cpp_neighbors"].radius_neighbors = radius_neighbors
    if patch_kpconv:
        try:
            import models.blocks as ref_blocks
        except Exception:
            return False
        from .kpconv import KPConv
        ref_blocks.KPConv = KPConv
        if patch_pools:
            _patch_pools(ref_blocks)
    return True
