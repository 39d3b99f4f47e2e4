"""Installs the B200 path under the reference's own module names (SURVEY.md §8b / INTEGRATION.md).

After :func:`install`, the unchanged WeaSAL sources resolve
  * ``cpp_wrappers.cpp_neighbors.radius_neighbors``     -> :mod:`weasal_b200.radius_neighbors`
  * ``cpp_wrappers.cpp_subsampling.grid_subsampling``   -> :mod:`weasal_b200.grid_subsampling`
  * ``models.blocks.KPConv``                            -> :class:`weasal_b200.kpconv.KPConv` (when ``models.blocks`` is
    importable; call before ``models.architectures`` does ``from models.blocks import *``).
"""
import sys
import types


def install(patch_kpconv=True):
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
    return True
