"""Installs the B200 path under the reference's own module names (SURVEY.md §8b / INTEGRATION.md).

After :func:`install`, the unchanged WeaSAL sources resolve
  * ``cpp_wrappers.cpp_neighbors.radius_neighbors``     -> :mod:`weasal_b200.radius_neighbors`
  * ``cpp_wrappers.cpp_subsampling.grid_subsampling``   -> :mod:`weasal_b200.grid_subsampling`
  * ``models.blocks.KPConv``                            -> :class:`weasal_b200.kpconv.KPConv` (when ``models.blocks`` is
    importable; call before ``models.architectures`` does ``from models.blocks import *``);
  * ``models.blocks.max_pool`` / ``closest_pool``        -> the gather kernels of :mod:`weasal_b200.ops` for CUDA tensors
    (blocks.py:77-112; the block classes look these functions up in the module at call time, blocks.py:704, 737), the
    reference's own functions for anything else.
:func:`collate_device` hands a DEVICE pyramid (``weasal_b200.pyramid.segmentation_inputs`` / ``PyramidPrefetcher``) to the
reference's unchanged ``<DS>CustomBatch.__init__`` (datasets/Vaihingen3D_PseudoLabel.py:1407-1447).
"""
import contextlib
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


@contextlib.contextmanager
def _tensors_pass_from_numpy():
    import torch
    orig = torch.from_numpy
    torch.from_numpy = lambda a: a if torch.is_tensor(a) else orig(a)
    try:
        yield
    finally:
        torch.from_numpy = orig


def collate_device(batch_cls, input_list):
    """``batch_cls([input_list])`` for a flat list whose entries are already device tensors.

    The reference's collate (``Vaihingen3DPLCollate`` -> ``Vaihingen3DPLCustomBatch.__init__``,
    datasets/Vaihingen3D_PseudoLabel.py:1407-1447, 1544-1547) wraps every entry with ``torch.from_numpy`` because its
    pyramid comes out of DataLoader workers as numpy arrays. The device pyramid is built in the training process, on the
    GPU; for the duration of the constructor ``torch.from_numpy`` lets tensors through unchanged, so the UNMODIFIED
    class unflattens the list (same ``L = (len - 7) // 5`` arithmetic, same field names); ``batch.to(device)`` of a
    batch that is already on the device is a no-op. (``pin_memory()`` is a host-memory notion: do not wrap a device
    batch source in a ``DataLoader(pin_memory=True)``.)"""
    with _tensors_pass_from_numpy():
        return batch_cls([list(input_list)])
