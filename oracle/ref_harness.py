"""TEST INFRASTRUCTURE — not product code.

Import harness for the UNMODIFIED reference Python sources. The reference tree is looked up
in this order: ``$WEASAL_REF_ROOT``, ``baseline/_ref`` (the git-ignored install that travels to the GPU box, made by
``tools/install_reference.py``), ``/root/reference`` (the build container only).

    root = ref_harness.find_root()                      # None when no copy of the reference is on this machine
    ref_harness.install(root, backend="oracle_ref")     # or backend="weasal_b200" (the drop-in under test)
    from models.architectures import KPFCNN             # the reference's own modules, unmodified

What ``install`` does, and why:
  * stubs ``matplotlib{,.pyplot,.cm}`` (imported at kernels/kernel_points.py:28-29 for debug plots only) and
    ``torch_scatter`` (architectures.py:20; used by ``contrast_loss`` only) with a mean-``scatter`` on ``index_add_``;
  * registers ``datasets / utils / models / kernels`` as packages rooted in the reference (site-packages holds an
    unrelated ``datasets`` that would otherwise win);
  * backs the two extension-module names ``cpp_wrappers.cpp_neighbors.radius_neighbors`` and
    ``cpp_wrappers.cpp_subsampling.grid_subsampling`` either with the compiled reference cores (``oracle/_ref``) or with
    the product's drop-in modules (``weasal_b200.dropin.install``), which is the seam the product uses;
  * ``chdir`` to the reference root (``load_kernels`` opens the relative path ``kernels/dispositions``,
    kernel_points.py:410);
  * without a GPU, ``torch.Tensor.cuda`` becomes the identity (hard-coded ``.cuda()`` calls).
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)


def find_root():
    for cand in (os.environ.get("WEASAL_REF_ROOT"), os.path.join(_ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isdir(os.path.join(cand, "models")) and os.path.isdir(os.path.join(cand, "kernels")):
            return cand
    return None


def _scatter(src, index, dim=0, reduce="mean", **_):
    import torch
    n = int(index.max()) + 1 if index.numel() else 0
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    out.index_add_(0, index, src)
    if reduce == "mean":
        cnt = torch.zeros(n, dtype=src.dtype, device=src.device)
        cnt.index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        out = out / cnt.clamp_min(1).reshape((n,) + (1,) * (src.dim() - 1))
    return out


def install(root, backend="oracle_ref", chdir=True):
    import torch

    if _ROOT not in sys.path:
        sys.path.insert(0, _ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm"):
        sys.modules.setdefault(name, types.ModuleType(name))
    ts = types.ModuleType("torch_scatter")
    ts.scatter = _scatter
    sys.modules["torch_scatter"] = ts
    for n in ("datasets", "utils", "models", "kernels"):
        for k in [k for k in sys.modules if k == n or k.startswith(n + ".")]:
            del sys.modules[k]
        m = types.ModuleType(n)
        m.__path__ = [os.path.join(root, n)]
        sys.modules[n] = m
    for n in ("cpp_wrappers", "cpp_wrappers.cpp_subsampling", "cpp_wrappers.cpp_neighbors"):
        m = types.ModuleType(n)
        m.__path__ = []
        sys.modules[n] = m
    if backend == "oracle_ref":
        import oracle

        gs = types.ModuleType("cpp_wrappers.cpp_subsampling.grid_subsampling")

        def subsample(points, features=None, classes=None, sampleDl=0.1, method="barycenters", verbose=0):
            return oracle.ref_subsample(points, features, classes, sampleDl)

        def subsample_batch(points, batches, features=None, classes=None, sampleDl=0.1, method="barycenters",
                            max_p=0, verbose=0):
            return oracle.ref_subsample_batch(points, batches, features, classes, sampleDl, max_p)

        gs.subsample, gs.subsample_batch = subsample, subsample_batch
        sys.modules[gs.__name__] = gs
        sys.modules["cpp_wrappers.cpp_subsampling"].grid_subsampling = gs
        rn = types.ModuleType("cpp_wrappers.cpp_neighbors.radius_neighbors")

        def batch_query(queries, supports, q_batches, s_batches, radius=0.1):
            return oracle.ref_batch_neighbors(queries, supports, q_batches, s_batches, radius)

        rn.batch_query = batch_query
        sys.modules[rn.__name__] = rn
        sys.modules["cpp_wrappers.cpp_neighbors"].radius_neighbors = rn
    elif backend == "weasal_b200":
        from weasal_b200 import dropin
        dropin.install(patch_kpconv=False)  # the extension-module names only; KPConv is swapped by the caller
    else:
        raise KeyError(backend)
    if root not in sys.path:
        sys.path.insert(0, root)
    if chdir:
        os.chdir(root)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    return root
