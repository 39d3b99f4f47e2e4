"""weasal_b200 — B200 (sm_100a) implementation of WeaSAL's KPConv hot path.

  radius_neighbors / grid_subsampling   numpy drop-ins for the reference's two C++ extension modules
  ops                                   device-tensor entry points (radius search, grid subsampling, KPConv fwd/bwd)
  kpconv.KPConv                         drop-in nn.Module for models.blocks.KPConv
  pyramid                               device-side segmentation_inputs: one native call per batch (NativeBuild),
                                        PyramidPrefetcher (worker thread + side stream), static-shape layout
  engine                                static-shape batches + the training step as one CUDA graph
  voting                                sphere-voting inference, spheres sharded over ranks
  distributed                           gradient all-reduce, vote accumulation
  dropin.install()                      registers all of the above under the reference's module names
All arithmetic runs in libweasal_b200.so (hand-written CUDA, C-ABI in include/weasal_b200.h); there is no CPU path.
"""
__version__ = "0.1.0"
