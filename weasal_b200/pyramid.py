"""Device-side pyramid builder: the GPU counterpart of ``PointCloudDataset.segmentation_inputs``
(datasets/common.py:461-577) with ``batch_neighbors`` (common.py:185-196), ``batch_grid_subsampling`` incl. its random
grid orientation (common.py:77-182) and ``big_neighborhood_filter`` (common.py:336-346).

The reference runs this inside forked DataLoader workers on the CPU; CUDA cannot be used there, so this version runs
in the main process on the training stream and returns the same flat list
``points*L + neighbors*L + pools*L + upsamples*L + lengths*L + [features, labels]`` with device tensors
(indices int64 as collated by the reference, common.py:551-553). :class:`DeviceBatch` exposes that list with the
field names of ``<DS>CustomBatch`` (datasets/Vaihingen3D_PseudoLabel.py:1407-1481) so the unchanged networks consume it.
"""
import numpy as np
import torch

from . import ops


def axis_angle_rotations(axis, angle):
    """Rotation matrices about unit ``axis`` [B,3] by ``angle`` [B] (Rodrigues); float64 [B,3,3].
    Same matrices as the reference's create_3D_rotations (kernels/kernel_points.py:43-74)."""
    axis = np.asarray(axis, np.float64)
    angle = np.asarray(angle, np.float64)
    c, s = np.cos(angle), np.sin(angle)
    v = 1.0 - c
    x, y, z = axis[:, 0], axis[:, 1], axis[:, 2]
    R = np.empty((len(angle), 3, 3), np.float64)
    vx = v * x
    R[:, 0, 0] = c + v * (x * x)
    R[:, 0, 1] = vx * y - s * z
    R[:, 0, 2] = vx * z + s * y
    R[:, 1, 0] = vx * y + s * z
    R[:, 1, 1] = c + v * (y * y)
    R[:, 1, 2] = v * y * z - s * x
    R[:, 2, 0] = vx * z - s * y
    R[:, 2, 1] = v * y * z + s * x
    R[:, 2, 2] = c + v * (z * z)
    return R


def random_grid_rotations(B):
    """The per-batch-element grid orientation of datasets/common.py:98-111, drawing from ``np.random`` in the same
    order (theta, phi, alpha) so a seeded run reproduces the reference's pyramid."""
    theta = np.random.rand(B) * 2 * np.pi
    phi = (np.random.rand(B) - 0.5) * np.pi
    u = np.vstack([np.cos(theta) * np.cos(phi), np.sin(theta) * np.cos(phi), np.sin(phi)])
    alpha = np.random.rand(B) * 2 * np.pi
    return axis_angle_rotations(u.T, alpha).astype(np.float32)


def batch_neighbors(queries, supports, q_batches, s_batches, radius, limit=None, dtype=torch.int64):
    return ops.batch_query(queries, supports, q_batches, s_batches, radius, limit=limit, dtype=dtype)


def batch_grid_subsampling(points, batches_len, sampleDl=0.1, max_p=0, random_grid_orient=True, order="reference"):
    rot = random_grid_rotations(len(batches_len)) if random_grid_orient else None
    s_points, s_len = ops.grid_subsample(points, batches_len, sampleDl=sampleDl, max_p=max_p, order=order, rot=rot)
    return s_points, s_len


class PyramidBuilder:
    """The walk over ``config.architecture`` of datasets/common.py:461-577, one pyramid layer per :meth:`step`.

    Every radius search is issued without a host sync (``ops.PendingSearches``); grids are shared between the searches
    that use the same supports and radius. ``resolve()`` reads the widths of the searches issued so far with one
    device->host copy and swaps the placeholders for the column-sliced index matrices.
    """

    def __init__(self, stacked_points, stack_lengths, config, neighborhood_limits=None, random_grid_orient=True,
                 order="reference", index_dtype=torch.int64, device="cuda"):
        self.dev = torch.device(device)
        self.cfg, self.order, self.dtype, self.orient = config, order, index_dtype, random_grid_orient
        t = stacked_points if torch.is_tensor(stacked_points) else torch.from_numpy(np.ascontiguousarray(stacked_points))
        self.pts = t.to(self.dev, non_blocking=True).to(torch.float32)
        self.lens = np.ascontiguousarray(stack_lengths.cpu().numpy() if torch.is_tensor(stack_lengths)
                                         else stack_lengths, dtype=np.int32)
        self.limits = list(neighborhood_limits) if neighborhood_limits is not None and len(neighborhood_limits) else None
        self.r_normal = config.first_subsampling_dl * config.conv_radius
        self.points, self.neighbors, self.pools, self.upsamples, self.lengths = [], [], [], [], []
        self.pending = ops.PendingSearches(self.dev)
        self.resolved = 0
        self.grids = {}
        self.blocks = list(config.architecture)
        self.pos = 0
        self.done = False
        # number of layers = pooling / strided blocks before the first 'global' / 'upsample' block, plus one
        self.n_layers = 0
        for b in self.blocks:
            if 'pool' in b or 'strided' in b or 'global' in b or 'upsample' in b:
                self.n_layers += 1
                if 'global' in b or 'upsample' in b:
                    break

    def _lim(self, layer):
        return int(self.limits[layer]) if self.limits is not None and layer < len(self.limits) else None

    def _grid(self, p, b, r):
        key = (p.data_ptr(), p.shape[0], float(r))
        if key not in self.grids:
            self.grids[key] = ops.SearchGrid(p, b, r)
        return self.grids[key]

    def _empty(self):
        return torch.zeros((0, 1), dtype=self.dtype, device=self.dev)

    def step(self):
        """Issue the work of the next layer (conv search, subsampling, pool and upsample searches)."""
        if self.done:
            return False
        cfg, layer_blocks = self.cfg, []
        while self.pos < len(self.blocks):
            block = self.blocks[self.pos]
            self.pos += 1
            if not ('pool' in block or 'strided' in block or 'global' in block or 'upsample' in block):
                layer_blocks.append(block)
                continue
            layer, pts, lens, r_normal = len(self.points), self.pts, self.lens, self.r_normal
            if layer_blocks:
                deform = any('deformable' in b for b in layer_blocks)
                r = r_normal * cfg.deform_radius / cfg.conv_radius if deform else r_normal
                conv_i = self.pending.add(pts, pts, lens, lens, r, limit=self._lim(layer), dtype=self.dtype,
                                          grid=self._grid(pts, lens, r))
            else:
                conv_i = self._empty()
            if 'pool' in block or 'strided' in block:
                dl = 2 * r_normal / cfg.conv_radius
                pool_p, pool_b = batch_grid_subsampling(pts, lens, sampleDl=dl, random_grid_orient=self.orient,
                                                        order=self.order)
                r = r_normal * cfg.deform_radius / cfg.conv_radius if 'deformable' in block else r_normal
                pool_i = self.pending.add(pool_p, pts, pool_b, lens, r, limit=self._lim(layer), dtype=self.dtype,
                                          grid=self._grid(pts, lens, r))
                up_i = self.pending.add(pts, pool_p, lens, pool_b, 2 * r, limit=self._lim(layer + 1), dtype=self.dtype,
                                        grid=self._grid(pool_p, pool_b, 2 * r))
            else:
                pool_i, up_i = self._empty(), self._empty()
                pool_p = torch.zeros((0, 3), dtype=torch.float32, device=self.dev)
                pool_b = np.zeros((0,), np.int32)
            self.points.append(pts)
            self.neighbors.append(conv_i)
            self.pools.append(pool_i)
            self.upsamples.append(up_i)
            self.lengths.append(torch.from_numpy(lens.copy()).to(self.dev, non_blocking=True))
            self.pts, self.lens = pool_p, pool_b
            self.r_normal *= 2
            if 'global' in block or 'upsample' in block:
                self.done = True
            return True
        self.done = True
        return False

    def resolve(self):
        found = self.pending.resolve(start=self.resolved)
        for lst in (self.neighbors, self.pools, self.upsamples):
            for i, v in enumerate(lst):
                if isinstance(v, int):
                    lst[i] = found[v - self.resolved]
        self.resolved += len(found)
        for key in [k for k in self.grids if k[0] != self.pts.data_ptr()]:
            del self.grids[key]  # only the grid over the newest layer can be needed again


def segmentation_inputs(stacked_points, stacked_features, labels, stack_lengths, config, neighborhood_limits=None,
                        random_grid_orient=True, order="reference", index_dtype=torch.int64, device="cuda"):
    """Device counterpart of datasets/common.py:461-577: the flat list ``points*L + neighbors*L + pools*L +
    upsamples*L + lengths*L + [features, labels]`` of device tensors. All searches of the batch are issued before the
    single synchronisation that reads their widths back."""
    dev = torch.device(device)
    pb = PyramidBuilder(stacked_points, stack_lengths, config, neighborhood_limits, random_grid_orient, order,
                        index_dtype, dev)
    while pb.step():
        pass
    pb.resolve()

    def to_dev(a, dtype=None):
        t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
        t = t.to(dev, non_blocking=True)
        return t.to(dtype) if dtype is not None else t

    feats = to_dev(stacked_features, torch.float32) if stacked_features is not None else None
    labs = to_dev(labels) if labels is not None else None
    return pb.points + pb.neighbors + pb.pools + pb.upsamples + pb.lengths + [feats, labs]


class DeviceBatch:
    """The fields of the reference's ``<DS>CustomBatch`` (Vaihingen3D_PseudoLabel.py:1407-1481) over device tensors."""

    def __init__(self, input_list, extras=None):
        L = (len(input_list) - 2) // 5
        ind = 0
        self.points = list(input_list[ind:ind + L]); ind += L
        self.neighbors = list(input_list[ind:ind + L]); ind += L
        self.pools = list(input_list[ind:ind + L]); ind += L
        self.upsamples = list(input_list[ind:ind + L]); ind += L
        self.lengths = list(input_list[ind:ind + L]); ind += L
        self.features = input_list[ind]; ind += 1
        self.labels = input_list[ind]
        for k, v in (extras or {}).items():
            setattr(self, k, v)

    def pin_memory(self):
        return self

    def to(self, device):
        return self
