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


def segmentation_inputs(stacked_points, stacked_features, labels, stack_lengths, config, neighborhood_limits=None,
                        random_grid_orient=True, order="reference", index_dtype=torch.int64, device="cuda"):
    """Same walk over ``config.architecture`` as datasets/common.py:461-577; every array is a device tensor."""
    dev = torch.device(device)

    def to_dev(a, dtype=None):
        t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))
        t = t.to(dev, non_blocking=True)
        return t.to(dtype) if dtype is not None else t

    pts = to_dev(stacked_points, torch.float32)
    lens = np.ascontiguousarray(stack_lengths.cpu().numpy() if torch.is_tensor(stack_lengths) else stack_lengths,
                                dtype=np.int32)
    limits = list(neighborhood_limits) if neighborhood_limits is not None and len(neighborhood_limits) > 0 else None

    def lim(layer):
        return int(limits[layer]) if limits is not None else None

    r_normal = config.first_subsampling_dl * config.conv_radius
    layer_blocks = []
    in_points, in_neighbors, in_pools, in_upsamples, in_lengths = [], [], [], [], []
    empty_idx = lambda: torch.zeros((0, 1), dtype=index_dtype, device=dev)
    # every search is issued without a host sync; their widths are read back together at the end
    pending = ops.PendingSearches(dev)
    # one hash grid per (point set, radius): the grid over layer l+1 at radius 2r serves the upsample search of layer l
    # and the conv and pool searches of layer l+1
    grids = {}

    def grid_of(p, b, r):
        key = (p.data_ptr(), p.shape[0], float(r))
        if key not in grids:
            grids[key] = ops.SearchGrid(p, b, r)
        return grids[key]

    for block in config.architecture:
        if not ('pool' in block or 'strided' in block or 'global' in block or 'upsample' in block):
            layer_blocks.append(block)
            continue
        layer = len(in_points)
        if layer_blocks:
            if any('deformable' in b for b in layer_blocks):
                r = r_normal * config.deform_radius / config.conv_radius
            else:
                r = r_normal
            conv_i = pending.add(pts, pts, lens, lens, r, limit=lim(layer), dtype=index_dtype, grid=grid_of(pts, lens, r))
        else:
            conv_i = empty_idx()
        if 'pool' in block or 'strided' in block:
            dl = 2 * r_normal / config.conv_radius
            pool_p, pool_b = batch_grid_subsampling(pts, lens, sampleDl=dl, random_grid_orient=random_grid_orient,
                                                    order=order)
            if 'deformable' in block:
                r = r_normal * config.deform_radius / config.conv_radius
            else:
                r = r_normal
            pool_i = pending.add(pool_p, pts, pool_b, lens, r, limit=lim(layer), dtype=index_dtype,
                                 grid=grid_of(pts, lens, r))
            up_i = pending.add(pts, pool_p, lens, pool_b, 2 * r, limit=lim(layer + 1) if limits is not None and
                               layer + 1 < len(limits) else None, dtype=index_dtype, grid=grid_of(pool_p, pool_b, 2 * r))
        else:
            pool_i = empty_idx()
            pool_p = torch.zeros((0, 3), dtype=torch.float32, device=dev)
            pool_b = np.zeros((0,), np.int32)
            up_i = empty_idx()
        in_points.append(pts)
        in_neighbors.append(conv_i)
        in_pools.append(pool_i)
        in_upsamples.append(up_i)
        in_lengths.append(torch.from_numpy(lens.copy()).to(dev, non_blocking=True))
        pts, lens = pool_p, pool_b
        r_normal *= 2
        layer_blocks = []
        if 'global' in block or 'upsample' in block:
            break

    found = pending.resolve()
    for lst in (in_neighbors, in_pools, in_upsamples):
        for i, v in enumerate(lst):
            if isinstance(v, int):
                lst[i] = found[v]

    feats = to_dev(stacked_features, torch.float32) if stacked_features is not None else None
    labs = to_dev(labels) if labels is not None else None
    return in_points + in_neighbors + in_pools + in_upsamples + in_lengths + [feats, labs]


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
