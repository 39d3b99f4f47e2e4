"""Input spheres and test-time votes on the device: the two stages the reference runs on the CPU either side of the
network.

* :func:`extract_spheres` — ``potential_item``'s ``query_radius`` + re-centring
  (datasets/Vaihingen3D_PseudoLabel.py:345-365): one counting pass, one scan, one fill pass over the cloud.
* :func:`draw_augmentation` / :func:`augment` — ``augmentation_transform`` (datasets/common.py:252-334): the random draws
  come from ``np.random`` in the reference's order (rotation angle, 3 scales, 3 symmetry bits, N x 3 normals), so a seeded
  run reproduces the reference's augmented points bit for bit; the arithmetic and the feature assembly
  (Vaihingen3D_PseudoLabel.py:383, 423-430) run in one kernel.
* :class:`VoteBuffer` — ``test_probs`` of utils/tester_PseudoLabel.py:176-195 with the reference's EMA update or the
  order-independent accumulation, reprojection (:270-283) and ``fast_confusion`` (utils/metrics.py:35-118).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def extract_spheres(cloud, centres, radius, cap=None):
    """cloud [N,3] f32 CUDA; centres [B,3] (host, float64). Returns (points [n,3] f32 centred, lengths int32 [B] (numpy),
    input_inds int64 [n]) — rows of a sphere in ascending cloud index."""
    if not cloud.is_cuda:
        raise RuntimeError("weasal_b200: tensors must be CUDA tensors (there is no CPU fallback)")
    cloud = cloud.contiguous().float()
    cen = np.ascontiguousarray(centres, np.float64).reshape(-1, 3)
    nb = len(cen)
    cap = int(cap) if cap else max(int(cloud.shape[0]), 1)
    while True:
        pts = torch.empty((cap, 3), dtype=torch.float32, device=cloud.device)
        inds = torch.empty((cap,), dtype=torch.int64, device=cloud.device)
        lens = np.zeros(nb, np.int32)
        rc = _lib.lib().kp_extract_spheres_dev(cloud.data_ptr(), cloud.shape[0], cen.ctypes.data, nb, float(radius),
                                               pts.data_ptr(), inds.data_ptr(), cap, lens.ctypes.data, _stream())
        if rc == _lib.KP_ERR_CAPACITY:
            cap = int(lens.sum()) + 1
            continue
        _lib.check(rc, "extract_spheres")
        n = int(lens.sum())
        return pts[:n], lens, inds[:n]


def draw_augmentation(lengths, config):
    """The random draws of ``augmentation_transform`` for every sphere of a batch, from ``np.random`` in the reference's
    order (datasets/common.py:262-304): R [B,3,3] f32, scale [B,3] f32, noise [N,3] f32 (host arrays)."""
    Rs, scales, noises = [], [], []
    for n in lengths:
        R = np.eye(3)
        if config.augment_rotation == 'vertical':
            theta = np.random.rand() * 2 * np.pi
            c, s = np.cos(theta), np.sin(theta)
            R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float32)
        elif config.augment_rotation == 'all':
            from .pyramid import axis_angle_rotations
            theta = np.random.rand() * 2 * np.pi
            phi = (np.random.rand() - 0.5) * np.pi
            u = np.array([np.cos(theta) * np.cos(phi), np.sin(theta) * np.cos(phi), np.sin(phi)])
            alpha = np.random.rand() * 2 * np.pi
            R = axis_angle_rotations(np.reshape(u, (1, -1)), np.reshape(alpha, (1,)))[0]
        R = R.astype(np.float32)
        min_s, max_s = config.augment_scale_min, config.augment_scale_max
        if config.augment_scale_anisotropic:
            scale = np.random.rand(3) * (max_s - min_s) + min_s
        else:
            scale = np.random.rand() * (max_s - min_s) + min_s
        symmetries = np.array(config.augment_symmetries).astype(np.int32)
        symmetries *= np.random.randint(2, size=3)
        scale = (scale * (1 - symmetries * 2)).astype(np.float32)
        noise = (np.random.randn(int(n), 3) * config.augment_noise).astype(np.float32)
        Rs.append(R)
        scales.append(np.broadcast_to(scale, (3,)).astype(np.float32))
        noises.append(noise)
    return np.stack(Rs), np.stack(scales), np.concatenate(noises, 0) if noises else np.zeros((0, 3), np.float32)


def augment(points, lengths, R, scale, noise=None, colors=None, input_inds=None, centre_z=None, color_keep=None, fdim=0):
    """``augmented = (points . R_b) * scale_b + noise`` in float32 in numpy's operation order; with ``fdim`` also the input
    features ``[1, colors[input_inds] * keep_b, z_aug + centre_z_b, z_aug][:fdim]``. Returns (points, features or None)."""
    p = points.contiguous().float()
    lens = np.ascontiguousarray(lengths, np.int32)
    nb = len(lens)
    Rn, sn = np.ascontiguousarray(R, np.float32), np.ascontiguousarray(scale, np.float32)
    nz = None
    if noise is not None:
        nz = noise if torch.is_tensor(noise) else torch.from_numpy(np.ascontiguousarray(noise, np.float32))
        nz = nz.to(p.device, non_blocking=True).contiguous()
    out = torch.empty_like(p)
    feats = torch.empty((p.shape[0], fdim), dtype=torch.float32, device=p.device) if fdim else None
    col = colors.contiguous().float() if colors is not None else None
    cz = np.ascontiguousarray(centre_z, np.float32) if centre_z is not None else None
    keep = np.ascontiguousarray(color_keep, np.float32) if color_keep is not None else None
    _lib.check(_lib.lib().kp_augment_spheres_dev(
        p.data_ptr(), lens.ctypes.data, nb, Rn.ctypes.data, sn.ctypes.data, nz.data_ptr() if nz is not None else None,
        out.data_ptr(), col.data_ptr() if col is not None else None, col.shape[1] if col is not None else 0,
        input_inds.data_ptr() if input_inds is not None else None, cz.ctypes.data if cz is not None else None,
        keep.ctypes.data if keep is not None else None, feats.data_ptr() if feats is not None else None, int(fdim), _stream()),
        "augment_spheres")
    return out, feats


class VoteBuffer:
    """``test_probs`` of one cloud on the device. ``mode='ema'``: the reference's update
    ``test_probs[inds] = smooth * test_probs[inds] + (1 - smooth) * probs`` (spheres in order); ``mode='sum'``: the
    order-independent accumulation (sum of probabilities, votes) for spheres sharded over ranks."""

    def __init__(self, n_points, n_classes, device, mode="ema", smooth=0.95):
        self.mode, self.smooth, self.C = mode, float(smooth), int(n_classes)
        self.probs = torch.zeros((n_points, n_classes), dtype=torch.float32, device=device)
        self.weight = torch.zeros((n_points,), dtype=torch.float32, device=device) if mode == "sum" else None

    @torch.no_grad()
    def update(self, probs, points, input_inds, lengths, radius_limit=0.0):
        """One batch of spheres: ``probs`` [n,C] (softmax outputs), ``points`` [n,3] (centred input points), ``input_inds``
        [n] int64 cloud indices, ``lengths`` [B]; only points with |p|^2 < radius_limit^2 vote (0: all)."""
        pr, pt, ii = probs.contiguous().float(), points.contiguous().float(), input_inds.contiguous()
        lens = np.ascontiguousarray(lengths.cpu().numpy() if torch.is_tensor(lengths) else lengths, np.int32)
        _lib.check(_lib.lib().kp_vote_update_dev(pr.data_ptr(), pt.data_ptr(), ii.data_ptr(), lens.ctypes.data, len(lens), self.C,
                                                 float(radius_limit), self.smooth, 0 if self.mode == "ema" else 1,
                                                 self.probs.data_ptr(), self.weight.data_ptr() if self.weight is not None else None,
                                                 _stream()), "vote_update")

    @torch.no_grad()
    def reduce(self, group=None):
        import torch.distributed as dist
        if self.mode == "sum" and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.probs, group=group)
            dist.all_reduce(self.weight, group=group)

    @torch.no_grad()
    def reproject(self, proj=None, truth=None, return_probs=True):
        """(probabilities [M,C], predictions int32 [M], confusion int64 [C,C] or None) on the evaluation points
        ``test_proj`` (None: the cloud's own points); ``truth`` int32 [M] labels in 0..C-1 for the confusion matrix."""
        dev = self.probs.device
        m = int(proj.shape[0]) if proj is not None else int(self.probs.shape[0])
        out = torch.empty((m, self.C), dtype=torch.float32, device=dev) if return_probs else None
        pred = torch.empty((m,), dtype=torch.int32, device=dev)
        conf = torch.zeros((self.C, self.C), dtype=torch.int64, device=dev) if truth is not None else None
        tr = truth.to(torch.int32).contiguous() if truth is not None else None
        pj = proj.contiguous() if proj is not None else None
        _lib.check(_lib.lib().kp_vote_reproject_dev(self.probs.data_ptr(), self.weight.data_ptr() if self.weight is not None else None,
                                                    pj.data_ptr() if pj is not None else None, m, self.C,
                                                    out.data_ptr() if out is not None else None, pred.data_ptr(),
                                                    tr.data_ptr() if tr is not None else None,
                                                    conf.data_ptr() if conf is not None else None, _stream()), "vote_reproject")
        return out, pred, conf
