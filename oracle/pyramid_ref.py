"""TEST INFRASTRUCTURE — not product code.

CPU pyramid builder over the compiled reference cores (oracle/_ref) or, when those are absent, the C restatement:
the same walk as the reference's ``segmentation_inputs`` (datasets/common.py:461-577) with ``batch_neighbors``
(:185-196), ``batch_grid_subsampling`` incl. its numpy grid rotation (:77-135) and ``big_neighborhood_filter``
(:336-346). bench.py times it as the reference's CPU precompute; tests compare the device pyramid against it.
"""
import numpy as np

import oracle
from weasal_b200.pyramid import random_grid_rotations


def _search(q, s, qb, sb, r, use_ref):
    if use_ref:
        return oracle.ref_batch_neighbors(q, s, qb, sb, r)
    return oracle.batch_neighbors(q, s, qb, sb, r)


def _subsample(p, lens, dl, use_ref, random_grid_orient):
    R = None
    if random_grid_orient:
        R = random_grid_rotations(len(lens))
        p = p.copy()
        i0 = 0
        for bi, n in enumerate(lens):
            p[i0:i0 + n] = np.sum(np.expand_dims(p[i0:i0 + n], 2) * R[bi], axis=1)  # common.py:118
            i0 += n
    sp, sl = (oracle.ref_subsample_batch if use_ref else oracle.grid_subsample_batch)(p, lens, sampleDl=dl)
    if random_grid_orient:
        i0 = 0
        for bi, n in enumerate(sl):
            sp[i0:i0 + n] = np.sum(np.expand_dims(sp[i0:i0 + n], 2) * R[bi].T, axis=1)  # common.py:134
            i0 += n
    return sp, sl


def segmentation_inputs_cpu(points, features, labels, lengths, config, neighborhood_limits=None,
                            random_grid_orient=True, use_ref=None):
    if use_ref is None:
        use_ref = oracle.ref_available()
    pts = np.ascontiguousarray(points, np.float32)
    lens = np.ascontiguousarray(lengths, np.int32)
    limits = list(neighborhood_limits) if neighborhood_limits is not None and len(neighborhood_limits) else None
    crop = (lambda m, l: m[:, :limits[l]]) if limits is not None else (lambda m, l: m)
    r_normal = config.first_subsampling_dl * config.conv_radius
    blocks, P, N, PO, UP, LE = [], [], [], [], [], []
    for block in config.architecture:
        if not ('pool' in block or 'strided' in block or 'global' in block or 'upsample' in block):
            blocks.append(block)
            continue
        layer = len(P)
        conv_i = _search(pts, pts, lens, lens, r_normal, use_ref) if blocks else np.zeros((0, 1), np.int32)
        if 'pool' in block or 'strided' in block:
            dl = 2 * r_normal / config.conv_radius
            pool_p, pool_b = _subsample(pts, lens, dl, use_ref, random_grid_orient)
            pool_i = _search(pool_p, pts, pool_b, lens, r_normal, use_ref)
            up_i = _search(pts, pool_p, lens, pool_b, 2 * r_normal, use_ref)
        else:
            pool_i = np.zeros((0, 1), np.int32)
            pool_p = np.zeros((0, 3), np.float32)
            pool_b = np.zeros((0,), np.int32)
            up_i = np.zeros((0, 1), np.int32)
        conv_i, pool_i = crop(conv_i, layer), crop(pool_i, layer)
        if up_i.shape[0] > 0 and limits is not None and layer + 1 < len(limits):
            up_i = crop(up_i, layer + 1)
        P.append(pts); N.append(conv_i.astype(np.int64)); PO.append(pool_i.astype(np.int64))
        UP.append(up_i.astype(np.int64)); LE.append(lens)
        pts, lens = pool_p, pool_b
        r_normal *= 2
        blocks = []
        if 'global' in block or 'upsample' in block:
            break
    return P + N + PO + UP + LE + [features, labels]
