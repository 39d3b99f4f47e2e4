"""Kernel-point dispositions for standalone use (no reference tree on the machine).

Under the reference, ``models.blocks.KPConv.init_KP`` calls ``kernels.kernel_points.load_kernels`` (blocks.py:222-236),
which reads the cached 15-point disposition ``kernels/dispositions/k_015_center_3D.ply`` and applies a random
z-rotation, N(0, 0.01) noise and the radius scale (kernel_points.py:452-487). That call is used unchanged whenever
the reference's ``kernels`` package is importable. On a machine without the reference (the GPU bench box) this
module supplies a stand-in disposition of the same shape — a centre point plus the 14 vertices of a rhombic
dodecahedron at 0.66 of the unit radius, which is what the reference's repulsion optimiser converges close to — and
then applies the same rotation / noise / scale recipe from ``np.random``. Trained checkpoints are unaffected:
``kernel_points`` is part of the state_dict and is loaded, never regenerated.
"""
import numpy as np


def _stand_in_disposition():
    pts = [[0.0, 0.0, 0.0]]
    for a in range(3):
        for s in (-1.0, 1.0):
            v = [0.0, 0.0, 0.0]
            v[a] = s
            pts.append(v)
    for sx in (-1.0, 1.0):
        for sy in (-1.0, 1.0):
            for sz in (-1.0, 1.0):
                pts.append([sx / np.sqrt(3.0), sy / np.sqrt(3.0), sz / np.sqrt(3.0)])
    return 0.66 * np.asarray(pts, np.float64)


def load_kernels(radius, num_kpoints, dimension=3, fixed="center"):
    try:
        from kernels.kernel_points import load_kernels as ref_load  # the reference's own, when present
        return ref_load(radius, num_kpoints, dimension=dimension, fixed=fixed)
    except Exception:
        pass
    if num_kpoints != 15 or dimension != 3 or fixed != "center":
        raise NotImplementedError("stand-in disposition exists for 15 kernel points, 3-D, fixed='center' only")
    kp = _stand_in_disposition()
    theta = np.random.rand() * 2 * np.pi
    c, s = np.cos(theta), np.sin(theta)
    R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float32)
    kp = kp + np.random.normal(scale=0.01, size=kp.shape)
    kp = radius * kp
    return np.matmul(kp, R).astype(np.float32)
