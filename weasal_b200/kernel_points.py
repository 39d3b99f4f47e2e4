"""Kernel-point dispositions for KPConv (models/blocks.py:222-236 `init_KP` -> kernels/kernel_points.py:407-488).

``load_kernels`` of the reference reads the cached 15-point disposition ``kernels/dispositions/k_015_center_3D.ply``
(float64, unit radius, point 0 = centre) and applies, from the global ``np.random`` stream and in this order: one
``rand()`` for a z-rotation angle, one ``normal(scale=0.01)`` of the table's shape, the radius scale, the rotation
(kernel_points.py:452-487). This module is that recipe on the SAME table, kept here as a constant (a data table has
one spelling), so a network built on a machine without the reference tree (the GPU bench box) has the reference's
kernel geometry and, for equal ``np.random`` seeds, bit-identical ``kernel_points``. Every shipped WeaSAL config uses
15 kernel points, 3-D, ``fixed='center'``; anything else raises. Trained checkpoints are unaffected either way:
``kernel_points`` is part of the state_dict and is loaded, never regenerated.
"""
import numpy as np

# kernels/dispositions/k_015_center_3D.ply, columns x y z, float64 as stored
K015_CENTER_3D = np.array([
    [0.0, 0.0, 0.0],
    [0.21565234317174373, -0.4097143410725172, 0.4717650114994032],
    [0.31987454147346756, 0.16360462180037819, 0.5548403205633189],
    [0.37552668279302426, 0.2996755631037512, -0.4539914841882228],
    [-0.40191500040081674, -0.4807627518926672, 0.21039364243609365],
    [0.6426173999561108, -0.11961323509685517, 0.01931999761315348],
    [-0.36977619681370755, -0.40220590672700146, -0.37206178700167797],
    [0.3997962924097483, 0.5183047602539482, 0.09196275592168036],
    [-0.3277127915676838, -0.023541523575169566, 0.5735714178617418],
    [-0.22140282692111046, 0.5122446853093123, 0.35428825970911637],
    [-0.26344508459577765, 0.1335763672949114, -0.5913443074273487],
    [0.27131493869758383, -0.2736453445641754, -0.5370664813263086],
    [-0.6426174011244471, 0.11961322832361423, -0.01931999640633534],
    [-0.18927448055730278, 0.5908034757259142, -0.22816748196551412],
    [0.19136158396576897, -0.6283396028006597, -0.07418986659617464],
], dtype=np.float64)


def load_kernels(radius, num_kpoints, dimension=3, fixed="center"):
    if num_kpoints != 15 or dimension != 3 or fixed != "center":
        raise NotImplementedError("weasal_b200 ships the 15-point, 3-D, fixed='center' disposition only "
                                  "(the one every WeaSAL config uses)")
    theta = np.random.rand() * 2 * np.pi                                        # kernel_points.py:455
    c, s = np.cos(theta), np.sin(theta)
    R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float32)          # :463-464
    kp = K015_CENTER_3D + np.random.normal(scale=0.01, size=K015_CENTER_3D.shape)  # :481
    kp = radius * kp                                                            # :484
    return np.matmul(kp, R).astype(np.float32)                                  # :487-488
