"""``KPConv`` — drop-in for the class of the same name in the reference's models/blocks.py:144-379.

Same constructor signature, attributes, parameter names (state_dict keys ``weights`` [K,Cin,Cout] and
``kernel_points`` [K,3]), initialisation calls (``kaiming_uniform_(weights, a=sqrt(5))``, blocks.py:217-218) and
``forward(q_pts, s_pts, neighb_inds, x) -> [Nq, Cout]``. Only the rigid / 'linear' / 'sum' configuration — the one
every shipped WeaSAL config uses — is implemented; the other branches raise at construction.
"""
import math

import torch
import torch.nn as nn
from torch.nn.init import kaiming_uniform_
from torch.nn.parameter import Parameter

from . import ops
from .kernel_points import load_kernels


class KPConv(nn.Module):

    def __init__(self, kernel_size, p_dim, in_channels, out_channels, KP_extent, radius,
                 fixed_kernel_points='center', KP_influence='linear', aggregation_mode='sum',
                 deformable=False, modulated=False):
        super(KPConv, self).__init__()
        if deformable or modulated:
            raise NotImplementedError("weasal_b200.KPConv: deformable / modulated KPConv is outside the hot path "
                                      "(no shipped WeaSAL config enables it, blocks.py:244-271)")
        if KP_influence != 'linear':
            raise NotImplementedError("weasal_b200.KPConv: only KP_influence='linear' (blocks.py:335-338)")
        if aggregation_mode != 'sum':
            raise NotImplementedError("weasal_b200.KPConv: only aggregation_mode='sum' (blocks.py:352-354)")
        if p_dim != 3:
            raise NotImplementedError("weasal_b200.KPConv: only 3-D points")
        self.K = kernel_size
        self.p_dim = p_dim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.radius = radius
        self.KP_extent = KP_extent
        self.fixed_kernel_points = fixed_kernel_points
        self.KP_influence = KP_influence
        self.aggregation_mode = aggregation_mode
        self.deformable = deformable
        self.modulated = modulated
        # read by p2p_fitting_regularizer (architectures.py:29-57) for deformable convs only
        self.min_d2 = None
        self.deformed_KP = None
        self.offset_features = None
        self.offset_dim = None
        self.offset_conv = None
        self.offset_bias = None
        self.weights = Parameter(torch.zeros((self.K, in_channels, out_channels), dtype=torch.float32),
                                 requires_grad=True)
        self.reset_parameters()
        self.kernel_points = self.init_KP()

    def reset_parameters(self):
        kaiming_uniform_(self.weights, a=math.sqrt(5))

    def init_KP(self):
        K_points_numpy = load_kernels(self.radius, self.K, dimension=self.p_dim, fixed=self.fixed_kernel_points)
        return Parameter(torch.tensor(K_points_numpy, dtype=torch.float32), requires_grad=False)

    def forward(self, q_pts, s_pts, neighb_inds, x):
        return ops.kpconv(q_pts, s_pts, neighb_inds, x, self.weights, self.kernel_points, self.KP_extent)

    def __repr__(self):
        return 'KPConv(radius: {:.2f}, in_feat: {:d}, out_feat: {:d})'.format(self.radius, self.in_channels,
                                                                              self.out_channels)
