"""``KPConv`` module — drop-in for ``models.blocks.KPConv`` of the reference (models/blocks.py:144-379).

Contract kept from the reference so that checkpoints, optimiser parameter groups and the regulariser keep working:
  * constructor ``KPConv(kernel_size, p_dim, in_channels, out_channels, KP_extent, radius, fixed_kernel_points,
    KP_influence, aggregation_mode, deformable, modulated)`` (blocks.py:146-148);
  * parameters ``weights`` [K, Cin, Cout] (trainable, ``kaiming_uniform_(a=sqrt(5))`` as at blocks.py:217-218) and
    ``kernel_points`` [K, 3] (``requires_grad=False``, from ``load_kernels``, blocks.py:222-236): the state_dict keys;
  * attributes read elsewhere: ``K, p_dim, in_channels, out_channels, radius, KP_extent, deformable, min_d2,
    deformed_KP, offset_features`` (``p2p_fitting_regularizer``, architectures.py:29-57) and ``__repr__``;
  * ``forward(q_pts, s_pts, neighb_inds, x) -> [Nq, Cout]``, differentiable w.r.t. ``x`` and ``weights``.
Only the configuration every shipped WeaSAL script uses is implemented (rigid kernel, 'linear' influence, 'sum'
aggregation, 3-D); anything else raises at construction instead of silently computing something different.
"""
import math

import torch
from torch import nn

from . import ops
from .kernel_points import load_kernels

_UNSUPPORTED = "weasal_b200.KPConv supports the rigid / linear / sum / 3-D configuration only ({})"


class KPConv(nn.Module):

    def __init__(self, kernel_size, p_dim, in_channels, out_channels, KP_extent, radius,
                 fixed_kernel_points='center', KP_influence='linear', aggregation_mode='sum',
                 deformable=False, modulated=False):
        super().__init__()
        for bad, why in ((deformable or modulated, "deformable / modulated kernels, blocks.py:244-271"),
                         (KP_influence != 'linear', "KP_influence=%r, blocks.py:330-351" % (KP_influence,)),
                         (aggregation_mode != 'sum', "aggregation_mode=%r, blocks.py:352-354" % (aggregation_mode,)),
                         (p_dim != 3, "p_dim=%r" % (p_dim,))):
            if bad:
                raise NotImplementedError(_UNSUPPORTED.format(why))
        config = dict(K=kernel_size, p_dim=p_dim, in_channels=in_channels, out_channels=out_channels, radius=radius,
                      KP_extent=KP_extent, fixed_kernel_points=fixed_kernel_points, KP_influence=KP_influence,
                      aggregation_mode=aggregation_mode, deformable=False, modulated=False)
        # slots the reference fills for deformable kernels only; kept so that code probing them finds None
        state = dict(min_d2=None, deformed_KP=None, offset_features=None, offset_dim=None, offset_conv=None,
                     offset_bias=None)
        for name, value in {**config, **state}.items():
            setattr(self, name, value)
        self.weights = nn.Parameter(torch.zeros(kernel_size, in_channels, out_channels, dtype=torch.float32))
        self.reset_parameters()
        self.kernel_points = self.init_KP()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weights, a=math.sqrt(5))

    def init_KP(self):
        disposition = load_kernels(self.radius, self.K, dimension=self.p_dim, fixed=self.fixed_kernel_points)
        return nn.Parameter(torch.as_tensor(disposition, dtype=torch.float32), requires_grad=False)

    def forward(self, q_pts, s_pts, neighb_inds, x):
        return ops.kpconv(q_pts, s_pts, neighb_inds, x, self.weights, self.kernel_points, self.KP_extent)

    def __repr__(self):
        return f"KPConv(radius: {self.radius:.2f}, in_feat: {self.in_channels:d}, out_feat: {self.out_channels:d})"
