"""TEST INFRASTRUCTURE — not product code.

PyTorch restatement of the reference's rigid KPConv operator as the chain of aten ops the reference executes
(models/blocks.py:277-374, active lines 278-298, 335-338, 357-374): shadow point / zero feature row appended,
index gather, [Nq,H,K,3] differences, clamp(1 - sqrt(d2)/extent), batched matmul, K matmuls, sum. It is what
bench.py times as the reference's CPU implementation of the operator (the reference itself cannot travel to the GPU
box) and what tests use for autograd gradients on CPU. Pinned against the reference's own KPConv outputs in
tests/golden/kpconv_ref.npz (tests/test_boundary_cpu.py).
"""
import math

import torch
import torch.nn as nn


def kpconv_reference_ops(q_pts, s_pts, neighb_inds, x, weights, kernel_points, KP_extent):
    s_pad = torch.cat((s_pts, torch.zeros_like(s_pts[:1, :]) + 1e6), 0)          # blocks.py:278
    neighbors = s_pad[neighb_inds, :] - q_pts.unsqueeze(1)                       # :281-284
    differences = neighbors.unsqueeze(2) - kernel_points                         # :294-295
    sq_distances = torch.sum(differences ** 2, dim=3)                            # :298
    all_weights = torch.clamp(1 - torch.sqrt(sq_distances) / KP_extent, min=0.0)  # :337
    all_weights = torch.transpose(all_weights, 1, 2)                             # :338
    x_pad = torch.cat((x, torch.zeros_like(x[:1, :])), 0)                        # :357
    # blocks.py:36-65 `gather` method 2: expand + Tensor.gather so that backward is a scatter_add
    idx = neighb_inds.unsqueeze(2).expand(-1, -1, x_pad.shape[1])
    neighb_x = x_pad.unsqueeze(1).expand(-1, neighb_inds.shape[1], -1).gather(0, idx)  # :360
    weighted = torch.matmul(all_weights, neighb_x)                               # :363
    weighted = weighted.permute((1, 0, 2))                                       # :370
    return torch.sum(torch.matmul(weighted, weights), dim=0)                     # :371-374


class KPConvTorch(nn.Module):
    """Same parameters / call signature as models.blocks.KPConv; forward = the aten chain above."""

    def __init__(self, kernel_size, p_dim, in_channels, out_channels, KP_extent, radius, kernel_points=None, **_):
        super().__init__()
        self.K, self.in_channels, self.out_channels = kernel_size, in_channels, out_channels
        self.KP_extent, self.radius = KP_extent, radius
        self.weights = nn.Parameter(torch.zeros((kernel_size, in_channels, out_channels), dtype=torch.float32))
        nn.init.kaiming_uniform_(self.weights, a=math.sqrt(5))
        if kernel_points is None:
            from weasal_b200.kernel_points import load_kernels
            kernel_points = load_kernels(radius, kernel_size, dimension=p_dim, fixed="center")
        self.kernel_points = nn.Parameter(torch.as_tensor(kernel_points, dtype=torch.float32), requires_grad=False)

    def forward(self, q_pts, s_pts, neighb_inds, x):
        return kpconv_reference_ops(q_pts, s_pts, neighb_inds.long(), x, self.weights, self.kernel_points,
                                    self.KP_extent)
