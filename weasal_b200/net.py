"""Caller harness for the hot path: a KPFCNN-shaped segmentation network (the consumer of KPConv and of the
pyramid), parameterised by the KPConv class so the same network runs on the B200 operator or on the CPU
restatement of the reference operator.

It mirrors the structure the reference assembles in models/architectures.py:196-290 from models/blocks.py:387-755 —
'simple' (KPConv -> LeakyReLU 0.1), 'resnetb' / 'resnetb_strided' (unary -> KPConv -> unary, shortcut with max-pool
when strided), 'nearest_upsample' + 'unary' decoder with skip concatenation, two-layer head — including the
reference's quirk that BatchNormBlock is the identity on 2-D features when use_batch_norm=True (blocks.py:453-463), so
conv blocks have no normalisation and only the head blocks carry a bias. Feature widths follow
architectures.py:214-251. It is bench / test support, not part of the drop-in boundary.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


import os

# Unary blocks up to this many channels (in and out) run on the fused tcgen05 linear kernels; wider ones are plain large
# GEMMs where the library kernel wins. Measured in the training step (bench.py, ms/step Vaihingen / DALES): 64: 3.64 /
# 4.07, 128: 3.59 / 3.85, 256: 3.58 / 3.88, 512: 3.66 / 3.98, all: 3.98 / 4.67.
UNARY_FUSED_MAX_CHANNELS = int(os.environ.get("WEASAL_UNARY_MAX_C", "256"))
FORK_SHORTCUT = os.environ.get("WEASAL_FORK_SHORTCUT", "1") != "0"


def max_pool(x, inds, width=None):
    """blocks.py:93-112: shadow row is zeros, so shadow entries contribute 0 to the max. ``width``: the true width of a
    fixed-width matrix (int32 device scalar, static-shape batches), columns beyond it are ignored."""
    if x.is_cuda:
        from . import ops
        return ops.max_pool(x, inds, width)
    xp = torch.cat((x, torch.zeros_like(x[:1, :])), 0)  # CPU: the reference's own formulation (reference arm only)
    idx = inds.unsqueeze(2).expand(-1, -1, xp.shape[1])
    return xp.unsqueeze(1).expand(-1, inds.shape[1], -1).gather(0, idx).max(dim=1)[0]


def closest_pool(x, inds):
    """blocks.py:77-90: nearest upsampling through the first (closest) neighbour column."""
    if x.is_cuda:
        from . import ops
        return ops.closest_pool(x, inds)
    xp = torch.cat((x, torch.zeros_like(x[:1, :])), 0)
    idx = inds[:, 0].unsqueeze(1).expand(-1, xp.shape[1])
    return xp.gather(0, idx)


class Unary(nn.Module):
    def __init__(self, cin, cout, bias=False, relu=True):
        super().__init__()
        self.mlp = nn.Linear(cin, cout, bias=False)
        self.bias = nn.Parameter(torch.zeros(cout)) if bias else None
        self.relu = relu

    def forward(self, x):
        # Narrow layers (the shallow, many-row ones): Linear (+ bias) + LeakyReLU as one tcgen05 kernel each way
        # (kp_linear_*_dev). Wide layers are plain large GEMMs, where the library GEMM is the faster tool (measured:
        # tools/bench_linear.py, profiles/).
        if x.is_cuda and max(self.mlp.in_features, self.mlp.out_features) <= UNARY_FUSED_MAX_CHANNELS:
            from . import ops
            return ops.linear_act(x, self.mlp.weight, self.bias, 0.1 if self.relu else 1.0)
        x = self.mlp(x)  # CPU: the reference's own formulation (reference arm only)
        if self.bias is not None:
            x = x + self.bias
        return F.leaky_relu(x, 0.1) if self.relu else x


def fused_linear_weights(net):
    """``nn.Linear`` weights of the unary blocks that run on the fused tcgen05 linear kernels (for plan.WeightPacker)."""
    return [m.mlp.weight for m in net.modules()
            if isinstance(m, Unary) and max(m.mlp.in_features, m.mlp.out_features) <= UNARY_FUSED_MAX_CHANNELS]


class ConvBlock(nn.Module):
    """'simple', 'resnetb' or 'resnetb_strided' block at pyramid level `layer`."""

    def __init__(self, kind, cin, cout, radius, layer, cfg, conv_cls):
        super().__init__()
        self.kind, self.layer, self.strided = kind, layer, 'strided' in kind
        extent = radius * cfg["KP_extent"] / cfg["conv_radius"]
        if kind == 'simple':
            self.conv = conv_cls(cfg["num_kernel_points"], 3, cin, cout // 2, extent, radius)
        else:
            mid = cout // 4
            self.unary1 = Unary(cin, mid) if cin != mid else nn.Identity()
            self.conv = conv_cls(cfg["num_kernel_points"], 3, mid, mid, extent, radius)
            self.unary2 = Unary(mid, cout, relu=False)
            self.shortcut = Unary(cin, cout, relu=False) if cin != cout else nn.Identity()

    def forward(self, x, batch):
        l = self.layer
        if self.strided:
            q, s, idx = batch.points[l + 1], batch.points[l], batch.pools[l]
        else:
            q, s, idx = batch.points[l], batch.points[l], batch.neighbors[l]
        if self.kind == 'simple':
            return F.leaky_relu(self.conv(q, s, idx, x), 0.1)
        widths = getattr(batch, "pool_widths", None)

        def shortcut():
            return self.shortcut(max_pool(x, idx, widths[l] if widths is not None else None) if self.strided else x)

        # The shortcut branch (max-pool and / or a Linear) does not depend on the conv branch: it runs on a second stream
        # beside unary1 -> KPConv -> unary2, forward and (autograd keeps an op's backward on its forward's stream)
        # backward; inside a captured step this becomes a fork / join of the graph. The deep layers' kernels are a few
        # CTAs each, so the two branches share the SMs instead of queueing.
        # (only while a step is being captured: launched eagerly, the extra stream hand-overs cost host time instead)
        fork = FORK_SHORTCUT and x.is_cuda and (self.strided or not isinstance(self.shortcut, nn.Identity))
        if fork:
            from . import ops
            fork = ops.forking_for_capture()
        if fork:
            from . import ops
            cur, side = torch.cuda.current_stream(x.device), ops._side_stream(x.device, 1)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                sc = shortcut()
            x.record_stream(side)
        y = self.unary2(F.leaky_relu(self.conv(q, s, idx, self.unary1(x)), 0.1))
        if fork:
            cur.wait_stream(side)
            sc.record_stream(cur)
        else:
            sc = shortcut()
        return F.leaky_relu(y + sc, 0.1)


class KPFCNNHarness(nn.Module):
    def __init__(self, cfg, conv_cls):
        super().__init__()
        arch = cfg["architecture"]
        layer, r = 0, cfg["first_subsampling_dl"] * cfg["conv_radius"]
        cin, cout = cfg["in_features_dim"], cfg["first_features_dim"]
        self.encoder, self.enc_kinds, skip_dims = nn.ModuleList(), [], []
        self.skip_at = []
        for i, blk in enumerate(arch):
            if any(t in blk for t in ('pool', 'strided', 'upsample', 'global')):
                self.skip_at.append(i)
                skip_dims.append(cin)
            if 'upsample' in blk:
                break
            self.encoder.append(ConvBlock(blk, cin, cout, r, layer, cfg, conv_cls))
            cin = cout // 2 if 'simple' in blk else cout
            if 'strided' in blk:
                layer += 1
                r *= 2
                cout *= 2
        start = next(i for i, blk in enumerate(arch) if 'upsample' in blk)
        self.decoder, self.dec_kinds, self.concat_at = nn.ModuleList(), [], []
        for j, blk in enumerate(arch[start:]):
            if j > 0 and 'upsample' in arch[start + j - 1]:
                cin += skip_dims[layer]
                self.concat_at.append(j)
            if 'upsample' in blk:
                self.decoder.append(nn.Identity())
                self.dec_kinds.append(('up', layer))
                layer -= 1
                r *= 0.5
                cin_next = cin
            else:
                self.decoder.append(Unary(cin, cout))
                self.dec_kinds.append(('unary', layer))
                cin_next = cout
            cin = cin_next
            if 'upsample' in blk:
                cout = cout // 2
        self.head_mlp = Unary(cout, cfg["first_features_dim"], bias=True)
        self.head_softmax = Unary(cfg["first_features_dim"], cfg["num_classes"], bias=True)
        self.dropout = cfg.get("dropout", 0.5)

    def forward(self, batch):
        x = batch.features
        skips = []
        for i, blk in enumerate(self.encoder):
            if i in self.skip_at:
                skips.append(x)
            x = blk(x, batch)
        for j, (blk, (kind, layer)) in enumerate(zip(self.decoder, self.dec_kinds)):
            if j in self.concat_at:
                x = torch.cat([x, skips.pop()], dim=1)
            x = closest_pool(x, batch.upsamples[layer - 1]) if kind == 'up' else blk(x)
        if self.dropout and self.training:
            x = F.dropout(x, self.dropout)
        return self.head_softmax(self.head_mlp(x))


VAIHINGEN_PL_ARCH = ['simple', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb_strided', 'resnetb',
                     'resnetb_strided', 'resnetb', 'resnetb_strided', 'resnetb', 'nearest_upsample', 'unary',
                     'nearest_upsample', 'unary', 'nearest_upsample', 'unary', 'nearest_upsample', 'unary']


def net_config(name="vaihingen_pl"):
    """Constants of the reference's entry scripts (train_Vaihingen3D_PseudoLabel.py:70-121,
    train_DALES_PseudoLabel.py:98-121)."""
    base = dict(architecture=VAIHINGEN_PL_ARCH, num_kernel_points=15, conv_radius=2.5, deform_radius=6.0,
                KP_extent=1.0, dropout=0.5)
    if name == "vaihingen_pl":
        base.update(first_subsampling_dl=0.24, in_features_dim=4, first_features_dim=64, num_classes=9)
    elif name == "dales_pl":
        base.update(first_subsampling_dl=0.4, in_features_dim=3, first_features_dim=128, num_classes=8)
    else:
        raise KeyError(name)
    return base


class CfgView:
    """attribute view of a config dict (what pyramid.segmentation_inputs expects)"""

    def __init__(self, d):
        self.__dict__.update(d)
