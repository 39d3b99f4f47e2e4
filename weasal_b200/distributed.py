"""Multi-GPU plumbing for the hot path (one process per GPU, torch.distributed).

The path shards over independent units: a batch is a stack of independent spheres and batches are
independent given the weights, so
  * training is data parallel: every rank draws its own spheres and builds its own pyramid on its GPU; the only
    exchange is ONE gradient all-reduce per step (NCCL over NVLink on GPUs, gloo in the CPU tests);
  * test-time voting shards sphere centres round-robin over ranks; every rank accumulates (sum of weighted
    probabilities, sum of weights) for the points its spheres touch and the accumulators are all-reduced once at
    the end. (The reference's in-place EMA ``0.95*old + 0.05*new``, utils/tester_PseudoLabel.py:194, depends on the
    order spheres are visited in, so a sharded run cannot reproduce it bit-for-bit; the sum/weight form is
    order-independent.)
No data-path collective exists inside the pyramid or KPConv kernels.
"""
import torch
import torch.distributed as dist


class GradAllReducer:
    """Flat-buffer gradient averaging. Parameters without a gradient (the reference's BatchNorm layers are
    identities on 2-D features and never receive one) contribute zeros and stay ``grad=None``."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(n, dtype=torch.float32, device=ref.device)
        self.views = []
        o = 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()

    def bytes(self):
        return self.flat.numel() * 4

    @torch.no_grad()
    def step(self):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        had = []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
                had.append(False)
            else:
                v.copy_(p.grad)
                had.append(True)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(dist.get_world_size(self.group))
        for p, v, h in zip(self.params, self.views, had):
            if h:
                p.grad.copy_(v)


def shard_indices(n_items, rank, world_size):
    """Round-robin shard of sphere (or batch) indices: rank r takes r, r + W, r + 2W, ..."""
    return list(range(rank, n_items, world_size))


class VoteAccumulator:
    """Order-independent vote accumulation for sharded test-time inference."""

    def __init__(self, n_points, n_classes, device):
        self.sum = torch.zeros((n_points, n_classes), dtype=torch.float32, device=device)
        self.weight = torch.zeros((n_points,), dtype=torch.float32, device=device)

    @torch.no_grad()
    def add(self, point_inds, probs, w=1.0):
        self.sum.index_add_(0, point_inds, probs * w)
        self.weight.index_add_(0, point_inds, torch.full((len(point_inds),), float(w), device=self.weight.device))

    @torch.no_grad()
    def reduce(self, group=None):
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.sum, group=group)
            dist.all_reduce(self.weight, group=group)
        return self.sum / self.weight.clamp_min(1e-12).unsqueeze(1)
