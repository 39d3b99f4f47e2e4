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
    """Flat-buffer gradient averaging: one concatenation, ONE all-reduce, one fused copy back per step.
    Parameters without a gradient (the reference's BatchNorm layers are identities on 2-D features and never receive
    one) are skipped and stay ``grad=None``; every rank runs the same network, so every rank
    skips the same ones."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group

    def bytes(self):
        return sum(p.numel() for p in self.params) * 4

    @torch.no_grad()
    def step(self):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        grads = [p.grad for p in self.params if p.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.div_(dist.get_world_size(self.group))
        torch._foreach_copy_(grads, [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in grads]), grads)])


    @torch.no_grad()
    def allreduce_flat(self, flat):
        """Average a flat gradient buffer in place: ONE collective, no copies (the captured training step gathers the
        gradients into ``flat`` inside its graph and reads them back from there)."""
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        if dist.get_backend(self.group) == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        else:  # (gloo, the CPU tests: no AVG)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(dist.get_world_size(self.group))

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
