"""Data-parallel plumbing: one process per GPU, batch sharded across ranks, weights replicated.

The reference is single-process (no collective anywhere); the only exchange batch sharding needs is the gradient
all-reduce between ``loss.backward()`` and ``optimizer.step()`` (Learner.py:121-122).  Because every trainable
parameter's gradient lives in ONE flat fp32 buffer (engine.GradSink), that is a single ``all_reduce`` call — NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests of this logic — and the 1/world_size averaging is folded into the
fused Adam kernel (``grad_scale``).  BatchNorm statistics and Dice sums stay local per shard ("DDP semantics"): the
result equals the reference run on each shard with averaged gradients.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank` (the first n_items % world ranks get one extra item)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch, rank, world):
    """Slice every tensor of a reference batch dict along dim 0."""
    n = next(v.shape[0] for v in batch.values() if torch.is_tensor(v))
    lo, hi = shard_range(n, rank, world)
    return {k: (v[lo:hi] if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n else v) for k, v in batch.items()}


def allreduce_flat_(flat, group=None):
    """Sum `flat` over the group in place (no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def broadcast_parameters(module, src=0, group=None):
    """Make every rank start from rank `src`'s parameters and buffers."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


class GradientAllReduce:
    """Callable placed between backward and the optimizer step."""

    def __init__(self, optimizer, group=None):
        self.optimizer = optimizer
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        sink = optimizer.attach_grad_sink() if hasattr(optimizer, "attach_grad_sink") else None
        self.flat = sink.flat if sink is not None else None
        optimizer.grad_scale = 1.0 / self.world
        self.calls = 0

    def __call__(self):
        if self.world == 1:
            return
        if self.flat is not None:
            allreduce_flat_(self.flat, self.group)
        else:
            for g in self.optimizer.param_groups:
                for p in g["params"]:
                    if p.grad is not None:
                        dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group)
        self.calls += 1
