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
    """Callable placed between backward and the optimizer step (Learner.py:121-122).

    The gradient exchange is OVERLAPPED with the backward pass: the engine reports every layer plan whose parameter
    gradients are final (``engine.plan_backward_hooks``: the decoder's plan fires before the encoder's backward has started,
    every U-Net block fires as the backward walks up the network), and the slice of the flat gradient buffer that holds
    that plan's parameters is all-reduced asynchronously right away (NCCL orders it after the kernels already queued on the
    compute stream and runs it next to the remaining backward kernels).  ``__call__`` reduces whatever was not covered,
    then makes the compute stream wait for all pieces.  Reducing a buffer in slices is the same elementwise sum as reducing
    it at once, so results do not depend on the overlap.  ``overlap=False`` restores the single blocking all-reduce."""

    def __init__(self, optimizer, group=None, overlap=True):
        self.optimizer = optimizer
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        sink = optimizer.attach_grad_sink() if hasattr(optimizer, "attach_grad_sink") else None
        self.sink = sink
        self.flat = sink.flat if sink is not None else None
        optimizer.grad_scale = 1.0 / self.world
        self.calls = 0
        self.overlap = bool(overlap) and self.flat is not None
        self._done = []          # [lo, hi) element ranges of the flat buffer already handed to the collective this step
        self._work = []
        self.early_elements = 0  # elements reduced before __call__ in the last step (diagnostics / tests)
        self._hook = None
        if self.overlap:
            from . import engine
            self._hook = self.plan_ready
            engine.plan_backward_hooks.append(self._hook)

    def close(self):
        if self._hook is not None:
            from . import engine
            if self._hook in engine.plan_backward_hooks:
                engine.plan_backward_hooks.remove(self._hook)
            self._hook = None

    # ------------------------------------------------------------------------------------------ buckets
    def plan_range(self, plan):
        """[lo, hi) of the flat buffer covered by `plan`'s parameters, or None if they are not ours / not one contiguous run."""
        offs = self.sink._offsets
        spans = []
        for p in plan.params():
            o = offs.get(id(p))
            if o is None:
                if p.requires_grad:
                    return None
                continue
            spans.append((o, o + (p.numel() + 3) // 4 * 4))
        if not spans:
            return None
        spans.sort()
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            if b0 != a1:
                return None
        return spans[0][0], min(spans[-1][1], self.flat.numel())

    def plan_ready(self, plan):
        if self.world == 1 or not self.overlap or not self.sink.params:
            return
        r = self.plan_range(plan)
        if r is not None:
            self.bucket_ready(r[0], r[1])

    def bucket_ready(self, lo, hi):
        """Start the all-reduce of flat[lo:hi) (skipping parts already started this step)."""
        for a, b in self._uncovered(lo, hi):
            self._work.append(dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self._done.append((a, b))
            self.early_elements += b - a

    def _uncovered(self, lo, hi):
        gaps, cur = [], lo
        for a, b in sorted(self._done):
            if b <= cur or a >= hi:
                continue
            if a > cur:
                gaps.append((cur, min(a, hi)))
            cur = max(cur, b)
            if cur >= hi:
                break
        if cur < hi:
            gaps.append((cur, hi))
        return gaps

    def __call__(self):
        if self.world == 1:
            return
        if self.flat is not None:
            early = self.early_elements
            for a, b in self._uncovered(0, self.flat.numel()):
                self._work.append(dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            for w in self._work:
                w.wait()                 # stream-ordered for NCCL (the compute stream waits), blocking for gloo
            self._work, self._done = [], []
            self.last_early_elements, self.early_elements = early, 0
        else:
            for g in self.optimizer.param_groups:
                for p in g["params"]:
                    if p.grad is not None:
                        dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group)
        self.calls += 1
