"""FusedAdam — torch.optim.Adam semantics (L2-coupled weight decay, bias correction, per-group betas read at every
step) with ONE kernel launch for all parameter tensors (sp_adam_multi).

Replaces ``torch.optim.Adam`` as configured by the reference (train_shape_reconstruction.py:40,
train_unet_segmentation.py:32) and stepped in ``Learner.train_batch`` (Learner.py:120-122).  State layout
(``step``, ``exp_avg``, ``exp_avg_sq`` per parameter) and ``state_dict()`` format are torch's, so ``.optim`` files of
the reference load (Learner.py:96-110) and ``adapt_betas`` (CaeReconstructionLearner.py:28-40) keeps working on
``param_groups``.
"""
import torch
from torch.optim.optimizer import Optimizer

from . import engine, ops


class FusedAdam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0):
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0:
            raise ValueError("invalid hyper-parameter")
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError("invalid betas %r" % (betas,))
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._tables = {}
        self._sink = None
        self.grad_scale = 1.0       # multiplied into every gradient (1/world_size after a sum all-reduce)
        self.fuse_zero_grad = False  # clear gradients inside the update kernel (saves the separate memset pass)
        self.launches = 0

    @classmethod
    def from_torch(cls, opt):
        """Adopt the param groups and state of an existing torch.optim.Adam (the scripts construct that class)."""
        if isinstance(opt, cls):
            return opt
        if not isinstance(opt, torch.optim.Adam):
            raise TypeError("FusedAdam.from_torch expects torch.optim.Adam, got %r" % type(opt))
        for g in opt.param_groups:
            if g.get('amsgrad') or g.get('maximize'):
                raise ValueError("amsgrad / maximize are not used by the reference and not supported")
        new = cls.__new__(cls)
        Optimizer.__init__(new, opt.param_groups, dict(opt.defaults))
        new.param_groups = opt.param_groups   # share: adapt_betas / schedulers mutate these dicts
        new.state = opt.state
        new.defaults = opt.defaults
        new._tables = {}
        new._sink = None
        new.grad_scale = 1.0
        new.fuse_zero_grad = False
        new.launches = 0
        return new

    # ------------------------------------------------------------------------------------------ gradient sink
    def attach_grad_sink(self):
        """Route all gradients of this optimizer's parameters into one persistent flat buffer (engine.GradSink)."""
        if getattr(self, '_sink', None) is None:
            params = [p for g in self.param_groups for p in g['params'] if p.requires_grad]
            self._sink = engine.register_grad_sink(engine.GradSink(params))
        return self._sink

    def detach_grad_sink(self):
        sink = getattr(self, '_sink', None)
        if sink is not None:
            engine.unregister_grad_sink(sink)
            sink.release()
            self._sink = None

    def __del__(self):
        # a dropped optimizer must not keep routing its parameters' gradients into a buffer nobody reads
        try:
            sink = self.__dict__.get('_sink')
            if sink is not None:
                engine.unregister_grad_sink(sink)
        except Exception:
            pass

    def load_state_dict(self, state_dict):
        """torch's ``load_state_dict`` REPLACES ``param_groups`` and ``state`` by new objects.  The reference loads into
        the very optimizer its scheduler was built on (Learner.py:96-103), and ``from_torch`` shares these objects with
        the script's ``torch.optim.Adam`` so that ``MultiStepLR`` / ``adapt_betas`` keep acting on the optimizer that
        steps: load, then move the loaded content back INTO the original objects."""
        groups, state = self.param_groups, self.state
        super().load_state_dict(state_dict)
        new_groups, new_state = self.param_groups, self.state
        if new_groups is not groups:
            for old, new in zip(groups, new_groups):
                if old is not new:
                    keep = {k: old[k] for k in ('initial_lr',) if k in old and k not in new}
                    old.clear()
                    old.update(new)
                    old.update(keep)
            self.param_groups = groups
        if new_state is not state:
            state.clear()
            state.update(new_state)
            self.state = state
        for st in self.state.values():
            for k in ('exp_avg', 'exp_avg_sq'):
                if k in st and not st[k].is_contiguous():
                    st[k] = st[k].contiguous()
        self._tables = {}

    def zero_grad(self, set_to_none=True):
        sink = getattr(self, '_sink', None)
        if sink is None:
            return super().zero_grad(set_to_none=set_to_none)
        if sink.dirty or not self.fuse_zero_grad:
            sink.zero()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            entries = []
            steps = set()
            for p in group['params']:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("FusedAdam: CUDA parameters only — there is no CPU path")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("FusedAdam: fp32 master parameters expected")
                st = self.state[p]
                if len(st) == 0:
                    st['step'] = torch.tensor(0.0, dtype=torch.float32)
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                g = p.grad
                if not p.is_contiguous() or not g.is_contiguous():
                    raise RuntimeError("FusedAdam: contiguous parameters and gradients expected")
                st['step'] += 1
                steps.add(int(st['step']))
                entries.append((p, g, st['exp_avg'], st['exp_avg_sq']))
            if not entries:
                continue
            if len(steps) != 1:
                # parameters that joined later (e.g. unfrozen): one launch per distinct step count
                by_step = {}
                for e in entries:
                    by_step.setdefault(int(self.state[e[0]]['step']), []).append(e)
            else:
                by_step = {steps.pop(): entries}
            b1, b2 = group['betas']
            for t, ents in by_step.items():
                # the pointer table does not depend on the step count: key it on the pointer set (bounded: one entry per
                # distinct set of tensors that ever stepped together), never on t
                ptrs = tuple((p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel()) for p, g, m, v in ents)
                hit = self._tables.get((gi, ptrs))
                if hit is None:
                    if len(self._tables) >= 16:
                        self._tables.clear()
                    table, blocks = ops.make_adam_table(ents, ents[0][0].device)
                    hit = (ptrs, table, blocks)
                    self._tables[(gi, ptrs)] = hit
                ops.adam_multi(hit[1], len(ents), hit[2], group['lr'], b1, b2, group['eps'], group['weight_decay'],
                               t, self.grad_scale, self.fuse_zero_grad)
                self.launches += 1
        sink = getattr(self, '_sink', None)
        if sink is not None and self.fuse_zero_grad:
            sink.dirty = False          # the update kernel cleared the gradients it consumed
        engine.bump_weights_epoch()
        return loss
