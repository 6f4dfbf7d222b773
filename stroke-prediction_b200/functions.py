"""autograd nodes for the loss terms and the latent interpolation (all arithmetic in libstroke_b200.so).

Reference semantics: BatchDiceLoss (common/metrics.py:16-28); hinge ``mean(abs(d) - d)`` and L1 ``mean(abs(a - b))``
(learner/CaeReconstructionLearner.py:59-62,68); ``Enc3D._interpolate`` (common/model/Cae3D.py:78-89).
"""
import torch

from . import ops


def _flat_pair(a, b):
    """Reductions over two same-shaped tensors only need both to share one dense element order."""
    if a.shape != b.shape:
        raise RuntimeError("shape mismatch %s vs %s" % (tuple(a.shape), tuple(b.shape)))
    if a.dtype != torch.float32:
        a = a.float()
    if b.dtype != torch.float32:
        b = b.float()
    if a.dim() == 5 and ops.is_ndhwc(a) and ops.is_ndhwc(b):
        return a, b
    if a.is_contiguous() and b.is_contiguous():
        return a, b
    if a.dim() == 5 and ops.is_ndhwc(a):
        return a, ops.as_vol(b.contiguous()) if b.is_cuda else b
    if a.dim() == 5 and ops.is_ndhwc(b):
        return ops.as_vol(a.contiguous()), b
    return a.contiguous(), b.contiguous()


class DiceTermFunction(torch.autograd.Function):
    """1 - w * (2 sum(o t) + eps) / (sum(o^2) + sum(t^2) + eps) over the whole batch; gradient to `o` only."""

    @staticmethod
    def forward(ctx, o, t, w, eps):
        o, t = _flat_pair(o, t)
        sums = ops.dice_sums(o, t)
        ctx.save_for_backward(o, t, sums)
        ctx.w, ctx.eps = w, eps
        return ops.dice_loss(sums, w, eps)

    @staticmethod
    def backward(ctx, g):
        o, t, sums = ctx.saved_tensors
        go = torch.empty_like(o)
        g = g.contiguous().float()
        ops.dice_bwd(o, t, sums, ctx.w, ctx.eps, g, 1.0, go, False)
        return go, None, None, None


class AbsDiffMeanFunction(torch.autograd.Function):
    """mode 0: mean(|a-b| - (a-b)) (monotonicity hinge); mode 1: mean(|a-b|) (latent L1)."""

    @staticmethod
    def forward(ctx, a, b, mode):
        a, b = _flat_pair(a, b)
        ctx.save_for_backward(a, b)
        ctx.mode = mode
        return ops.absdiff_mean(a, b, mode)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        ga = torch.empty_like(a) if need_a else None
        gb = torch.empty_like(b) if need_b else None
        if need_a or need_b:
            ops.absdiff_bwd(a, b, ctx.mode, g.contiguous().float(), 1.0, ga, False, gb, False)
        return ga, gb, None


class LatentInterpFunction(torch.autograd.Function):
    """z_c + s * (z_p - z_c) with one step s per sample."""

    @staticmethod
    def forward(ctx, zc, zp, step):
        zc, zp = _flat_pair(zc, zp)
        s = step.reshape(-1).contiguous().float()
        if s.numel() != zc.shape[0]:
            raise RuntimeError("one interpolation step per sample expected, got %d for batch %d" % (s.numel(), zc.shape[0]))
        ctx.save_for_backward(zc, zp, s)
        ctx.step_shape = step.shape
        return ops.latent_interp_fwd(zc, zp, s)

    @staticmethod
    def backward(ctx, g):
        zc, zp, s = ctx.saved_tensors
        g2, _ = _flat_pair(g, zc)
        nc, np_, ns = ctx.needs_input_grad
        dzc, dzp, ds = ops.latent_interp_bwd(g2, zc, zp, s, nc, np_, ns)
        if ds is not None:
            ds = ds.reshape(ctx.step_shape)
        return dzc, dzp, ds


def dice_term(o, t, w=1.0, eps=1e-7):
    return DiceTermFunction.apply(o, t, float(w), float(eps))


def hinge_mean(a, b):
    return AbsDiffMeanFunction.apply(a, b, 0)


def l1_mean(a, b):
    return AbsDiffMeanFunction.apply(a, b, 1)


def latent_interp(zc, zp, step):
    return LatentInterpFunction.apply(zc, zp, step)
