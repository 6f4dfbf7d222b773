"""Layer plans and hand-written forward/backward chains over the sm_100a kernels.

A reference ``nn.Sequential`` on the hot path is a chain of pre-norm units ``[BatchNorm3d] -> Conv3d |
ConvTranspose3d -> [ELU | LeakyReLU | Sigmoid]`` (Cae3D.py:39-76,176-220; Unet3D.py:17-24,49-54).  ``SeqPlan`` parses
such a Sequential (whose modules keep owning the parameters, so ``state_dict`` names/shapes stay those of the
reference) into fused units and runs them through the C-ABI:

forward per unit   : bn_stats -> bn_finalize (scale/shift)  ->  corr / corrT with BN applied while staging the
                     source and bias + activation fused in the epilogue.  Only the activation output is stored.
backward per unit  : wgrad (BN re-applied on the fly), bias gradient (column sums fused into the kernel that wrote gz), corrT / corr for the gradient w.r.t. the BN output,
                     bn_bwd_reduce -> bn_bwd_finalize (dgamma, dbeta, coefficients) -> bn_act_bwd_apply, which fuses
                     the BN backward with the derivative of the *previous* unit's activation.

``SeqFunction`` / ``UnetFunction`` expose the chains to autograd as single nodes.
"""
import weakref

import torch
import torch.nn as nn

from . import ops
from .ops import ACT_ELU, ACT_LEAKY, ACT_NONE, ACT_SIGMOID

_weights_epoch = 0
# Called as hook(plan) when ALL backward passes of a SeqPlan belonging to the current step are done, i.e. its parameters'
# gradients are final (as many seq_backward calls as gradient-tracked seq_forward calls).  parallel.GradientAllReduce uses
# it to start the all-reduce of that plan's slice of the flat gradient buffer while the rest of the backward still runs.
plan_backward_hooks = []
DEBUG_GZ = None     # diagnostics only (tools/diag_chain.py): a list receiving (plan, unit index, dL/d(conv output))
DEBUG_ACTS = None   # diagnostics only: a list receiving (plan, [input, unit outputs...]) per forward pass


def bump_weights_epoch():
    """Called by the fused optimizer after it has rewritten parameters behind autograd's back."""
    global _weights_epoch
    _weights_epoch += 1


class GradSink:
    """Persistent flat gradient buffer.

    Every trainable parameter's ``.grad`` is a view into one fp32 buffer; the backward chains accumulate weight /
    bias / BN-affine gradients straight into these views (kernel-side ``beta = 1``) instead of returning fresh
    tensors to autograd's AccumulateGrad.  Consequences: the CAE's 3 encoder / 4 decoder passes sum their parameter
    gradients inside the wgrad kernels, gradient pointers are stable (the fused Adam pointer table is built once),
    clearing is one memset (or fused into the Adam kernel), and data-parallel training all-reduces ONE tensor.
    """

    def __init__(self, params):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("GradSink: no trainable parameters")
        dev = params[0].device
        offs, total = [], 0
        for p in params:
            if not p.is_cuda or p.dtype != torch.float32:
                raise RuntimeError("GradSink: fp32 CUDA parameters expected")
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4          # keep every view 16-byte aligned
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.params = params
        self._views = {}
        self._offsets = {id(p): o for p, o in zip(params, offs)}
        for p in params:
            self._attach(p)
        self.dirty = False

    def _attach(self, p):
        o = self._offsets[id(p)]
        v = self.flat[o:o + p.numel()].view(p.shape)
        p.grad = v
        self._views[id(p)] = v
        return v

    def owns(self, p):
        return id(p) in self._offsets

    def view(self, p):
        """The gradient view of `p` inside ``flat`` (None if `p` is not ours).  Self-healing: ``zero_grad(set_to_none=
        True)`` drops ``p.grad``, and ``Module._apply`` (``model.cpu()`` / ``.to(device)``) re-binds the STORAGE of the
        very tensor object we handed out (``param.grad.data = fn(grad)``), so identity alone proves nothing — the
        view's address is compared with its slot of the flat buffer and the view is rebuilt when they differ."""
        o = self._offsets.get(id(p))
        if o is None:
            return None
        v = self._views[id(p)]
        if (p.grad is not v or v.device != self.flat.device
                or v.data_ptr() != self.flat.data_ptr() + 4 * o):
            v = self._attach(p)
        return v

    def reattach(self):
        """Re-validate every view (after anything that may have moved the module)."""
        for p in self.params:
            self.view(p)

    def release(self):
        """Detach: parameters get ordinary (None) gradients again."""
        for p in self.params:
            if p.grad is self._views.get(id(p)):
                p.grad = None
        self._views.clear()
        self._offsets.clear()
        self.params = []

    def zero(self):
        self.flat.zero_()
        self.dirty = False


# id(parameter) -> (weak reference to the parameter, the sink that owns its gradient).  Keyed by id because tensors
# compare element-wise; the weak reference's callback drops the entry with the parameter (and with it the last reference
# to the sink's flat buffer).  The LATEST sink attached to a parameter wins, e.g. when a second Learner / optimizer is
# built over a model a first one trained (the reference trains the CAE, then CaeStepLearner or CaePredictionLearner on
# the same modules).
_sink_of = {}


def _owner(p):
    hit = _sink_of.get(id(p))
    if hit is None or hit[0]() is not p:
        return None
    return hit[1]


def register_grad_sink(sink):
    for p in sink.params:
        old = _owner(p)
        if old is not None and old is not sink:
            old._offsets.pop(id(p), None)      # the previous owner no longer serves this parameter
            old._views.pop(id(p), None)
            old.params = [q for q in old.params if q is not p]
        key = id(p)
        _sink_of[key] = (weakref.ref(p, lambda _r, key=key: _sink_of.pop(key, None)), sink)
    return sink


def unregister_grad_sink(sink):
    for p in list(sink.params):
        if _owner(p) is sink:
            del _sink_of[id(p)]


def _sink_view(p):
    s = _owner(p)
    if s is None:
        return None
    v = s.view(p)
    if v is not None:
        s.dirty = True
    return v


def _triple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


def _act_of(m):
    if m is None:
        return ACT_NONE, 0.0
    if isinstance(m, nn.ELU):
        return ACT_ELU, float(m.alpha)
    if isinstance(m, nn.LeakyReLU):
        return ACT_LEAKY, float(m.negative_slope)
    if isinstance(m, nn.Sigmoid):
        return ACT_SIGMOID, 0.0
    raise TypeError("unsupported activation %r" % (m,))


class FusedUnit:
    """One [BN] -> conv/convT -> [act] unit."""

    def __init__(self, bn, conv, act_module):
        self.bn = bn
        self.conv = conv
        self.transposed = isinstance(conv, nn.ConvTranspose3d)
        k, s, p = _triple(conv.kernel_size), _triple(conv.stride), _triple(conv.padding)
        if not (k[0] == k[1] == k[2] and s[0] == s[1] == s[2]):
            raise ValueError("only cubic kernels / isotropic strides are on the hot path")
        if _triple(conv.dilation) != (1, 1, 1) or conv.groups != 1:
            raise ValueError("dilation/groups are not used by the reference and not supported")
        if self.transposed and _triple(conv.output_padding) != (0, 0, 0):
            raise ValueError("output_padding is not used by the reference and not supported")
        self.k, self.s, self.pad = k[0], s[0], p
        self.act, self.alpha = _act_of(act_module)
        self.cin, self.cout = conv.in_channels, conv.out_channels
        self._packed = {}

    def out_size(self, size):
        if self.transposed:
            return tuple((i - 1) * self.s - 2 * p + self.k for i, p in zip(size, self.pad))
        return tuple((i + 2 * p - self.k) // self.s + 1 for i, p in zip(size, self.pad))

    def desc(self, N, in_size, act=None, alpha=None):
        """Correlation-geometry descriptor; conv: I-side = input, convT: O-side = input."""
        out_size = self.out_size(in_size)
        a = self.act if act is None else act
        al = self.alpha if alpha is None else alpha
        if self.transposed:
            return ops.conv_desc(N, out_size, self.cout, in_size, self.cin, self.k, self.s, self.pad, a, al)
        return ops.conv_desc(N, in_size, self.cin, out_size, self.cout, self.k, self.s, self.pad, a, al)

    @staticmethod
    def _geom(d, which):
        return (which, d.N, d.Di, d.Hi, d.Wi, d.Ci, d.ldi, d.Do, d.Ho, d.Wo, d.Co, d.ldo, d.k, d.s, d.pd, d.ph, d.pw)

    def has_pack(self, d, which):
        """True when packed(d, which) would be served from the cache."""
        w = self.conv.weight
        hit = self._packed.get(self._geom(d, which))
        return hit is not None and hit[0] == (w._version, w.data_ptr(), _weights_epoch)

    def packed(self, d, which):
        """Packed weights for direction `which` under geometry `d`.  The layout sp_pack_weights writes (FFMA / GEMM /
        tensor-core image, and the buffer size) depends on the tier the GEOMETRY selects, so the cache is keyed on the
        geometry as well as on the parameter version: the same unchanged weights at another volume size (validation,
        Tester, visualisation) get their own pack."""
        w = self.conv.weight
        geom = self._geom(d, which)
        key = (w._version, w.data_ptr(), _weights_epoch)
        hit = self._packed.get(geom)
        if hit is not None and hit[0] == key:
            return hit[1]
        if len(self._packed) >= 8:          # bounded: drop packs of geometries / versions no longer in use
            for g in [g for g, h in self._packed.items() if h[0] != key]:
                del self._packed[g]
            if len(self._packed) >= 8:
                self._packed.clear()
        wp = ops.pack_weights(d, which, w)
        self._packed[geom] = (key, wp)
        return wp

    def params(self):
        ps = []
        if self.bn is not None and self.bn.affine:
            ps += [self.bn.weight, self.bn.bias]
        ps.append(self.conv.weight)
        if self.conv.bias is not None:
            ps.append(self.conv.bias)
        return ps


class SeqPlan:
    def __init__(self, seq):
        mods = list(seq.children()) if isinstance(seq, nn.Sequential) else list(seq)
        self.units = []
        i = 0
        while i < len(mods):
            bn = None
            if isinstance(mods[i], nn.BatchNorm3d):
                bn = mods[i]
                i += 1
            if i >= len(mods) or not isinstance(mods[i], (nn.Conv3d, nn.ConvTranspose3d)):
                raise TypeError("expected Conv3d/ConvTranspose3d at position %d of the Sequential" % i)
            conv = mods[i]
            i += 1
            act = None
            if i < len(mods) and isinstance(mods[i], (nn.ELU, nn.LeakyReLU, nn.Sigmoid)):
                act = mods[i]
                i += 1
            self.units.append(FusedUnit(bn, conv, act))

    def params(self):
        ps = []
        for u in self.units:
            ps += u.params()
        return ps

    def out_shape(self, shape):
        N, C, D, H, W = shape
        size = (D, H, W)
        for u in self.units:
            size = u.out_size(size)
        return (N, self.units[-1].cout) + size


class SeqSaved:
    __slots__ = ("acts", "bn", "G")

    def __init__(self):
        self.acts = []   # input + every unit's activation output
        self.bn = []     # per unit: None or (scale, shift, mean, invstd, training)
        self.G = 1


# After an optimizer step every weight image is stale: 2 packs per unit and step (forward and dgrad layout), ~10 us kernels that
# sat between the convolutions on the compute stream (44 launches = 0.45 ms of the CAE step, 1.2 %).  Only the first unit of a
# pass needs its image at once: the others — and every dgrad image — are packed on a side stream while that unit's BatchNorm
# statistics and convolution run.
PREPACK_AHEAD = True
_pack_streams = {}


def _prepack_ahead(plan, x, track):
    """Launch the stale packs of `plan` (all but unit 0's forward image) on the pack stream.  Returns (event after the forward
    images, event after the dgrad images); None where nothing was launched."""
    units = plan.units
    if not (PREPACK_AHEAD and x.is_cuda):
        return None, None
    N, size = x.shape[0], tuple(x.shape[2:])
    fwd, bwd = [], []
    for i, u in enumerate(units):
        d = u.desc(N, size)
        fw = 1 if u.transposed else 0
        if i > 0 and not u.has_pack(d, fw):
            fwd.append((u, d, fw))
        if track and not u.has_pack(d, 1 - fw):
            bwd.append((u, d, 1 - fw))
        size = u.out_size(size)
    if not fwd and not bwd:
        return None, None
    main = torch.cuda.current_stream(x.device)
    side = _pack_streams.get(x.device)
    if side is None:
        side = _pack_streams[x.device] = torch.cuda.Stream(x.device)
    side.wait_stream(main)       # the optimizer step that changed the weights is ahead of us on the compute stream
    ev_f = ev_b = None
    with torch.cuda.stream(side):
        for u, d, w in fwd:
            u.packed(d, w)
        if fwd:
            ev_f = torch.cuda.Event()
            ev_f.record(side)
        for u, d, w in bwd:
            u.packed(d, w)
        if bwd:
            ev_b = torch.cuda.Event()
            ev_b.record(side)
    return ev_f, ev_b


def seq_forward(plan, x, G=1, track=False):
    """x: NDHWC volume with N = G * B.  Returns (y, SeqSaved).  track: a backward pass will follow (counted for the
    plan_backward_hooks)."""
    if track:
        plan.pending_backward = getattr(plan, 'pending_backward', 0) + 1
    saved = SeqSaved()
    saved.G = G
    saved.acts.append(x)
    ev_fwd, ev_bwd = _prepack_ahead(plan, x, track)
    if ev_bwd is not None:
        plan.pack_event = ev_bwd
    for iu, u in enumerate(plan.units):
        if iu == 1 and ev_fwd is not None:
            torch.cuda.current_stream(x.device).wait_event(ev_fwd)
        N, C, D, H, W = x.shape
        if C != u.cin:
            raise RuntimeError("channel mismatch: unit expects %d, got %d" % (u.cin, C))
        scale = shift = None
        if u.bn is not None:
            bn = u.bn
            training = bn.training or bn.running_mean is None
            mom = 0.1 if bn.momentum is None else bn.momentum
            scale, shift, mean, invstd = ops.bn_forward(
                x, G, bn.weight if bn.affine else None, bn.bias if bn.affine else None,
                bn.running_mean, bn.running_var, bn.num_batches_tracked, mom, bn.eps, training)
            saved.bn.append((scale, shift, mean, invstd, training))
        else:
            saved.bn.append(None)
        d = u.desc(N, (D, H, W))
        osz = u.out_size((D, H, W))
        y = ops.new_vol(N, u.cout, osz[0], osz[1], osz[2], x.device)
        bias = u.conv.bias
        if u.transposed:
            ops.corrT(d, x, u.packed(d, 1), bias, scale, shift, G, y)
        else:
            ops.corr(d, x, u.packed(d, 0), bias, scale, shift, G, y)
        saved.acts.append(y)
        x = y
    if DEBUG_ACTS is not None:
        DEBUG_ACTS.append((plan, list(saved.acts)))
    return x, saved


def seq_backward(plan, saved, gy, need_input_grad, want):
    out = _seq_backward_impl(plan, saved, gy, need_input_grad, want)
    pending = getattr(plan, 'pending_backward', 0)
    if pending > 0:
        plan.pending_backward = pending - 1
        if pending == 1:
            for hook in plan_backward_hooks:
                hook(plan)
    return out


# First unit of a network (BatchNorm3d on the data -> Conv3d without padding, e.g. Unet3D.py:16-18 block1): the input needs no
# gradient, and the BatchNorm parameter gradients follow from the weight gradient taken against the normalised input
# (include/stroke_b200.h: sp_bn_grads_from_wgrad) — no transposed correlation, no reduction pass over its result.
FIRST_UNIT_SHORTCUT = True


def _first_unit_shortcut(u, bnrec, x, gz, gz_colsum, d, G, want, grads):
    """Weight / bias / BatchNorm gradients of unit 0 without its dgrad.  Returns False when the unit does not qualify."""
    conv, bn = u.conv, u.bn
    padded = max(u.pad)
    if not (FIRST_UNIT_SHORTCUT and bnrec is not None and bn is not None and bn.affine and not u.transposed
            and (padded == 0 or (u.k == 3 and u.s == 1 and padded <= 2 and min(u.pad) >= 0)) and gz_colsum is not None
            and conv.weight.is_contiguous()
            and want(conv.weight) and conv.bias is not None and want(conv.bias) and (want(bn.weight) or want(bn.bias))):
        return False
    mean, invstd = bnrec[2], bnrec[3]
    nshift = torch.mul(mean, invstd).neg_()
    dwh = torch.empty_like(conv.weight, memory_format=torch.contiguous_format)
    ops.wgrad(d, x, invstd, nshift, gz, None, None, G, dwh, 0.0)          # sum_v dZ[co, v] * xhat[ci, v + tap]
    dw, beta_dw = _sink_view(conv.weight), 1.0
    if dw is None:
        dw, beta_dw = torch.empty_like(conv.weight, memory_format=torch.contiguous_format), 0.0
        grads[conv.weight] = dw
    dgamma, dbeta, beta_acc = _sink_view(bn.weight), _sink_view(bn.bias), 1.0
    if dgamma is None or dbeta is None:
        dgamma, dbeta, beta_acc = torch.empty_like(bn.weight), torch.empty_like(bn.bias), 0.0
        grads[bn.weight], grads[bn.bias] = dgamma, dbeta
    # zero padding is applied after BatchNorm (Cae3D.py:40-41): taps of border voxels that read it see neither xhat nor beta
    excl = ops.border_tap_sums(gz, u.pad) if padded > 0 else None
    ops.bn_grads_from_wgrad(conv.weight.detach(), dwh, gz_colsum, excl, bn.weight.detach(), bn.bias.detach(), dw, beta_dw, dgamma,
                            dbeta, beta_acc)
    db, beta = _sink_view(conv.bias), 1.0
    if db is None:
        db, beta = torch.empty_like(conv.bias), 0.0
        grads[conv.bias] = db
    ops.bias_from_colsum(gz_colsum, db, beta)
    return True


def _seq_backward_impl(plan, saved, gy, need_input_grad, want):
    """gy: gradient w.r.t. the chain output (post-activation).  `want(param)` tells whether a parameter gradient is
    needed.  Returns (gx or None, {param: grad})."""
    G = saved.G
    grads = {}
    units = plan.units
    last = units[-1]
    # the dgrad images may have been launched on the pack stream by ANY forward pass of this plan (a later pass of the same step finds
    # them in the cache and launches nothing): wait for the plan's latest pack event — a no-op once it has completed
    ev = getattr(plan, 'pack_event', None)
    if ev is not None and gy.is_cuda:
        torch.cuda.current_stream(gy.device).wait_event(ev)
    # gradient w.r.t. the last conv output: gy * act'(y)
    def colsum_for(unit, like):
        """fp64 scratch for the bias gradient of `unit`, filled by the kernel that writes its output gradient."""
        if unit.conv.bias is not None and want(unit.conv.bias):
            return torch.empty(unit.cout, device=like.device, dtype=torch.float64)
        return None

    gz_colsum = None
    if last.act != ACT_NONE:
        gz_colsum = colsum_for(last, gy)
        gz = ops.bn_act_bwd_apply(gy, saved.acts[-1], None, G, last.act, last.alpha, colsum=gz_colsum)
    else:
        gz = gy if ops.is_ndhwc(gy) else ops.as_vol(gy)
    for i in range(len(units) - 1, -1, -1):
        u = units[i]
        x = saved.acts[i]
        if DEBUG_GZ is not None:
            DEBUG_GZ.append((plan, i, gz))
        N, C, D, H, W = x.shape
        bnrec = saved.bn[i]
        scale = shift = None
        if bnrec is not None:
            scale, shift = bnrec[0], bnrec[1]
        d = u.desc(N, (D, H, W), ACT_NONE, 0.0)
        conv = u.conv
        if i == 0 and not need_input_grad and _first_unit_shortcut(u, bnrec, x, gz, gz_colsum, d, G, want, grads):
            break
        # ---- parameter gradients
        if want(conv.weight):
            dw, beta = _sink_view(conv.weight), 1.0
            if dw is None:
                dw, beta = torch.empty_like(conv.weight, memory_format=torch.contiguous_format), 0.0
                grads[conv.weight] = dw
            if u.transposed:
                ops.wgrad(d, gz, None, None, x, scale, shift, G, dw, beta)
            else:
                ops.wgrad(d, x, scale, shift, gz, None, None, G, dw, beta)
        if conv.bias is not None and want(conv.bias):
            db, beta = _sink_view(conv.bias), 1.0
            if db is None:
                db, beta = torch.empty_like(conv.bias), 0.0
                grads[conv.bias] = db
            if gz_colsum is not None:
                ops.bias_from_colsum(gz_colsum, db, beta)
            else:
                gC = gz.shape[1]
                ops.bias_grad(gz, gz.numel() // gC, gC, gC, db, beta)
        bn_grads = u.bn is not None and u.bn.affine and (want(u.bn.weight) or want(u.bn.bias))
        need_dx = need_input_grad if i == 0 else True
        if not (need_dx or bn_grads):
            break
        # ---- gradient w.r.t. the BN output (= conv input)
        gxh = ops.new_vol(N, C, D, H, W, x.device)
        if u.transposed:
            ops.corr(d, gz, u.packed(d, 0), None, None, None, G, gxh)
        else:
            ops.corrT(d, gz, u.packed(d, 1), None, None, None, G, gxh)
        coef = None
        if bnrec is not None:
            bn = u.bn
            dgamma = dbeta = None
            beta_acc = 0.0
            if bn.affine and (want(bn.weight) or want(bn.bias)):
                dgamma, dbeta = _sink_view(bn.weight), _sink_view(bn.bias)
                if dgamma is not None and dbeta is not None:
                    beta_acc = 1.0
                else:
                    dgamma, dbeta = torch.empty_like(bn.weight), torch.empty_like(bn.bias)
                    grads[bn.weight], grads[bn.bias] = dgamma, dbeta
            coef = ops.bn_backward_coef(gxh, x, G, bn.weight if bn.affine else None, bnrec[2], bnrec[3], bnrec[4],
                                        dgamma, dbeta, beta_acc)
        if not need_dx:
            break
        prev_act, prev_alpha = (units[i - 1].act, units[i - 1].alpha) if i > 0 else (ACT_NONE, 0.0)
        if coef is None and prev_act == ACT_NONE:
            gz, gz_colsum = gxh, None
        else:
            gz_colsum = colsum_for(units[i - 1], gxh) if i > 0 else None
            gz = ops.bn_act_bwd_apply(gxh, x, coef, G, prev_act, prev_alpha, colsum=gz_colsum)
    else:
        return gz, grads
    return None, grads


class SeqFunction(torch.autograd.Function):
    """autograd node for one pass of a whole Sequential (encoder / decoder / step MLP)."""

    @staticmethod
    def forward(ctx, x, plan, G, *params):
        xv = ops.as_vol(x)
        y, saved = seq_forward(plan, xv, G, track=any(ctx.needs_input_grad))
        ctx.plan, ctx.saved, ctx.params = plan, saved, params
        return y

    @staticmethod
    def backward(ctx, gy):
        params = ctx.params
        needs = ctx.needs_input_grad
        wanted = {id(p) for j, p in enumerate(params) if needs[3 + j]}
        gy = ops.as_vol(gy)
        gx, grads = seq_backward(ctx.plan, ctx.saved, gy, needs[0], lambda p: id(p) in wanted)
        ctx.saved = None
        return (gx, None, None) + tuple(grads.get(p) for p in params)


def run_sequential(plan, x, G=1):
    return SeqFunction.apply(x, plan, G, *plan.params())


def _adjacent_views(ts):
    """True when the NDHWC volumes `ts` are consecutive batch slices of one dense buffer (zero-copy grouping)."""
    t0 = ts[0]
    if not ops.is_ndhwc(t0):
        return False
    base = t0.untyped_storage().data_ptr()
    off = t0.storage_offset()
    for t in ts:
        if (t.shape != t0.shape or t.dtype != t0.dtype or not ops.is_ndhwc(t) or t.untyped_storage().data_ptr() != base
                or t.storage_offset() != off or t.requires_grad):
            return False
        off += t.numel()
    return off * t0.element_size() <= t0.untyped_storage().nbytes()


def group_volumes(ts):
    """Stack same-shaped volumes along the batch axis as one NDHWC volume [G*B, C, D, H, W]."""
    if len(ts) == 1:
        return ts[0]
    if all(t.is_cuda for t in ts) and _adjacent_views(ts):
        B, C, D, H, W = ts[0].shape
        return ts[0].as_strided((len(ts) * B, C, D, H, W), ts[0].stride())
    # (a tensor that requires grad must reach the autograd node untouched: the layout kernel is not differentiable)
    return torch.cat([ops.as_vol(t) if (t.is_cuda and not t.requires_grad) else t for t in ts], dim=0)


class _SplitGroups(torch.autograd.Function):
    """[G*B, ...] -> G batch slices (views); the backward stitches the G incoming gradients into one volume."""

    @staticmethod
    def forward(ctx, y, G):
        ctx.G = G
        ctx.shape = y.shape
        B = y.shape[0] // G
        return tuple(y.narrow(0, g * B, B) for g in range(G))

    @staticmethod
    def backward(ctx, *gs):
        N, C, D, H, W = ctx.shape
        B = N // ctx.G
        dev = next(g for g in gs if g is not None).device
        out = ops.new_vol(N, C, D, H, W, dev)
        for i, g in enumerate(gs):
            dst = out.narrow(0, i * B, B)
            if g is None:
                dst.zero_()
            else:
                dst.copy_(g)
        return out, None


def run_sequential_grouped(plan, xs):
    """Run `plan` on G same-shaped inputs as ONE stacked pass with per-group BatchNorm statistics (identical to G
    separate calls in the reference's order, Cae3D.py:105-110,230-233).  Returns the G outputs."""
    G = len(xs)
    if G == 1:
        return [run_sequential(plan, xs[0], 1)]
    y = run_sequential(plan, group_volumes(xs), G)
    return list(_SplitGroups.apply(y, G))


# =====================================================================================================================
class UnetPlan:
    """The 3-scale U-Net graph of Unet3D.forward (Unet3D.py:56-79)."""

    def __init__(self, unet):
        self.blocks = [SeqPlan(getattr(unet, "block%d" % i).bn_conv_relu_2x) for i in range(1, 6)]
        self.classify = SeqPlan(unet.classify)
        self.unet = unet

    def params(self):
        ps = []
        for b in self.blocks:
            ps += b.params()
        return ps + self.classify.params()


_warned_legacy_upsample = False


def _align(unet, name):
    """align_corners of the trilinear x2 upsampling `unet.<name>` (Unet3D.py:44,46).

    ``unet.align_corners`` (None | bool), when set, overrides everything.  Otherwise the module's own attribute decides:
    None = the installed torch's default (False).  A module WITHOUT the attribute is a pickle written by the reference's
    pinned torch 0.3.1 (``nn.Upsample`` had no such field and interpolated with corner alignment): it keeps the
    behaviour it was trained with (True), with a one-time warning."""
    global _warned_legacy_upsample
    override = getattr(unet, "align_corners", None)
    if override is not None:
        return bool(override)
    mod = getattr(unet, name)
    if "align_corners" not in vars(mod):
        if not _warned_legacy_upsample:
            import warnings
            warnings.warn("Unet3D.%s has no align_corners attribute (checkpoint pickled by torch < 0.4): using "
                          "align_corners=True, the interpolation it was trained with; set unet.align_corners to override"
                          % name)
            _warned_legacy_upsample = True
        return True
    return bool(mod.align_corners)  # None (installed default) == False


def _crop_offsets(big, small):
    return tuple((b - s) // 2 for b, s in zip(big.shape[2:], small.shape[2:]))


def unet_forward(plan, x, track=False):
    b = plan.blocks
    S = {}
    y1, S["b1"] = seq_forward(b[0], x, track=track)
    p1 = ops.maxpool2_fwd(y1)
    y2, S["b2"] = seq_forward(b[1], p1, track=track)
    p2 = ops.maxpool2_fwd(y2)
    y3, S["b3"] = seq_forward(b[2], p2, track=track)
    N, C3, D3, H3, W3 = y3.shape
    C2 = y2.shape[1]
    cat4 = ops.new_vol(N, C3 + C2, 2 * D3, 2 * H3, 2 * W3, x.device)
    ops.upsample2_fwd(y3, cat4, 0, _align(plan.unet, "upsa34"))
    off4 = _crop_offsets(y2, cat4)
    ops.crop_into(y2, cat4, C3, off4)
    y4, S["b4"] = seq_forward(b[3], cat4, track=track)
    N, C4, D4, H4, W4 = y4.shape
    C1 = y1.shape[1]
    cat5 = ops.new_vol(N, C4 + C1, 2 * D4, 2 * H4, 2 * W4, x.device)
    ops.upsample2_fwd(y4, cat5, 0, _align(plan.unet, "upsa45"))
    off5 = _crop_offsets(y1, cat5)
    ops.crop_into(y1, cat5, C4, off5)
    y5, S["b5"] = seq_forward(b[4], cat5, track=track)
    seg, S["cls"] = seq_forward(plan.classify, y5, track=track)
    S.update(p1=p1, p2=p2, off4=off4, off5=off5, C1=C1, C2=C2, C3=C3, C4=C4)
    return seg, S


def unet_backward(plan, S, gseg, need_input_grad, want):
    b = plan.blocks
    grads = {}

    def run(p, key, g, need=True):
        gx, gr = seq_backward(p, S[key], g, need, want)
        grads.update(gr)
        return gx

    g5 = run(plan.classify, "cls", gseg)
    gcat5 = run(b[4], "b5", g5)
    g4 = ops.upsample2_bwd(gcat5, 0, S["C4"], _align(plan.unet, "upsa45"))
    gcat4 = run(b[3], "b4", g4)
    g3 = ops.upsample2_bwd(gcat4, 0, S["C3"], _align(plan.unet, "upsa34"))
    gp2 = run(b[2], "b3", g3)
    y2 = S["b2"].acts[-1]
    g2 = ops.maxpool2_bwd(y2, S["p2"], gp2)
    ops.crop_add(g2, gcat4, S["C3"], S["off4"])
    gp1 = run(b[1], "b2", g2)
    y1 = S["b1"].acts[-1]
    g1 = ops.maxpool2_bwd(y1, S["p1"], gp1)
    ops.crop_add(g1, gcat5, S["C4"], S["off5"])
    gx = run(b[0], "b1", g1, need_input_grad)
    return gx, grads


class UnetFunction(torch.autograd.Function):
    """autograd node for the whole U-Net: returns the two dense single-channel probability volumes."""

    @staticmethod
    def forward(ctx, x, plan, *params):
        xv = ops.as_vol(x)
        seg, S = unet_forward(plan, xv, track=any(ctx.needs_input_grad))
        ctx.plan, ctx.S, ctx.params = plan, S, params
        ctx.seg_shape = seg.shape
        outs = tuple(ops.extract_channel(seg, c) for c in range(seg.shape[1]))
        return outs

    @staticmethod
    def backward(ctx, *gouts):
        params = ctx.params
        needs = ctx.needs_input_grad
        wanted = {id(p) for j, p in enumerate(params) if needs[2 + j]}
        N, C, D, H, W = ctx.seg_shape
        dev = ctx.S["p1"].device
        if any(g is None for g in gouts):
            gseg = ops.zeros_vol(N, C, D, H, W, dev)
        else:
            gseg = ops.new_vol(N, C, D, H, W, dev)
        for c, g in enumerate(gouts):
            if g is not None:
                ops.insert_channel(ops.as_vol(g), gseg, c)
        gx, grads = unet_backward(ctx.plan, ctx.S, gseg, needs[0], lambda p: id(p) in wanted)
        ctx.S = None
        return (gx, None) + tuple(grads.get(p) for p in params)


def run_unet(plan, x):
    return UnetFunction.apply(x, plan, *plan.params())
