"""Thin Python wrappers over the C-ABI (include/stroke_b200.h): argument checking, pointers, current stream.

Volumes are torch tensors of *logical* shape N x C x D x H x W whose memory is dense NDHWC (what torch calls
``channels_last_3d``), so the reference's B x C x D x H x W API (README.md:13) is unchanged while the kernels see
channel-innermost voxels.  Nothing in this module computes on the CPU: a CUDA tensor is required everywhere.
"""
import ctypes

import torch

from . import _lib
from ._lib import ACT_ELU, ACT_LEAKY, ACT_NONE, ACT_SIGMOID, SpAdamTensor, SpConvDesc, check

__all__ = ["ACT_NONE", "ACT_ELU", "ACT_LEAKY", "ACT_SIGMOID"]


# ---------------------------------------------------------------------------------------------------- call proxy
# kernels launched per C-ABI entry point (for the `gpu_launches` figure bench.py reports)
_KERNELS_PER_CALL = {
    "sp_pack_weights": 1, "sp_corr": 1, "sp_corrT": 1, "sp_wgrad": 2, "sp_bias_grad": 2, "sp_bn_stats": 1,
    "sp_bn_finalize": 1, "sp_bn_bwd_reduce": 1, "sp_bn_bwd_finalize": 1, "sp_bn_act_bwd_apply": 1,
    "sp_maxpool2_fwd": 1, "sp_maxpool2_bwd": 1, "sp_upsample2_fwd": 1, "sp_upsample2_bwd": 1, "sp_crop_copy": 1,
    "sp_crop_add": 1, "sp_ncdhw_to_ndhwc": 1, "sp_ndhwc_to_ncdhw": 1, "sp_dice_sums": 1, "sp_dice_loss": 1,
    "sp_dice_bwd": 1, "sp_absdiff_mean": 2, "sp_absdiff_bwd": 1, "sp_binary_counts": 1, "sp_latent_interp_fwd": 1,
    "sp_latent_interp_bwd": 1, "sp_adam_multi": 1, "sp_surface_distances": 13, "sp_signed_distance": 11,
    "sp_gauss3d": 3, "sp_elastic_warp": 1, "sp_zoom_plane_xy": 1, "sp_flip_w": 1, "sp_pad_volume": 1,
    "sp_bias_from_colsum": 1, "sp_bn_grads_from_wgrad": 1, "sp_border_tap_sums": 2,
}


class _Stats:
    launches = 0          # kernels launched through the C-ABI since the last reset
    calls = 0
    profile = None        # None, or a list receiving (name, key, start_event, end_event)


def reset_launch_count():
    _Stats.launches = 0
    _Stats.calls = 0


def launch_count():
    return _Stats.launches


def start_profile():
    """Record a CUDA-event pair around every C-ABI call on the current stream (bench.py's per-kernel attribution)."""
    _Stats.profile = []


def stop_profile():
    """-> {(name, key): [ms, ...]} ; synchronises."""
    rec, _Stats.profile = _Stats.profile, None
    torch.cuda.synchronize()
    out = {}
    for name, key, e0, e1 in rec or []:
        out.setdefault((name, key), []).append(e0.elapsed_time(e1))
    return out


def _desc_key(args):
    a0 = args[0] if args else None
    d = getattr(a0, "_obj", None)
    if isinstance(d, SpConvDesc):
        return "N%d I%dx%dx%dx%d O%dx%dx%dx%d k%d s%d" % (d.N, d.Di, d.Hi, d.Wi, d.Ci, d.Do, d.Ho, d.Wo, d.Co, d.k, d.s)
    return ""


class _Proxy:
    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        nk = _KERNELS_PER_CALL.get(name, 0)
        if nk == 0:
            setattr(self, name, fn)
            return fn

        def call(*args):
            _Stats.launches += nk
            _Stats.calls += 1
            if _Stats.profile is None:
                return fn(*args)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            _Stats.profile.append((name, _desc_key(args), e0, e1))
            return rc

        setattr(self, name, call)
        return call


_proxy = None


def _L():
    global _proxy
    if _proxy is None:
        _proxy = _Proxy(_lib.load())
    return _proxy


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _req_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("stroke_prediction_b200 ops run on CUDA tensors only (no CPU fallback)")


# ---------------------------------------------------------------------------------------------------- volumes
def new_vol(N, C, D, H, W, device, dtype=torch.float32):
    """Uninitialised volume: logical (N,C,D,H,W), memory NDHWC dense."""
    return torch.empty((N, D, H, W, C), device=device, dtype=dtype).permute(0, 4, 1, 2, 3)


def zeros_vol(N, C, D, H, W, device, dtype=torch.float32):
    return torch.zeros((N, D, H, W, C), device=device, dtype=dtype).permute(0, 4, 1, 2, 3)


def is_ndhwc(x):
    return x.dim() == 5 and x.permute(0, 2, 3, 4, 1).is_contiguous()


def as_vol(x):
    """Return x as a dense-NDHWC fp32 CUDA volume (converting from dense NCDHW with our transpose kernel)."""
    _req_cuda(x)
    if x.dtype != torch.float32:
        x = x.float()
    if is_ndhwc(x):
        return x
    if not x.is_contiguous():
        x = x.contiguous()
    N, C, D, H, W = x.shape
    out = new_vol(N, C, D, H, W, x.device)
    check(_L().sp_ncdhw_to_ndhwc(_p(x), _p(out), N, C, D * H * W, _stream()), "sp_ncdhw_to_ndhwc")
    return out


def to_ncdhw(x):
    """Dense NCDHW copy of an NDHWC volume (for users who need torch-contiguous results)."""
    _req_cuda(x)
    N, C, D, H, W = x.shape
    if not is_ndhwc(x):
        return x.contiguous()
    out = torch.empty((N, C, D, H, W), device=x.device, dtype=x.dtype)
    check(_L().sp_ndhwc_to_ncdhw(_p(x), _p(out), N, C, D * H * W, _stream()), "sp_ndhwc_to_ncdhw")
    return out


# ---------------------------------------------------------------------------------------------------- conv family
def conv_desc(N, isize, Ci, osize, Co, k, s, pad, act=ACT_NONE, alpha=0.0, ldi=None, ldo=None):
    d = SpConvDesc()
    d.N = N
    d.Di, d.Hi, d.Wi = isize
    d.Ci = Ci
    d.ldi = Ci if ldi is None else ldi
    d.Do, d.Ho, d.Wo = osize
    d.Co = Co
    d.ldo = Co if ldo is None else ldo
    d.k = k
    d.s = s
    d.pd, d.ph, d.pw = pad
    d.act = act
    d.alpha = alpha
    return d


def with_act(d, act, alpha):
    e = SpConvDesc.from_buffer_copy(d)
    e.act = act
    e.alpha = alpha
    return e


def pack_weights(d, which, w):
    """w: torch-layout weight (Co,Ci,k,k,k) of the correlation geometry -> packed fp32 tensor."""
    _req_cuda(w)
    w = w.detach()
    if not w.is_contiguous():
        w = w.contiguous()
    n = _L().sp_packed_weight_floats(ctypes.byref(d), which)
    out = torch.empty(n, device=w.device, dtype=torch.float32)
    check(_L().sp_pack_weights(ctypes.byref(d), which, _p(w), _p(out), _stream()), "sp_pack_weights")
    return out


def set_tc_terms(terms):
    """Mode of the tensor-core tier of the 16-channel 3x3x3 stride-1 correlations (include/stroke_b200.h): 4 = pipelined
    split-accumulator tcgen05 kernel (default, fp32-grade), 0 = exact fp32 FFMA tier, 2 / 3 = first-generation kernels.
    Packed weights depend on the mode, so cached packs are dropped."""
    check(_L().sp_set_tc_terms(int(terms)), "sp_set_tc_terms")
    from . import engine
    engine.bump_weights_epoch()


def get_tc_terms():
    return int(_L().sp_get_tc_terms())


def set_wgrad_tc_options(generation=2, max_ctas=0):
    """Weight-gradient tier of the 16-channel 3x3x3 stride-1 layers (include/stroke_b200.h): generation 2 = M 128 x N 96
    MMAs on two round-to-nearest bf16 terms (default), 1 = the first-generation kernel; max_ctas caps the persistent grid."""
    check(_L().sp_set_wgrad_tc_options(int(generation), int(max_ctas)), "sp_set_wgrad_tc_options")


def _conv_ws(d, which, device):
    nbytes = _L().sp_conv_workspace_bytes(ctypes.byref(d), which)
    if nbytes == 0:
        return None, 0
    ws = workspace(nbytes, device, "conv")
    return ws, ws.numel()


def corr(d, src, wp, bias, scale, shift, G, dst):
    _req_cuda(src, wp, dst)
    ws, n = _conv_ws(d, 0, dst.device)
    check(_L().sp_corr(ctypes.byref(d), _p(src), _p(wp), _p(bias), _p(scale), _p(shift), G, _p(dst), _p(ws), n, _stream()), "sp_corr")
    return dst


def corrT(d, src, wp, bias, scale, shift, G, dst):
    _req_cuda(src, wp, dst)
    ws, n = _conv_ws(d, 1, dst.device)
    check(_L().sp_corrT(ctypes.byref(d), _p(src), _p(wp), _p(bias), _p(scale), _p(shift), G, _p(dst), _p(ws), n, _stream()), "sp_corrT")
    return dst


_ws_cache = {}


def workspace(nbytes, device, tag="ws"):
    """Grow-only scratch buffer per (device, tag); stream-ordered reuse on the current stream."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), tag)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), device=device, dtype=torch.uint8)
        _ws_cache[key] = buf
    return buf


def wgrad(d, iside, i_scale, i_shift, oside, o_scale, o_shift, G, dw, beta=0.0):
    _req_cuda(iside, oside, dw)
    assert dw.is_contiguous()
    nbytes = _L().sp_wgrad_workspace_bytes(ctypes.byref(d))
    ws = workspace(nbytes, dw.device, "wgrad")
    check(_L().sp_wgrad(ctypes.byref(d), _p(iside), _p(i_scale), _p(i_shift), _p(oside), _p(o_scale), _p(o_shift), G,
                        _p(dw), beta, _p(ws), ws.numel(), _stream()), "sp_wgrad")
    return dw


def bias_grad(g, rows, C, ld, db, beta=0.0):
    ws = workspace(8 * C, db.device, "f64")
    check(_L().sp_bias_grad(_p(g), rows, C, ld, _p(db), beta, _p(ws), _stream()), "sp_bias_grad")
    return db


# ---------------------------------------------------------------------------------------------------- batch norm
def bn_forward(x, G, gamma, beta, running_mean, running_var, nbt, momentum, eps, training):
    """x: volume (N = G*B).  Returns (scale, shift, mean, invstd), each [G, C]; updates running stats in training."""
    N, C, D, H, W = x.shape
    vox = D * H * W
    dev = x.device
    sums = None
    if training:
        sums = torch.empty((G, C, 2), device=dev, dtype=torch.float64)
        check(_L().sp_bn_stats(_p(x), N, vox, C, C, G, _p(sums), _stream()), "sp_bn_stats")
    out = torch.empty((4, G, C), device=dev, dtype=torch.float32)
    check(_L().sp_bn_finalize(_p(sums), (N // G) * vox, C, G, _p(gamma), _p(beta), _p(running_mean), _p(running_var),
                              _p(nbt), float(momentum), float(eps), int(bool(training)),
                              _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]), _stream()), "sp_bn_finalize")
    return out[0], out[1], out[2], out[3]


def bn_backward_coef(gxh, x, G, gamma, mean, invstd, training, dgamma, dbeta, beta_acc=0.0):
    """Reduce + finalise BN backward.  Returns coef [4, G, C]; writes dgamma / dbeta when given."""
    N, C, D, H, W = x.shape
    vox = D * H * W
    bsums = torch.empty((G, C, 2), device=x.device, dtype=torch.float64)
    check(_L().sp_bn_bwd_reduce(_p(gxh), C, _p(x), C, N, vox, C, G, _p(bsums), _stream()), "sp_bn_bwd_reduce")
    coef = torch.empty((4, G, C), device=x.device, dtype=torch.float32)
    check(_L().sp_bn_bwd_finalize(_p(bsums), (N // G) * vox, C, G, _p(gamma), _p(mean), _p(invstd), int(bool(training)),
                                  _p(dgamma), _p(dbeta), beta_acc, _p(coef), _stream()), "sp_bn_bwd_finalize")
    return coef


def bn_act_bwd_apply(gxh, x, coef, G, act, alpha, out=None, accumulate=False, colsum=None):
    """out (+)= A*((gxh - m1) - (x - mu)*k) * act'(x); coef None -> gxh * act'(x).  colsum: optional fp64 [C] tensor
    receiving the column sums of `out` (fused bias gradient, see bias_from_colsum)."""
    N, C, D, H, W = gxh.shape
    if out is None:
        out = new_vol(N, C, D, H, W, gxh.device)
        accumulate = False
    check(_L().sp_bn_act_bwd_apply(_p(gxh), C, _p(x), C, _p(coef), N, D * H * W, C, G, act, float(alpha), _p(out), C,
                                   int(accumulate), _p(colsum), _stream()), "sp_bn_act_bwd_apply")
    return out


def bias_from_colsum(colsum, db, beta=0.0):
    check(_L().sp_bias_from_colsum(_p(colsum), db.numel(), _p(db), beta, _stream()), "sp_bias_from_colsum")
    return db


TAP_EXCL_REPLICAS = 16      # SP_TAP_EXCL_REPLICAS of include/stroke_b200.h


def border_tap_sums(gz, pad):
    """[Co, 27] fp64: per tap of a 3x3x3 stride-1 convolution with padding `pad`, the sum of gz over the output voxels whose tap
    reads zero padding (include/stroke_b200.h)."""
    _req_cuda(gz)
    assert is_ndhwc(gz)
    N, C, D, H, W = gz.shape
    excl = torch.empty((TAP_EXCL_REPLICAS, C, 27), device=gz.device, dtype=torch.float64)
    check(_L().sp_border_tap_sums(_p(gz), C, N, D, H, W, C, int(pad[0]), int(pad[1]), int(pad[2]), _p(excl), _stream()),
          "sp_border_tap_sums")
    return excl[0]


def bn_grads_from_wgrad(w, dw_hat, colsum, tap_excl, gamma, beta, dw, beta_dw, dgamma, dbeta, beta_acc):
    """First-unit shortcut (include/stroke_b200.h): BN parameter gradients and dW from the weight gradient against the normalised
    input; w / dw_hat / dw in torch layout (Co, Ci, k, k, k)."""
    _req_cuda(w, dw_hat, colsum, dw, tap_excl)
    assert w.is_contiguous() and dw_hat.is_contiguous() and dw.is_contiguous() and colsum.dtype == torch.float64
    Co, Ci = w.shape[0], w.shape[1]
    k3 = w.numel() // (Co * Ci)
    assert tap_excl is None or (tap_excl.dtype == torch.float64 and tuple(tap_excl.shape) == (Co, k3))
    check(_L().sp_bn_grads_from_wgrad(_p(w), _p(dw_hat), _p(colsum), _p(tap_excl), _p(gamma), _p(beta), Co, Ci, k3, _p(dw), beta_dw,
                                      _p(dgamma), _p(dbeta), beta_acc, _stream()), "sp_bn_grads_from_wgrad")


# ---------------------------------------------------------------------------------------------------- resampling
def maxpool2_fwd(x):
    N, C, D, H, W = x.shape
    y = new_vol(N, C, D // 2, H // 2, W // 2, x.device)
    check(_L().sp_maxpool2_fwd(_p(x), N, D, H, W, C, C, _p(y), C, _stream()), "sp_maxpool2_fwd")
    return y


def maxpool2_bwd(x, y, gy):
    N, C, D, H, W = x.shape
    gx = new_vol(N, C, D, H, W, x.device)
    check(_L().sp_maxpool2_bwd(_p(x), _p(y), _p(gy), N, D, H, W, C, _p(gx), _stream()), "sp_maxpool2_bwd")
    return gx


def _chan_ptr(t, c0):
    return ctypes.c_void_p(t.data_ptr() + 4 * c0)


def upsample2_fwd(x, out, c0, align_corners):
    """Trilinear x2 of x into channels [c0, c0+C) of the (wider) volume `out`."""
    N, C, D, H, W = x.shape
    ldy = out.shape[1]
    check(_L().sp_upsample2_fwd(_p(x), N, D, H, W, C, C, _chan_ptr(out, c0), ldy, int(align_corners), _stream()),
          "sp_upsample2_fwd")
    return out


def upsample2_bwd(gcat, c0, C, align_corners):
    """Gradient w.r.t. the low-resolution input from channels [c0, c0+C) of the concat-buffer gradient."""
    N, Ct, Do, Ho, Wo = gcat.shape
    D, H, W = Do // 2, Ho // 2, Wo // 2
    gx = new_vol(N, C, D, H, W, gcat.device)
    check(_L().sp_upsample2_bwd(_chan_ptr(gcat, c0), Ct, N, D, H, W, C, _p(gx), C, int(align_corners), _stream()),
          "sp_upsample2_bwd")
    return gx


def crop_into(src, out, c0, offs):
    """out[:, c0:c0+C] = centre crop of src (offsets offs) — the skip connection half of the virtual concat."""
    N, C, Ds, Hs, Ws = src.shape
    _, Ct, Dd, Hd, Wd = out.shape
    check(_L().sp_crop_copy(_p(src), Ds, Hs, Ws, C, _chan_ptr(out, c0), Dd, Hd, Wd, Ct, N, C, offs[0], offs[1], offs[2],
                            _stream()), "sp_crop_copy")
    return out


def crop_add(big, gcat, c0, offs):
    """big[crop window] += gcat[:, c0:c0+C]  (gradient of crop_into, accumulated into the skip tensor's gradient)."""
    N, C, Db, Hb, Wb = big.shape
    _, Ct, Ds, Hs, Ws = gcat.shape
    check(_L().sp_crop_add(_p(big), Db, Hb, Wb, C, _chan_ptr(gcat, c0), Ds, Hs, Ws, Ct, N, C, offs[0], offs[1], offs[2],
                           _stream()), "sp_crop_add")
    return big


def extract_channel(x, c):
    """Dense (N,1,D,H,W) copy of channel c of an NDHWC volume."""
    N, C, D, H, W = x.shape
    out = new_vol(N, 1, D, H, W, x.device)
    check(_L().sp_crop_copy(_chan_ptr(x, c), D, H, W, C, _p(out), D, H, W, 1, N, 1, 0, 0, 0, _stream()), "sp_crop_copy")
    return out


def insert_channel(src, out, c):
    """out[:, c] = src (src dense single-channel volume)."""
    N, C, D, H, W = out.shape
    check(_L().sp_crop_copy(_p(src), D, H, W, 1, _chan_ptr(out, c), D, H, W, C, N, 1, 0, 0, 0, _stream()), "sp_crop_copy")
    return out


# ---------------------------------------------------------------------------------------------------- losses
def _dense(t):
    """1-D view semantics: any tensor whose elements are densely packed (any dim order) is fine for reductions."""
    if t.is_contiguous() or (t.dim() == 5 and is_ndhwc(t)):
        return t
    return t.contiguous()


def dice_sums(o, t):
    _req_cuda(o, t)
    sums = torch.empty(3, device=o.device, dtype=torch.float64)
    check(_L().sp_dice_sums(_p(o), _p(t), o.numel(), _p(sums), _stream()), "sp_dice_sums")
    return sums


def dice_loss(sums, w, eps):
    loss = torch.empty((), device=sums.device, dtype=torch.float32)
    check(_L().sp_dice_loss(_p(sums), float(w), float(eps), _p(loss), _stream()), "sp_dice_loss")
    return loss


def dice_bwd(o, t, sums, w, eps, gscale, gmul, go, accumulate):
    check(_L().sp_dice_bwd(_p(o), _p(t), o.numel(), _p(sums), float(w), float(eps), _p(gscale), float(gmul), _p(go),
                           int(accumulate), _stream()), "sp_dice_bwd")
    return go


def absdiff_mean(a, b, mode):
    ws = torch.empty(1, device=a.device, dtype=torch.float64)
    out = torch.empty((), device=a.device, dtype=torch.float32)
    check(_L().sp_absdiff_mean(_p(a), _p(b), a.numel(), mode, _p(ws), _p(out), _stream()), "sp_absdiff_mean")
    return out


def absdiff_bwd(a, b, mode, gscale, gmul, ga, acc_a, gb, acc_b):
    check(_L().sp_absdiff_bwd(_p(a), _p(b), a.numel(), mode, _p(gscale), float(gmul), _p(ga), int(acc_a), _p(gb),
                              int(acc_b), _stream()), "sp_absdiff_bwd")


def binary_counts(result, target, threshold, out=None):
    """TP, FP, FN, TN of (result > threshold) vs (target > threshold) as 4 doubles on the device (metrics.py:31-47)."""
    _req_cuda(result, target)
    assert result.numel() == target.numel()
    result, target = result.contiguous(), target.contiguous()
    if out is None:
        out = torch.empty(4, device=result.device, dtype=torch.float64)
    check(_L().sp_binary_counts(_p(result), _p(target), result.numel(), float(threshold), _p(out), _stream()), "sp_binary_counts")
    return out


def _lattice(t):
    """Dense tensor of any rank -> (n0, n1, n2, n3, had_singleton_axis): extent-1 axes dropped, at most 4 axes left."""
    dims = [int(d) for d in t.shape]
    keep = [d for d in dims if d != 1]
    if len(keep) > 4:
        raise RuntimeError("surface distances: at most 4 axes of extent > 1 are supported, got shape %r" % (tuple(dims),))
    while len(keep) < 4:
        keep.insert(0, 1)
    return keep[0], keep[1], keep[2], keep[3], len([d for d in dims if d == 1]) > 0


def surface_distances(result, target, threshold=0.5, out=None):
    """hd / assd (+ the directed parts) of (result > threshold) vs (target > threshold) as 8 doubles on the device; the
    arrays are taken as ONE lattice exactly like medpy takes the reference's whole B x 1 x D x H x W batch (metrics.py:43-45)."""
    _req_cuda(result, target)
    if tuple(result.shape) != tuple(target.shape):
        raise RuntimeError("surface_distances: shape mismatch %r vs %r" % (tuple(result.shape), tuple(target.shape)))
    result = _dense_c(result)
    target = _dense_c(target)
    n0, n1, n2, n3, single = _lattice(result)
    total = n0 * n1 * n2 * n3
    if out is None:
        out = torch.empty(8, device=result.device, dtype=torch.float64)
    ws = workspace(_L().sp_surface_distances_workspace_bytes(total), result.device, "metrics")
    check(_L().sp_surface_distances(_p(result), _p(target), n0, n1, n2, n3, int(single), float(threshold), _p(out), _p(ws),
                                    ws.numel(), _stream()), "sp_surface_distances")
    return out


def signed_distance(mask, threshold=0.5, outside_is_lt=True, sign=1.0):
    """sign * (edt(mask > thr) - edt(outside)) over the dense lattice of `mask` (extent-1 axes ignored), fp32."""
    _req_cuda(mask)
    mask = _dense_c(mask)
    n0, n1, n2, n3, _ = _lattice(mask)
    total = n0 * n1 * n2 * n3
    out = torch.empty_like(mask)
    ws = workspace(_L().sp_signed_distance_workspace_bytes(total), mask.device, "metrics")
    check(_L().sp_signed_distance(_p(mask), n0, n1, n2, n3, float(threshold), int(bool(outside_is_lt)), float(sign), _p(out),
                                  _p(ws), ws.numel(), _stream()), "sp_signed_distance")
    return out


def _dense_c(t):
    """fp32, C-contiguous in its LOGICAL axis order (a single-channel NDHWC volume already is)."""
    if t.dtype != torch.float32:
        t = t.float()
    if t.is_contiguous():
        return t
    if t.dim() == 5 and t.shape[1] == 1 and is_ndhwc(t):
        return t.permute(0, 2, 3, 4, 1).reshape(t.shape)      # same memory, contiguous strides (C = 1)
    return t.contiguous()


# ---------------------------------------------------------------------------------------------------- augmentation
def _ncdhw(x):
    _req_cuda(x)
    if x.dim() != 5:
        raise RuntimeError("expected a B x C x D x H x W tensor, got shape %r" % (tuple(x.shape),))
    if x.dtype != torch.float32:
        x = x.float()
    return x if x.is_contiguous() else x.contiguous()


def gauss3d(fields, sigma=4.0, truncate=4.0):
    """scipy gaussian_filter(mode='constant', cval=0) of fp64 volumes [..., D, H, W] (any leading axes)."""
    _req_cuda(fields)
    f = fields.double().contiguous()
    D, H, W = f.shape[-3:]
    nvol = f.numel() // (D * H * W)
    out, tmp = torch.empty_like(f), torch.empty_like(f)
    check(_L().sp_gauss3d(_p(f), nvol, D, H, W, float(sigma), float(truncate), _p(out), _p(tmp), _stream()), "sp_gauss3d")
    return out


def elastic_warp(x, f1, f2, f3, alpha=100.0, zscale=0.22):
    """x: B x C x D x H x W fp32; f1..f3: smoothed fp64 noise fields of the same shape (reference draw order)."""
    x = _ncdhw(x)
    B, C, D, H, W = x.shape
    fs = [f.double().contiguous() for f in (f1, f2, f3)]
    for f in fs:
        if tuple(f.shape) != tuple(x.shape):
            raise RuntimeError("elastic_warp: displacement fields must have the volume's shape")
    out = torch.empty_like(x)
    check(_L().sp_elastic_warp(_p(x), _p(fs[0]), _p(fs[1]), _p(fs[2]), B * C, D, H, W, float(alpha), float(zscale), _p(out),
                               _stream()), "sp_elastic_warp")
    return out


def zoom_plane_xy(x, scale_factor, order=0):
    x = _ncdhw(x)
    B, C, D, H, W = x.shape
    Ho, Wo = int(round(H * scale_factor)), int(round(W * scale_factor))       # scipy.ndimage.zoom: round(n * zoom)
    out = torch.empty((B, C, D, Ho, Wo), device=x.device, dtype=torch.float32)
    check(_L().sp_zoom_plane_xy(_p(x), B * C * D, H, W, Ho, Wo, int(order), _p(out), _stream()), "sp_zoom_plane_xy")
    return out


def flip_w(x):
    x = _ncdhw(x)
    out = torch.empty_like(x)
    W = x.shape[-1]
    check(_L().sp_flip_w(_p(x), x.numel() // W, W, _p(out), _stream()), "sp_flip_w")
    return out


def pad_volume(x, pd, ph, pw, value=0.0):
    x = _ncdhw(x)
    B, C, D, H, W = x.shape
    out = torch.empty((B, C, D + 2 * pd, H + 2 * ph, W + 2 * pw), device=x.device, dtype=torch.float32)
    check(_L().sp_pad_volume(_p(x), B * C, D, H, W, pd, ph, pw, float(value), _p(out), _stream()), "sp_pad_volume")
    return out


def crop_volume(x, offs, size):
    """x[:, :, od:od+D, oh:oh+H, ow:ow+W] as a dense copy (RandomPatch, data.py:248-277)."""
    x = _ncdhw(x)
    B, C, Ds, Hs, Ws = x.shape
    D, H, W = size
    out = torch.empty((B, C, D, H, W), device=x.device, dtype=torch.float32)
    check(_L().sp_crop_copy(_p(x), Ds, Hs, Ws, 1, _p(out), D, H, W, 1, B * C, 1, offs[0], offs[1], offs[2], _stream()), "sp_crop_copy")
    return out


# ---------------------------------------------------------------------------------------------------- interpolation
def latent_interp_fwd(zc, zp, step):
    B = zc.shape[0]
    per = zc.numel() // B
    out = torch.empty_like(zc)
    check(_L().sp_latent_interp_fwd(_p(zc), _p(zp), _p(step), B, per, _p(out), _stream()), "sp_latent_interp_fwd")
    return out


def latent_interp_bwd(g, zc, zp, step, need_c, need_p, need_s):
    B = zc.shape[0]
    per = zc.numel() // B
    dzc = torch.empty_like(zc) if need_c else None
    dzp = torch.empty_like(zp) if need_p else None
    ds = torch.empty(B, device=zc.device, dtype=torch.float32) if need_s else None
    ws = torch.empty(B, device=zc.device, dtype=torch.float64) if need_s else None
    check(_L().sp_latent_interp_bwd(_p(g), _p(zc), _p(zp), _p(step), B, per, _p(dzc), 0, _p(dzp), 0, _p(ds), _p(ws),
                                    _stream()), "sp_latent_interp_bwd")
    return dzc, dzp, ds


# ---------------------------------------------------------------------------------------------------- optimizer
def adam_multi(table_dev, n_tensors, total_blocks, lr, beta1, beta2, eps, wd, step, grad_scale, zero_grad):
    check(_L().sp_adam_multi(_p(table_dev), n_tensors, total_blocks, float(lr), float(beta1), float(beta2), float(eps),
                             float(wd), int(step), float(grad_scale), int(zero_grad), _stream()), "sp_adam_multi")


def make_adam_table(entries, device):
    """entries: list of (p, g, m, v) fp32 contiguous CUDA tensors -> (device uint8 tensor holding the table, blocks)."""
    arr = (SpAdamTensor * len(entries))()
    start = 0
    for i, (p, g, m, v) in enumerate(entries):
        n = p.numel()
        arr[i].p, arr[i].g, arr[i].m, arr[i].v = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
        arr[i].n = n
        arr[i].block_start = start
        start += (n + _lib.SP_ADAM_CHUNK - 1) // _lib.SP_ADAM_CHUNK
    raw = bytes(arr)
    host = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
    return host.to(device), start
