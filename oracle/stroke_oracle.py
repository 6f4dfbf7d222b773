"""CPU oracle for the volumetric hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module; nothing
under ``stroke-prediction_b200/`` does.  It restates, with plain ``torch.nn.functional`` calls on CPU tensors and an
explicit ``state_dict`` (reference parameter names), the algorithm of

  * ``Enc3D`` / ``Enc3DStep`` / ``Enc3DCtp`` / ``Dec3D`` / ``Cae3D``       /root/reference/common/model/Cae3D.py:35-260
  * ``Block3x3x3`` / ``Unet3D``                                            /root/reference/common/model/Unet3D.py:14-84
  * ``BatchDiceLoss``                                                      /root/reference/common/metrics.py:8-28
  * the four ``loss_step``s    /root/reference/learner/{CaeReconstruction,CaeStep,CaePrediction,UnetSegmentation}Learner.py
  * time normalisation                                                     /root/reference/common/inference/CaeInference.py:18-31
  * the Adam update as stepped from Learner.py:120-122 (installed-torch form)

The arithmetic itself lives in the third-party ``torch==0.3.1`` wheel (requirements.txt:7), which is not installable;
as BASELINE.json prescribes, the oracle runs on the closest installable CPU torch (2.11).  Parity pin: the reference
ships no golden vectors, so ``oracle/make_golden.py`` imports the *reference's own modules* from
/root/reference, checks this restatement against them bit-for-bit on seeded inputs and writes their outputs to
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` re-checks the restatement against those files everywhere.
"""
import math

import torch
import torch.nn.functional as F

EPS_BN = 1e-5
MOMENTUM_BN = 0.1


# ------------------------------------------------------------------------------------------------ layer tables
def enc_table(channels):
    """(cin, cout, stride, padding) of the ten encoder convolutions (Cae3D.py:40-75)."""
    c0, c1, c2, c4, c8, cf = channels[0], channels[1], channels[2], channels[3], channels[4], channels[5]
    z = (1, 0, 0)
    return [(c0, c1, 1, z), (c1, c1, 1, z), (c1, c2, 2, (1, 1, 1)), (c2, c2, 1, z), (c2, c2, 1, z),
            (c2, c4, 2, (1, 1, 1)), (c4, c4, 1, z), (c4, c4, 1, z), (c4, c8, 2, (0, 0, 0)), (c8, cf, 1, (0, 0, 0))]


def dec_table(channels):
    """(kind, cin, cout, k, stride, padding) of the twelve decoder layers (Cae3D.py:177-218)."""
    c1, c2, c4, c8, cf, ncls = channels[1], channels[2], channels[3], channels[4], channels[5], channels[-1]
    g = (1, 2, 2)
    o = (0, 0, 0)
    return [('T', cf, c8, 3, 1, o), ('T', c8, c4, 3, 2, o), ('C', c4, c4, 3, 1, g), ('C', c4, c2, 3, 1, g),
            ('T', c2, c2, 2, 2, o), ('C', c2, c2, 3, 1, g), ('C', c2, c1, 3, 1, g), ('T', c1, c1, 2, 2, o),
            ('C', c1, c1, 3, 1, g), ('C', c1, c1, 3, 1, g), ('C', c1, c1, 1, 1, o), ('C', c1, ncls, 1, 1, o)]


def _bn(x, sd, prefix, training):
    """BatchNorm3d with in-place running-stat update on the tensors of `sd`."""
    rm, rv = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
    y = F.batch_norm(x, rm, rv, sd[prefix + '.weight'], sd[prefix + '.bias'], training, MOMENTUM_BN, EPS_BN)
    if training and (prefix + '.num_batches_tracked') in sd:
        sd[prefix + '.num_batches_tracked'] += 1
    return y


def encoder_pass(x, sd, channels, alpha, training, prefix='enc.encoder'):
    """One call of Enc3D.encoder (Cae3D.py:91-94): ten BN -> Conv3d -> ELU units."""
    if x is None:
        return None
    for i, (cin, cout, stride, pad) in enumerate(enc_table(channels)):
        x = _bn(x, sd, '%s.%d' % (prefix, 3 * i), training)
        x = F.conv3d(x, sd['%s.%d.weight' % (prefix, 3 * i + 1)], sd['%s.%d.bias' % (prefix, 3 * i + 1)], stride, pad)
        x = F.elu(x, alpha)
    return x


def decoder_pass(z, sd, channels, alpha, training, prefix='dec.decoder'):
    """One call of Dec3D.decoder (Cae3D.py:222-225): eleven BN -> Conv/ConvT -> ELU units + BN -> 1x1 -> Sigmoid."""
    if z is None:
        return None
    table = dec_table(channels)
    for i, (kind, cin, cout, k, stride, pad) in enumerate(table):
        z = _bn(z, sd, '%s.%d' % (prefix, 3 * i), training)
        w, b = sd['%s.%d.weight' % (prefix, 3 * i + 1)], sd['%s.%d.bias' % (prefix, 3 * i + 1)]
        z = F.conv_transpose3d(z, w, b, stride, pad) if kind == 'T' else F.conv3d(z, w, b, stride, pad)
        z = torch.sigmoid(z) if i == len(table) - 1 else F.elu(z, alpha)
    return z


def interpolate(z_core, z_penu, step):
    """Cae3D.py:78-89: z_c + s (z_p - z_c), s broadcast from B x 1 x 1 x 1 x 1."""
    if z_core is None or z_penu is None:
        return None
    return z_core + step * (z_penu - z_core)


def step_from_globals(globals_, sd, alpha, prefix='enc'):
    """Enc3DStep._get_step with time_to_treatment None (Cae3D.py:138-142)."""
    h = F.elu(F.conv3d(globals_, sd[prefix + '.reduce.0.weight'], sd[prefix + '.reduce.0.bias']), alpha)
    h = F.elu(F.conv3d(h, sd[prefix + '.reduce.2.weight'], sd[prefix + '.reduce.2.bias']), alpha)
    return torch.sigmoid(F.conv3d(h, sd[prefix + '.step.weight'], sd[prefix + '.step.bias']))


def time_to_treatment(clinical, normalization_hours=10.0, step=None):
    """CaeInference.py:18-31: t_adm->rec / (norm - t_onset->adm) in fp32, B x 1 x 1 x 1 x 1."""
    t0 = clinical[:, 0].float().reshape(-1, 1)
    norm = torch.ones(t0.shape[0], 1) * normalization_hours - t0
    if step is None:
        t = clinical[:, 1].float().reshape(-1, 1) / norm
    else:
        t = (step * torch.ones(t0.shape[0], 1)) / norm
    return t.reshape(-1, 1, 1, 1, 1)


def cae_forward(sd, channels, alpha, training, core, penu, lesion, step, enc_prefix='enc', dec_prefix='dec'):
    """Cae3D.forward on the gtruth branch (Cae3D.py:103-110,228-233): 3 encoder + 4 decoder passes in the
    reference's order (running statistics are updated sequentially).  Returns dicts of latents, reconstructions."""
    e = lambda x: encoder_pass(x, sd, channels, alpha, training, enc_prefix + '.encoder')
    d = lambda z: decoder_pass(z, sd, channels, alpha, training, dec_prefix + '.decoder')
    lat = {}
    lat['core'], lat['penu'], lat['lesion'] = e(core), e(penu), e(lesion)
    lat['interpolation'] = interpolate(lat['core'], lat['penu'], step)
    rec = {k: d(lat[k]) for k in ('core', 'penu', 'lesion', 'interpolation')}
    return lat, rec


# ------------------------------------------------------------------------------------------------ U-Net
def block3x3x3(x, sd, prefix, training):
    """Unet3D.py:14-27."""
    for j in (0, 3):
        x = _bn(x, sd, '%s.bn_conv_relu_2x.%d' % (prefix, j), training)
        x = F.conv3d(x, sd['%s.bn_conv_relu_2x.%d.weight' % (prefix, j + 1)], sd['%s.bn_conv_relu_2x.%d.bias' % (prefix, j + 1)])
        x = F.leaky_relu(x, 0.01)
    return x


def center_crop(t, like):
    """Unet3D.py:6-11 over dims 2,3,4."""
    for dim in (2, 3, 4):
        t = t.narrow(dim, (t.size(dim) - like.size(dim)) // 2, like.size(dim))
    return t


def unet_forward(sd, x, training, align_corners=False):
    """Unet3D.forward (Unet3D.py:56-79).  Returns (core, penu) probability volumes."""
    up = lambda t: F.interpolate(t, scale_factor=2, mode='trilinear', align_corners=align_corners)
    b1 = block3x3x3(x, sd, 'block1', training)
    b2 = block3x3x3(F.max_pool3d(b1, 2, 2), sd, 'block2', training)
    b3 = block3x3x3(F.max_pool3d(b2, 2, 2), sd, 'block3', training)
    u3 = up(b3)
    b4 = block3x3x3(torch.cat((u3, center_crop(b2, u3)), 1), sd, 'block4', training)
    u4 = up(b4)
    b5 = block3x3x3(torch.cat((u4, center_crop(b1, u4)), 1), sd, 'block5', training)
    h = F.leaky_relu(F.conv3d(b5, sd['classify.0.weight'], sd['classify.0.bias']), 0.01)
    seg = torch.sigmoid(F.conv3d(h, sd['classify.2.weight'], sd['classify.2.bias']))
    return seg[:, 0:1], seg[:, 1:2]


# ------------------------------------------------------------------------------------------------ losses
def dice_loss(o, t, w=1.0, eps=1e-7):
    """BatchDiceLoss with one label (metrics.py:16-28): sums over the whole batch."""
    o, t = o.reshape(-1), t.reshape(-1)
    return 1.0 - w * (2.0 * (o * t).sum() + eps) / ((o * o).sum() + (t * t).sum() + eps)


def hinge(a, b, sign=None):
    """mean(|d| - d), d = a - b (CaeReconstructionLearner.py:59-62).  `sign` (tests only): evaluate the kink with a GIVEN
    sign pattern, i.e. mean(sign * d - d).  Where sign == sign(d) this is the same value and the same gradient
    (sign(d) - 1) / N; the parity tests use it to take the decision at |d| ~ 0 from the implementation under test, so
    that a gradient comparison is not dominated by voxels where two correct fp32 forwards land on opposite sides."""
    d = a - b
    if sign is None:
        return torch.mean(torch.abs(d) - d)
    return torch.mean(sign.to(d.dtype) * d - d)


def l1(a, b):
    return torch.mean(torch.abs(a - b))


def cae_reconstruction_loss(lat, rec, core, penu, lesion, epoch, signs=None):
    """CaeReconstructionLearner.loss_step (:52-70).  signs: optional (sign(penu - interpolation), sign(penu - core))."""
    f = min(0.04 * max(0, epoch - 25), 1)
    s0, s1 = signs if signs is not None else (None, None)
    loss = hinge(rec['penu'], rec['interpolation'], s0) + hinge(rec['penu'], rec['core'], s1)
    loss = loss + dice_loss(rec['core'], core) + dice_loss(rec['penu'], penu) + dice_loss(rec['lesion'], lesion)
    loss = loss + f * l1(lat['interpolation'], lat['lesion'])
    return loss / (5 + f)


def cae_step_loss(rec, lesion, signs=None):
    """CaeStepLearner.loss_step (:15-21)."""
    s0 = signs[0] if signs is not None else None
    return (hinge(rec['penu'], rec['interpolation'], s0) + dice_loss(rec['interpolation'], lesion)) / 2


def cae_prediction_loss(lat_in, rec_in, lat_gt, lesion, signs=None):
    """CaePredictionLearner.loss_step (:42-57)."""
    s0, s1 = signs if signs is not None else (None, None)
    loss = hinge(rec_in['penu'], rec_in['interpolation'], s0) + hinge(rec_in['penu'], rec_in['core'], s1)
    loss = loss + dice_loss(rec_in['interpolation'], lesion)
    loss = loss + l1(lat_gt['interpolation'], lat_in['interpolation']) + l1(lat_gt['core'], lat_in['core'])
    loss = loss + l1(lat_gt['penu'], lat_in['penu'])
    return loss / 6


def unet_loss(core_out, penu_out, core, penu):
    """UnetSegmentationLearner.loss_step (:21-28)."""
    return (dice_loss(core_out, core) + dice_loss(penu_out, penu)) / 2


# ------------------------------------------------------------------------------------------------ optimizer
def binary_measures(result, target, threshold=0.5):
    """metrics.py:31-47 without the surface distances: medpy.metric.binary (MedPy==0.3.0, requirements.txt:2, not
    vendored) dc / precision / sensitivity / specificity restated from their published definitions on boolean masks."""
    r = (result > threshold).reshape(-1)
    t = (target > threshold).reshape(-1)
    tp = int((r & t).sum()); fp = int((r & ~t).sum()); fn = int((~r & t).sum()); tn = int((~r & ~t).sum())
    size_r, size_t = int(r.sum()), int(t.sum())
    dc = 2.0 * tp / float(size_r + size_t) if size_r + size_t > 0 else 0.0
    precision = tp / float(tp + fp) if tp + fp > 0 else 0.0
    sensitivity = tp / float(tp + fn) if tp + fn > 0 else 0.0
    specificity = tn / float(tn + fp) if tn + fp > 0 else 0.0
    return {"dc": dc, "precision": precision, "sensitivity": sensitivity, "specificity": specificity,
            "counts": (tp, fp, fn, tn)}


def surface_distances_medpy(result, reference, connectivity=1):
    """MedPy==0.3.0 ``medpy.metric.binary.__surface_distances`` restated on scipy (requirements.txt:2,6; MedPy is not
    vendored): border = mask XOR binary_erosion(mask, cross of `connectivity`), distances = distance_transform_edt of the
    reference border's complement read at the result's border voxels (unit voxel spacing)."""
    import numpy as np
    from scipy.ndimage import binary_erosion, distance_transform_edt, generate_binary_structure
    result = np.atleast_1d(np.asarray(result).astype(bool))
    reference = np.atleast_1d(np.asarray(reference).astype(bool))
    footprint = generate_binary_structure(result.ndim, connectivity)
    if 0 == np.count_nonzero(result):
        raise RuntimeError('The first supplied array does not contain any binary object.')
    if 0 == np.count_nonzero(reference):
        raise RuntimeError('The second supplied array does not contain any binary object.')
    result_border = result ^ binary_erosion(result, structure=footprint, iterations=1)
    reference_border = reference ^ binary_erosion(reference, structure=footprint, iterations=1)
    dt = distance_transform_edt(~reference_border, sampling=None)
    return dt[result_border]


def surface_measures(result, target, threshold=0.5):
    """hd / assd exactly as metrics.py:31-47 obtains them: thresholded uint8 arrays of the FULL shape handed in (the reference
    hands in the B x 1 x D x H x W batch), ``mpm.hd`` = max of the two directed maxima, ``mpm.assd`` = mean of the two directed
    means; both stay inf unless both masks have voxels."""
    import numpy as np
    r = (np.asarray(result) > threshold).astype(np.uint8)
    t = (np.asarray(target) > threshold).astype(np.uint8)
    if not (r.any() and t.any()):
        return {"hd": float("inf"), "assd": float("inf")}
    s1, s2 = surface_distances_medpy(r, t), surface_distances_medpy(t, r)
    return {"hd": float(max(s1.max(), s2.max())), "assd": float(np.mean((s1.mean(), s2.mean()))),
            "asd_rt": float(s1.mean()), "asd_tr": float(s2.mean()), "n_r": int(s1.size), "n_t": int(s2.size)}


def signed_distance_map(volume, threshold=0.5, outside_is_lt=True, sign=1.0):
    """test_sdm_resampling.py:16-18 (penumbra form: edt(v > thr) - edt(v < thr)) and :31-32 (core form, sign = -1:
    edt(1 - bin) - edt(bin) = -(edt(bin) - edt(~bin)))."""
    import numpy as np
    from scipy.ndimage import distance_transform_edt
    v = np.asarray(volume)
    inside = v > threshold
    outside = (v < threshold) if outside_is_lt else ~inside
    return sign * (distance_transform_edt(inside) - distance_transform_edt(outside))


# ------------------------------------------------------------------------------------------------ data transforms
def elastic_transform(image, noise, alpha=100, sigma=4):
    """ElasticDeform.elastic_transform (data.py:331-341) with the three uniform [0, 1) noise fields handed in (the reference
    draws them from a numpy RandomState): numpy volume indexed [x][y][z]; np.meshgrid's default 'xy' indexing makes the
    FIRST coordinate follow `dy` and the second `dx` (requires X == Y)."""
    import numpy as np
    from scipy.ndimage import gaussian_filter, map_coordinates
    shape = image.shape
    dx = gaussian_filter((noise[0] * 2 - 1), sigma, mode="constant", cval=0) * alpha
    dy = gaussian_filter((noise[1] * 2 - 1), sigma, mode="constant", cval=0) * alpha
    dz = gaussian_filter((noise[2] * 2 - 1), sigma, mode="constant", cval=0) * alpha * 0.22
    x, y, z = np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), np.arange(shape[2]))
    indices = np.reshape(y + dy, (-1, 1)), np.reshape(x + dx, (-1, 1)), np.reshape(z + dz, (-1, 1))
    return map_coordinates(image, indices, order=1).reshape(shape)


def resample_plane_xy(volume, scale_factor, order=0):
    """ResamplePlaneXY (data.py:354-380): scipy.ndimage.zoom of every [x][y] slice of a [x][y][z][c] array."""
    import numpy as np
    import scipy.ndimage as ndi
    sx, sy = ndi.zoom(volume[:, :, 0, 0], scale_factor, order=0).shape[0:2]
    out = np.zeros((sx, sy) + volume.shape[2:], dtype=volume.dtype)
    for c in range(volume.shape[3]):
        for z in range(volume.shape[2]):
            out[:, :, z, c] = ndi.zoom(volume[:, :, z, c], scale_factor, order=order)
    return out


def pad_images(volume, px, py, pz, value=0.0):
    """PadImages (data.py:280-296) on a [x][y][z][c] array."""
    import numpy as np
    sx, sy, sz, sc = volume.shape
    out = np.ones((sx + 2 * px, sy + 2 * py, sz + 2 * pz, sc), dtype=np.float32) * float(value)
    out[px:-px, py:-py, pz:-pz, :] = volume
    return out


def to_tensor(volume):
    """ToTensor (data.py:299-310): [x][y][z][c] -> C x Z x Y x X."""
    import numpy as np
    return np.ascontiguousarray(np.transpose(volume, (3, 2, 1, 0)))


def adam_step(p, g, m, v, step, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-5):
    """One torch.optim.Adam update on tensors, installed-torch form.  Returns (p, m, v)."""
    g = g + weight_decay * p
    m = m + (g - m) * (1 - beta1)
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * (m / denom), m, v


# ------------------------------------------------------------------------------------------------ helpers
def clone_state(sd, requires_grad=False, dtype=None):
    out = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if dtype is not None and t.is_floating_point():
            t = t.to(dtype)
        if requires_grad and t.is_floating_point() and 'running_' not in k:
            t.requires_grad_(True)
        out[k] = t
    return out


def grads_of(loss, sd):
    names = [k for k, v in sd.items() if v.requires_grad]
    gs = torch.autograd.grad(loss, [sd[k] for k in names], allow_unused=True)
    return {k: g for k, g in zip(names, gs)}


def hinge_signs(rec):
    """(sign(penu - interpolation), sign(penu - core)) of a dict / DTO-like set of reconstructions, on the CPU."""
    get = (lambda k: rec[k]) if isinstance(rec, dict) else (lambda k: getattr(rec, k))
    p = get('penu').detach().cpu()
    return torch.sign(p - get('interpolation').detach().cpu()), torch.sign(p - get('core').detach().cpu())


def ulp_perturbed(sd, names, gen):
    """Copy of `sd` with every fp32 tensor listed in `names` moved one ulp up or down at random (tests only).

    Used to SAMPLE the fp32 noise floor of an ill-conditioned case: re-running the fp32 CPU oracle on parameters that
    differ in the last bit changes every rounding decision downstream, so the spread of |fp32 - fp64| over a few such
    runs is the accuracy the reference's own fp32 arithmetic has on that case (tests/test_gpu_models.py)."""
    out = {}
    for n, v in sd.items():
        if n in names and v.dtype == torch.float32:
            up = torch.nextafter(v, torch.full_like(v, float('inf')))
            dn = torch.nextafter(v, torch.full_like(v, -float('inf')))
            out[n] = torch.where(torch.randint(0, 2, v.shape, generator=gen).bool(), up, dn)
        else:
            out[n] = v.clone()
    return out


def rel_l2(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)
