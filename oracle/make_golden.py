"""Generate tests/golden/*.npz from the REFERENCE's own modules and pin oracle/stroke_oracle.py against them.

Runs only in the build container (it needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

For each case it (1) imports the reference classes (common.model.Cae3D / Unet3D, common.metrics.BatchDiceLoss,
common.inference.*, learner.*.loss_step) with the four absent third-party modules stubbed, (2) runs them
on seeded synthetic inputs on CPU torch, forward + loss + backward + one Adam step, (3) asserts that the functional
restatement in stroke_oracle.py reproduces every output bit-for-bit (same torch ops, same order), and (4) stores the
reference's inputs / weights / outputs as small fixtures.  Large volumes are stored as strided samples plus fp64
moments; parameter gradients, latents and losses are stored in full.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = '/root/reference'
OUT = os.path.join(ROOT, 'tests', 'golden')


def import_reference():
    for name in ('nibabel', 'medpy', 'medpy.metric', 'medpy.metric.binary', 'jsonpickle', 'matplotlib', 'matplotlib.pyplot'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib'].use = lambda *a, **k: None
    sys.modules['medpy'].metric = sys.modules['medpy.metric']
    sys.modules['medpy.metric'].binary = sys.modules['medpy.metric.binary']
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.path.insert(0, REF)
    import common.model.Cae3D as RCae
    import common.model.Unet3D as RUnet
    import common.metrics as RMetrics
    import common.inference.CaeInference as RCaeInf
    import common.inference.UnetInference as RUnetInf
    import common.dto.CaeDto as RCaeDto
    import learner.CaeReconstructionLearner as RRecL
    import learner.CaeStepLearner as RStepL
    import learner.CaePredictionLearner as RPredL
    import learner.UnetSegmentationLearner as RUnetL
    return types.SimpleNamespace(Cae=RCae, Unet=RUnet, Metrics=RMetrics, CaeInf=RCaeInf, UnetInf=RUnetInf,
                                 CaeDto=RCaeDto, RecL=RRecL, StepL=RStepL, PredL=RPredL, UnetL=RUnetL)


sys.path.insert(0, HERE)
import stroke_oracle as O  # noqa: E402


def sample(t, stride=13):
    return t.detach().reshape(-1)[::stride].clone().numpy()


def moments(t):
    d = t.detach().double()
    return np.array([d.sum().item(), (d * d).sum().item(), d.abs().max().item()])


def pack_masks(m):
    return np.packbits(m.numpy().astype(np.uint8).reshape(-1)), np.array(m.shape)


def sd_np(sd, prefix):
    return {prefix + k: v.detach().clone().numpy() for k, v in sd.items()}


def randomize_bn(model, gen):
    """Non-trivial BN affine parameters and running stats so every term of the BN forward/backward is exercised."""
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm3d):
            with torch.no_grad():
                m.weight.copy_(0.5 + torch.rand(m.weight.shape, generator=gen))
                m.bias.copy_(0.2 * torch.randn(m.bias.shape, generator=gen))
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=gen))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=gen))


def ellipsoids(gen, B, size, fracs):
    D, H, W = size
    zz, yy, xx = torch.meshgrid(torch.arange(D), torch.arange(H), torch.arange(W), indexing='ij')
    out = torch.zeros(B, len(fracs), D, H, W)
    for b in range(B):
        c = (0.35 + 0.3 * torch.rand(3, generator=gen)) * torch.tensor([D, H, W])
        for j, f in enumerate(fracs):
            r = torch.tensor([D, H, W]) * f
            out[b, j] = ((((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2) <= 1).float()
    return out


WORST = {}
GRAD_TOL = 5e-4   # max-abs error relative to the tensor's max-abs value


def assert_same(a, b, what, tol=0.0):
    """tol = 0: bit-for-bit (forward tensors: same torch ops in the same order).  Gradients go through autograd's
    accumulation, whose order differs between the module graph and the functional graph, so they are compared with
    a relative tolerance; the worst relative error per case is printed."""
    if a is None and b is None:
        return
    err = (a.detach().double() - b.detach().double()).abs().max().item()
    ref = b.detach().double().abs().max().item()
    if tol > 0:
        key = what.split('.')[0]
        WORST[key] = max(WORST.get(key, 0.0), err / max(ref, 1e-30))
    if err > tol * max(ref, 1e-30):
        raise AssertionError('oracle restatement differs from the reference at %s: max abs err %g (ref max %g)' % (what, err, ref))


# ======================================================================================================== CAE cases
def cae_case(R, name, channels, size, B, epoch, seed, mode):
    """mode: 'reconstruction' (Enc3D, CaeReconstructionLearner), 'step' (frozen CAE + Enc3DStep, CaeStepLearner)."""
    D, HW = size
    alpha = 1.0
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    EncCls = R.Cae.Enc3DStep if mode == 'step' else R.Cae.Enc3D
    enc = EncCls(HW, D, channels, 5, alpha)
    dec = R.Cae.Dec3D(HW, D, channels, 5, alpha)
    cae = R.Cae.Cae3D(enc, dec)
    randomize_bn(cae, gen)
    if mode == 'step':
        cae.freeze(True)
        for p in list(enc.reduce.parameters()) + list(enc.step.parameters()):
            p.requires_grad = True
    cae.train()
    sd0 = {k: v.detach().clone() for k, v in cae.state_dict().items()}

    labels = ellipsoids(gen, B, (D, HW, HW), (0.13, 0.27, 0.19))
    clinical = torch.zeros(B, 5, 1, 1, 1, dtype=torch.float64)
    clinical[:, 0, 0, 0, 0] = 0.5 + 3.5 * torch.rand(B, generator=gen, dtype=torch.float64)
    clinical[:, 1, 0, 0, 0] = 5.0 * torch.rand(B, generator=gen, dtype=torch.float64)
    clinical[:, 2:, 0, 0, 0] = torch.rand(B, 3, generator=gen, dtype=torch.float64)
    batch = {'images': torch.zeros(B, 2, 1, 1, 1), 'labels': labels, 'clinical': clinical}

    # ---- the reference, driven through its own inference + loss_step code
    crit = R.Metrics.BatchDiceLoss([1.0])
    inf = R.CaeInf.CaeInference(cae, 10)
    if mode == 'step':
        inf.get_time_to_treatment = types.MethodType(R.StepL.CaeStepLearner.get_time_to_treatment, inf)
    dto = inf.inference_step(batch)
    holder = types.SimpleNamespace(_criterion=crit)
    if mode == 'step':
        loss = R.StepL.CaeStepLearner.loss_step(holder, dto, epoch)
    else:
        loss = R.RecL.CaeReconstructionLearner.loss_step(holder, dto, epoch)
    params = [(n, p) for n, p in cae.named_parameters() if p.requires_grad]
    opt = torch.optim.Adam([p for _, p in params], lr=1e-3, weight_decay=1e-5, betas=(0.9, 0.999))
    opt.zero_grad()
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in params}
    opt.step()
    sd1 = {k: v.detach().clone() for k, v in cae.state_dict().items()}

    # ---- the restatement on the same weights/inputs
    sdo = O.clone_state(sd0)
    for n, _ in params:
        sdo[n].requires_grad_(True)
    core, penu, lesion = labels[:, 0:1], labels[:, 1:2], labels[:, 2:3]
    step = O.step_from_globals(clinical.float(), sdo, alpha) if mode == 'step' else O.time_to_treatment(clinical)
    lat, rec = O.cae_forward(sdo, channels, alpha, True, core, penu, lesion, step)
    oloss = O.cae_step_loss(rec, lesion) if mode == 'step' else O.cae_reconstruction_loss(lat, rec, core, penu, lesion, epoch)
    ogr = O.grads_of(oloss, sdo)
    assert_same(oloss, loss, name + '.loss')
    for k in ('core', 'penu', 'lesion', 'interpolation'):
        assert_same(lat[k], getattr(dto.latents.gtruth, k), name + '.latents.' + k)
        assert_same(rec[k], getattr(dto.reconstructions.gtruth, k), name + '.reconstructions.' + k)
    for n in grads:
        assert_same(ogr[n], grads[n], name + '.grad.' + n, tol=GRAD_TOL)
    for k in sd0:
        if 'running_' in k or 'num_batches' in k:
            assert_same(sdo[k].double(), sd1[k].double(), name + '.buffer.' + k)
    for n, p in params:
        m0 = torch.zeros_like(p)
        newp, _, _ = O.adam_step(sd0[n], grads[n], m0, m0.clone(), 1)
        assert_same(newp, sd1[n], name + '.adam.' + n, tol=GRAD_TOL)

    fx = {'channels': np.array(channels), 'size': np.array([D, HW, HW]), 'B': np.array(B), 'epoch': np.array(epoch),
          'alpha': np.array(alpha), 'clinical': clinical.numpy(), 'loss': np.array(loss.item(), dtype=np.float64)}
    fx['labels_bits'], fx['labels_shape'] = pack_masks(labels)
    fx.update(sd_np(sd0, 'sd0/'))
    fx.update({'sd1/' + k: v.numpy() for k, v in sd1.items() if 'running_' in k or 'num_batches' in k or k in grads})
    fx.update({'grad/' + n: g.numpy() for n, g in grads.items()})
    fx['step'] = step.detach().numpy()
    for k in ('core', 'penu', 'lesion', 'interpolation'):
        fx['lat/' + k] = getattr(dto.latents.gtruth, k).detach().numpy()
        r = getattr(dto.reconstructions.gtruth, k)
        fx['rec_sample/' + k] = sample(r)
        fx['rec_moments/' + k] = moments(r)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **fx)
    print('\nwrote', name, 'loss', loss.item())


def prediction_case(R, name, channels, size, B, seed):
    """Config 4 with the *intended* flag semantics: new encoder on soft segmentations + frozen CAE."""
    D, HW = size
    alpha = 1.0
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    cae = R.Cae.Cae3D(R.Cae.Enc3D(HW, D, channels, 5, alpha), R.Cae.Dec3D(HW, D, channels, 5, alpha))
    new_enc = R.Cae.Enc3D(HW, D, channels, 5, alpha)
    randomize_bn(cae, gen)
    randomize_bn(new_enc, gen)
    cae.freeze(True)
    cae.train()
    new_enc.train()
    sd_cae0 = {k: v.detach().clone() for k, v in cae.state_dict().items()}
    sd_enc0 = {k: v.detach().clone() for k, v in new_enc.state_dict().items()}
    labels = ellipsoids(gen, B, (D, HW, HW), (0.13, 0.27, 0.19))
    soft = (0.8 * labels[:, :2] + 0.2 * torch.rand(B, 2, D, HW, HW, generator=gen)).clamp(0, 1)
    soft = soft.half().float()   # fp16-representable so the fixture stores the exact inputs in half the bytes
    clinical = torch.zeros(B, 5, 1, 1, 1, dtype=torch.float64)
    clinical[:, 0, 0, 0, 0] = 0.5 + 3.5 * torch.rand(B, generator=gen, dtype=torch.float64)
    clinical[:, 1, 0, 0, 0] = 5.0 * torch.rand(B, generator=gen, dtype=torch.float64)
    batch = {'images': soft, 'labels': labels, 'clinical': clinical}

    inf = R.CaeInf.CaeInference(cae, 10)
    dto = inf.init_clinical_variables(batch, None)
    dto.flag = R.CaeDto.FLAG_INPUTS
    dto.given_variables.inputs.core = soft[:, 0:1]
    dto.given_variables.inputs.penu = soft[:, 1:2]
    dto = new_enc(dto)
    dto = cae.dec(dto)
    dto.flag = R.CaeDto.FLAG_GTRUTH
    dto = inf.init_gtruth_segm_variables(batch, dto)
    dto = cae(dto)
    holder = types.SimpleNamespace(_criterion=R.Metrics.BatchDiceLoss([1.0]))
    loss = R.PredL.CaePredictionLearner.loss_step(holder, dto, 0)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in new_enc.named_parameters()}
    assert all(p.grad is None for p in cae.parameters())

    # restatement
    sdo = O.clone_state(sd_cae0)
    sde = O.clone_state({'enc.' + k: v for k, v in sd_enc0.items()}, requires_grad=True)
    step = O.time_to_treatment(clinical)
    enc_in = lambda x: O.encoder_pass(x, sde, channels, alpha, True, 'enc.encoder')
    lat_in = {'core': enc_in(soft[:, 0:1]), 'penu': enc_in(soft[:, 1:2])}
    lat_in['interpolation'] = O.interpolate(lat_in['core'], lat_in['penu'], step)
    rec_in = {k: O.decoder_pass(lat_in[k], sdo, channels, alpha, True) for k in ('core', 'penu', 'interpolation')}
    lat_gt, rec_gt = O.cae_forward(sdo, channels, alpha, True, labels[:, 0:1], labels[:, 1:2], labels[:, 2:3], step)
    oloss = O.cae_prediction_loss(lat_in, rec_in, lat_gt, labels[:, 2:3])
    ogr = O.grads_of(oloss, sde)
    assert_same(oloss, loss, name + '.loss')
    for n in grads:
        assert_same(ogr['enc.' + n], grads[n], name + '.grad.' + n, tol=GRAD_TOL)
    for k, v in cae.state_dict().items():
        if 'running_' in k:
            assert_same(sdo[k], v, name + '.buffer.' + k)

    fx = {'channels': np.array(channels), 'size': np.array([D, HW, HW]), 'B': np.array(B), 'alpha': np.array(alpha),
          'clinical': clinical.numpy(), 'loss': np.array(loss.item(), dtype=np.float64), 'soft': soft.numpy().astype(np.float16)}
    fx['labels_bits'], fx['labels_shape'] = pack_masks(labels)
    fx.update(sd_np(sd_cae0, 'cae0/'))
    fx.update(sd_np(sd_enc0, 'enc0/'))
    fx.update({'grad/' + n: g.numpy() for n, g in grads.items()})
    fx.update({'cae1/' + k: v.numpy() for k, v in cae.state_dict().items() if 'running_' in k})
    for k in ('core', 'penu', 'interpolation'):
        fx['lat_in/' + k] = getattr(dto.latents.inputs, k).detach().numpy()
        fx['rec_in_moments/' + k] = moments(getattr(dto.reconstructions.inputs, k))
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **fx)
    print('\nwrote', name, 'loss', loss.item())


# ======================================================================================================== U-Net case
def unet_case(R, name, channels, out_size, B, seed, eval_too=True):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    unet = R.Unet.Unet3D(channels)
    randomize_bn(unet, gen)
    unet.train()
    sd0 = {k: v.detach().clone() for k, v in unet.state_dict().items()}
    D, H, W = out_size
    pad = 20
    img = torch.zeros(B, 2, D + 2 * pad, H + 2 * pad, W + 2 * pad)
    img[:, 0, pad:-pad, pad:-pad, pad:-pad] = 12 * torch.rand(B, D, H, W, generator=gen)
    img[:, 1, pad:-pad, pad:-pad, pad:-pad] = 40 * torch.rand(B, D, H, W, generator=gen)
    labels = ellipsoids(gen, B, out_size, (0.2, 0.4))
    batch = {'images': img, 'labels': labels}

    inf = R.UnetInf.UnetInference(unet)
    dto = inf.inference_step(batch)
    holder = types.SimpleNamespace(_criterion=R.Metrics.BatchDiceLoss([1.0]))
    loss = R.UnetL.UnetSegmentationLearner.loss_step(holder, dto, 0)
    params = list(unet.named_parameters())
    opt = torch.optim.Adam([p for _, p in params], lr=1e-3, weight_decay=1e-5, betas=(0.99, 0.999))
    opt.zero_grad()
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in params}
    opt.step()
    sd1 = {k: v.detach().clone() for k, v in unet.state_dict().items()}

    sdo = O.clone_state(sd0, requires_grad=True)
    oc, op = O.unet_forward(sdo, img, True)
    oloss = O.unet_loss(oc, op, labels[:, 0:1], labels[:, 1:2])
    ogr = O.grads_of(oloss, sdo)
    assert_same(oloss, loss, name + '.loss')
    assert_same(oc, dto.outputs.core, name + '.core')
    assert_same(op, dto.outputs.penu, name + '.penu')
    for n in grads:
        assert_same(ogr[n], grads[n], name + '.grad.' + n, tol=GRAD_TOL)

    fx = {'channels': np.array(channels), 'out_size': np.array(out_size), 'B': np.array(B),
          'loss': np.array(loss.item(), dtype=np.float64), 'images_interior': img[:, :, pad:-pad, pad:-pad, pad:-pad].numpy(),
          'core': dto.outputs.core.detach().numpy(), 'penu': dto.outputs.penu.detach().numpy()}
    fx['labels_bits'], fx['labels_shape'] = pack_masks(labels)
    fx.update(sd_np(sd0, 'sd0/'))
    fx.update({'sd1/' + k: v.numpy() for k, v in sd1.items()})
    fx.update({'grad/' + n: g.numpy() for n, g in grads.items()})
    if eval_too:
        unet.load_state_dict(sd0)
        unet.eval()
        with torch.no_grad():
            d2 = inf.inference_step(batch)
        fx['eval_core'] = d2.outputs.core.numpy()
        fx['eval_penu'] = d2.outputs.penu.numpy()
        ec, ep = O.unet_forward(O.clone_state(sd0), img, False)
        assert_same(ec, d2.outputs.core, name + '.eval_core')
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **fx)
    print('\nwrote', name, 'loss', loss.item())


def transforms_case(name, seed=11):
    """The reference's own data transforms (common/data.py:215-380) on one small synthetic sample; pins
    oracle.elastic_transform / resample_plane_xy / pad_images / hemispheric_flip / to_tensor."""
    import random
    import common.data as RD
    rng = np.random.RandomState(seed)
    X = Y = 24
    Z = 10
    labels = (rng.rand(X, Y, Z, 2) > 0.6).astype(np.float32)
    images = rng.rand(X, Y, Z, 2).astype(np.float32) * 12
    clinical = rng.rand(1, 1, 1, 5)
    fx = {'labels': labels, 'images': images, 'clinical': clinical, 'seed': np.array(seed)}

    # ElasticDeform exactly as __call__ chains it (data.py:343-351): channel 0 creates the random state, later channels and the
    # images continue the same stream.  (The reference seeds from the wall clock; the fixture pins the seed.)
    ed = RD.ElasticDeform(alpha=100, sigma=4, apply_to_images=True)
    rs = np.random.RandomState(seed + 1)
    out_l = np.zeros_like(labels)
    out_i = np.zeros_like(images)
    for c in range(labels.shape[3]):
        out_l[:, :, :, c], rs = ed.elastic_transform(labels[:, :, :, c].copy(), 100, 4, random_state=rs)
    for c in range(images.shape[3]):
        out_i[:, :, :, c], rs = ed.elastic_transform(images[:, :, :, c].copy(), 100, 4, random_state=rs)
    fx['elastic_labels'], fx['elastic_images'], fx['elastic_seed'] = out_l, out_i, np.array(seed + 1)
    # restatement check: same noise stream -> same result
    rs2 = np.random.RandomState(seed + 1)
    for c in range(labels.shape[3]):
        noise = [rs2.rand(X, Y, Z) for _ in range(3)]
        mine = O.elastic_transform(labels[:, :, :, c], noise, 100, 4)
        assert np.array_equal(mine, out_l[:, :, :, c]), 'elastic restatement differs from the reference'

    class Arr(np.ndarray):
        """`array != []` (data.py:367 etc.) was a scalar True on the reference's numpy 1.14; current numpy broadcasts and fails."""
        def __ne__(self, other):
            return True if isinstance(other, list) else np.ndarray.__ne__(self, other)

    def fresh():
        return {RD.KEY_IMAGES: images.copy().view(Arr), RD.KEY_LABELS: labels.copy().view(Arr),
                RD.KEY_GLOBAL: clinical.copy().view(Arr), RD.KEY_CASE_ID: 3}

    # (the reference's ResamplePlaneXY sizes its output by slicing the input, data.py:369, so it only works for factors <= 1)
    for sf, mode in ((0.5, 'nearest'), (0.75, 'nearest'), (0.5, 'bilinear'), (0.75, 'bilinear')):
        r = RD.ResamplePlaneXY(sf, mode)(fresh())
        key = 'zoom_%s_%s' % (str(sf).replace('.', 'p'), mode)
        fx[key + '_images'], fx[key + '_labels'] = np.array(r[RD.KEY_IMAGES]), np.array(r[RD.KEY_LABELS])
        assert np.array_equal(O.resample_plane_xy(images, sf, 0 if mode == 'nearest' else 1), np.array(r[RD.KEY_IMAGES]))
    sample = fresh()
    p = RD.PadImages(3, 2, 1, pad_value=0.5)(fresh())
    fx['pad_images'] = np.array(p[RD.KEY_IMAGES])
    assert np.array_equal(O.pad_images(images, 3, 2, 1, 0.5), np.array(p[RD.KEY_IMAGES]))
    random.seed(5)
    flips = []
    for _ in range(4):
        f = RD.HemisphericFlip()(fresh())
        flips.append(not np.array_equal(np.array(f[RD.KEY_LABELS]), labels))
    fx['flip_decisions_seed5'] = np.array(flips)
    fx['flipped_labels'] = np.flip(labels, RD.DIM_HORIZONTAL_NUMPY_3D).copy()
    ff = RD.HemisphericFlipFixedToCaseId(split_id=2)(fresh())
    assert np.array_equal(np.array(ff[RD.KEY_LABELS]), fx['flipped_labels'])
    random.seed(9)
    rp = RD.RandomPatch(16, 12, 6, 2, 1, 1)(fresh())
    fx['patch_images_seed9'], fx['patch_labels_seed9'] = np.array(rp[RD.KEY_IMAGES]), np.array(rp[RD.KEY_LABELS])
    tt = RD.ToTensor()(fresh())
    fx['to_tensor_labels'] = tt[RD.KEY_LABELS].contiguous().numpy()
    assert np.array_equal(O.to_tensor(labels), fx['to_tensor_labels'])
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **fx)


def main():
    os.makedirs(OUT, exist_ok=True)
    R = import_reference()
    transforms_case('transforms_tiny')
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    tiny = [1, 4, 6, 8, 10, 12, 1]
    cae_case(R, 'cae_rec_tiny', tiny, (28, 56), 2, 30, 4, 'reconstruction')
    cae_case(R, 'cae_step_tiny', tiny, (28, 56), 2, 0, 5, 'step')
    prediction_case(R, 'cae_pred_tiny', tiny, (28, 56), 2, 6)
    unet_case(R, 'unet_tiny', [2, 4, 6, 8, 6, 4, 6, 2], (4, 8, 12), 2, 7)
    print('worst relative gradient/Adam deviation restatement vs reference per case:', WORST)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, 'KiB')


if __name__ == '__main__':
    main()
