"""GPU parity, model level: the drop-in modules / learners on cuda:0 against (a) the committed fixtures produced by
the reference's own modules (tests/golden, see oracle/make_golden.py) and (b) the CPU oracle on fresh seeded inputs,
including the BASELINE.json shapes.  Everything goes through the public host API -> autograd nodes -> C-ABI."""
import numpy as np
import pytest
import torch

import stroke_oracle as O
from util import TOL_ACT, TOL_DICE, TOL_GRAD, load, rel_l2, rel_max, state_from, unpack_masks

pytestmark = pytest.mark.gpu


def _api():
    from stroke_prediction_b200.common import data
    from stroke_prediction_b200.common.inference.CaeEncInference import CaeEncInference
    from stroke_prediction_b200.common.inference.CaeInference import CaeInference
    from stroke_prediction_b200.common.inference.UnetInference import UnetInference
    from stroke_prediction_b200.common.metrics import BatchDiceLoss
    from stroke_prediction_b200.common.model.Cae3D import Cae3D, Dec3D, Enc3D, Enc3DCtp, Enc3DStep
    from stroke_prediction_b200.common.model.Unet3D import Unet3D
    from stroke_prediction_b200.learner.CaePredictionLearner import CaePredictionLearner
    from stroke_prediction_b200.learner.CaeReconstructionLearner import CaeReconstructionLearner
    from stroke_prediction_b200.learner.CaeStepLearner import CaeStepLearner
    from stroke_prediction_b200.learner.UnetSegmentationLearner import UnetSegmentationLearner
    from stroke_prediction_b200.optim import FusedAdam
    from stroke_prediction_b200.tester.Tester import Tester
    import types
    return types.SimpleNamespace(**locals())


def _dice_binary(a, b):
    a, b = (torch.as_tensor(a) > 0.5), (torch.as_tensor(b) > 0.5)
    den = a.sum().item() + b.sum().item()
    return 2.0 * (a & b).sum().item() / den if den else 0.0


def _check_grads(model, g64, g32, tol=TOL_GRAD, strip=""):
    """Noise-floor rule: against an fp64 oracle run, the CUDA gradient may not be further away than
    max(tol, 2 x the distance of the reference's own fp32 CPU gradient)."""
    n = 0
    for name, p in model.named_parameters():
        key = strip + name
        if g64.get(key) is None:
            assert p.grad is None or not p.requires_grad, name
            continue
        assert p.grad is not None, name
        e_gpu, e_cpu = rel_l2(p.grad, g64[key]), rel_l2(g32[key], g64[key])
        assert e_gpu <= max(tol, 2 * e_cpu), "%s: gpu-vs-fp64 %g, cpu32-vs-fp64 %g" % (name, e_gpu, e_cpu)
        n += 1
    assert n > 0


def _fixture_grads(fx, prefix="grad/"):
    return {k[len(prefix):]: torch.from_numpy(np.array(v)) for k, v in fx.items() if k.startswith(prefix)}


def _cae_oracle_grads(fx, mode, dtype=torch.float64, signs=None, sd0=None):
    """Oracle run (fp64 or fp32) of a CAE fixture case -> ({param: grad}, reconstructions).

    `signs`: hinge decisions taken from the implementation under test (O.hinge_signs).  The monotonicity hinge
    mean(|d| - d) has a kink at d = 0 and, for near-identical penumbra / core / interpolation reconstructions, many
    voxels with |d| of a few fp32 ulps: two correct fp32 forwards put some of them on opposite sides, which moves whole
    -2/N gradient quanta around and would dominate a gradient comparison.  Aligning the decision makes the comparison
    measure the backward arithmetic; `_check_signs` separately proves that decisions only differ at such near-ties.
    `sd0`: parameters to use instead of the fixture's (noise-floor sampling, see `_fp32_noise_floor`)."""
    ch = [int(c) for c in fx["channels"]]
    labels = unpack_masks(fx).to(dtype)
    clinical = torch.from_numpy(fx["clinical"])
    sd = O.clone_state(state_from(fx, "sd0/") if sd0 is None else sd0, dtype=dtype)
    names = list(_fixture_grads(fx))
    for n in names:
        sd[n].requires_grad_(True)
    core, penu, lesion = labels[:, 0:1], labels[:, 1:2], labels[:, 2:3]
    if mode == "step":
        step = O.step_from_globals(clinical.to(dtype), sd, 1.0)
    else:
        step = O.time_to_treatment(clinical).to(dtype)
    lat, rec = O.cae_forward(sd, ch, 1.0, True, core, penu, lesion, step)
    loss = (O.cae_step_loss(rec, lesion, signs) if mode == "step"
            else O.cae_reconstruction_loss(lat, rec, core, penu, lesion, int(fx["epoch"]), signs))
    return O.grads_of(loss, sd), rec


NOISE_SAMPLES = 6


def _fp32_noise_floor(fx, mode, samples=NOISE_SAMPLES):
    """Per-tensor fp32 noise floor of a fixture case: the LARGEST distance |fp32 CPU oracle - fp64 oracle| over the
    fixture's parameters and `samples - 1` copies of them moved by one ulp at random (O.ulp_perturbed).

    Why a sampled floor: the tiny fixtures decode a 1x1x1 latent at batch 2, i.e. the decoder's first BatchNorm
    normalises TWO values per channel with |mean|/std up to 56 (profiles/r01_diag_fwd_units.log).  Every fp32
    implementation is ill-conditioned there, and the distance of one fp32 run from fp64 is a draw from a wide
    distribution: the reference's own CPU arithmetic lands anywhere between 1.5e-5 and 1.4e-4 (encoder weights, median
    over tensors) when its parameters move by one ulp.  A single CPU draw is therefore not a bound for another correct
    implementation; the maximum over a few draws is."""
    names = set(_fixture_grads(fx))
    gen = torch.Generator().manual_seed(7)
    base = state_from(fx, "sd0/")
    floor = {}
    for k in range(samples):
        sd0 = base if k == 0 else O.ulp_perturbed(base, names, gen)
        g32, rec32 = _cae_oracle_grads(fx, mode, torch.float32, sd0=sd0)
        g64, _ = _cae_oracle_grads(fx, mode, torch.float64, O.hinge_signs(rec32), sd0=sd0)
        for n in names:
            if g64.get(n) is not None:
                floor[n] = max(floor.get(n, 0.0), rel_l2(g32[n], g64[n]))
    return floor


NEAR_TIE = 1e-5   # |d| below which a hinge decision may legitimately differ between two fp32 forwards (outputs in [0,1])


def _check_signs(signs, rec64, pairs=(("penu", "interpolation"), ("penu", "core"))):
    """Hinge decisions of the implementation under test may differ from the fp64 oracle's only at near-ties."""
    for s, (a, b) in zip(signs, pairs):
        d = (rec64[a] - rec64[b]).detach()
        differ = torch.sign(d) != s.to(d.dtype)
        assert float(d[differ].abs().max() if differ.any() else 0.0) < NEAR_TIE, (a, b, int(differ.sum()))


def _aligned_cae_check(model, fx, mode, rec_gpu, tol=TOL_GRAD, strip=""):
    """Noise-floor gradient check with the hinge decisions aligned: CUDA path vs fp64 oracle (CUDA's decisions) against
    the sampled fp32 floor of the reference arithmetic (`_fp32_noise_floor`): err_gpu <= max(tol, 2 x floor)."""
    s_gpu = O.hinge_signs(rec_gpu)
    g64_gpu, rec64 = _cae_oracle_grads(fx, mode, torch.float64, s_gpu)
    _check_signs(s_gpu, rec64)
    floor = _fp32_noise_floor(fx, mode)
    n, bad = 0, []
    for name, p in model.named_parameters():
        key = strip + name
        if g64_gpu.get(key) is None:
            assert p.grad is None or not p.requires_grad, name
            continue
        assert p.grad is not None, name
        e_gpu, e_cpu = rel_l2(p.grad, g64_gpu[key]), floor[key]
        print("%-28s gpu/f64 %.2e cpu32-floor/f64 %.2e |g| %.3e" % (name, e_gpu, e_cpu, g64_gpu[key].norm().item()))
        if e_gpu > max(tol, 2 * e_cpu):
            bad.append("%s: gpu-vs-fp64 %g, cpu32 floor %g" % (name, e_gpu, e_cpu))
        n += 1
    assert n > 0 and not bad, bad


def test_cae_reconstruction_step_against_reference_fixture():
    A = _api()
    fx = load("cae_rec_tiny")
    ch = [int(c) for c in fx["channels"]]
    D, H, W = (int(v) for v in fx["size"])
    cae = A.Cae3D(A.Enc3D(H, D, ch, 5, float(fx["alpha"])), A.Dec3D(H, D, ch, 5, float(fx["alpha"])))
    cae.load_state_dict(state_from(fx, "sd0/"))
    cae = cae.cuda().train()
    batch = {A.data.KEY_IMAGES: torch.zeros(2, 2, 1, 1, 1), A.data.KEY_LABELS: unpack_masks(fx),
             A.data.KEY_GLOBAL: torch.from_numpy(fx["clinical"])}
    opt = torch.optim.Adam([p for p in cae.parameters() if p.requires_grad], lr=1e-3, weight_decay=1e-5, betas=(0.9, 0.999))
    learner = A.CaeReconstructionLearner(None, None, cae, opt, None, 1, None, "/tmp/x", A.BatchDiceLoss([1.0]))
    dto = learner.inference_step(batch)
    assert rel_max(dto.given_variables.time_to_treatment, fx["step"]) < 1e-6
    loss = learner.loss_step(dto, int(fx["epoch"]))
    assert abs(loss.item() - float(fx["loss"])) < 1e-5
    for k in ("core", "penu", "lesion", "interpolation"):
        assert rel_l2(getattr(dto.latents.gtruth, k), fx["lat/" + k]) < TOL_ACT, k
        r = getattr(dto.reconstructions.gtruth, k)
        assert r.shape == (2, 1, D, H, W)
        assert rel_l2(r.detach().cpu().reshape(-1)[::13], fx["rec_sample/" + k]) < TOL_ACT, k
        m = fx["rec_moments/" + k]
        assert abs(r.double().sum().item() - m[0]) < 1e-4 * abs(m[0])
        assert abs((r.double() ** 2).sum().item() - m[1]) < 1e-4 * abs(m[1])
    learner._optimizer.zero_grad()
    loss.backward()
    _aligned_cae_check(cae, fx, "reconstruction", dto.reconstructions.gtruth)
    learner._optimizer.step()
    torch.cuda.synchronize()
    sd1 = cae.state_dict()
    for k, v in fx.items():
        if not k.startswith("sd1/"):
            continue
        name = k[4:]
        if "num_batches" in name:
            assert int(sd1[name]) == int(v), name
        elif "running_" in name:
            assert rel_max(sd1[name], v) < 1e-5, name
        else:   # parameters after one Adam step: compare the UPDATE (Adam's first step is lr * sign-like, so be lenient)
            before = state_from(fx, "sd0/")[name]
            upd_ref = torch.from_numpy(np.array(v)) - before
            upd = sd1[name].cpu() - before
            assert (upd - upd_ref).abs().max().item() < 2e-5, name


def test_cae_step_learner_against_reference_fixture():
    A = _api()
    fx = load("cae_step_tiny")
    ch = [int(c) for c in fx["channels"]]
    D, H, W = (int(v) for v in fx["size"])
    cae = A.Cae3D(A.Enc3DStep(H, D, ch, 5, 1.0), A.Dec3D(H, D, ch, 5, 1.0))
    cae.load_state_dict(state_from(fx, "sd0/"))
    cae.freeze(True)
    for p in list(cae.enc.reduce.parameters()) + list(cae.enc.step.parameters()):
        p.requires_grad = True
    cae = cae.cuda().train()
    batch = {A.data.KEY_IMAGES: torch.zeros(2, 2, 1, 1, 1), A.data.KEY_LABELS: unpack_masks(fx),
             A.data.KEY_GLOBAL: torch.from_numpy(fx["clinical"])}
    opt = A.FusedAdam([p for p in cae.parameters() if p.requires_grad], lr=1e-3, weight_decay=1e-5)
    learner = A.CaeStepLearner(None, None, cae, opt, None, 1, None, "/tmp/x", A.BatchDiceLoss([1.0]))
    dto = learner.inference_step(batch)
    assert dto.given_variables.time_to_treatment is None
    loss = learner.loss_step(dto, 0)
    assert abs(loss.item() - float(fx["loss"])) < 1e-5
    assert rel_l2(dto.latents.gtruth.interpolation, fx["lat/interpolation"]) < TOL_ACT
    loss.backward()
    _aligned_cae_check(cae, fx, "step", dto.reconstructions.gtruth)
    frozen = [n for n, p in cae.named_parameters() if not p.requires_grad]
    assert frozen and all(dict(cae.named_parameters())[n].grad is None for n in frozen)
    # frozen BN layers still ran in train mode: running stats drift exactly like the reference
    for k, v in fx.items():
        if k.startswith("sd1/") and "running_" in k:
            assert rel_max(cae.state_dict()[k[4:]], v) < 1e-5, k


def test_cae_prediction_learner_against_reference_fixture():
    A = _api()
    fx = load("cae_pred_tiny")
    ch = [int(c) for c in fx["channels"]]
    D, H, W = (int(v) for v in fx["size"])
    cae = A.Cae3D(A.Enc3D(H, D, ch, 5, 1.0), A.Dec3D(H, D, ch, 5, 1.0))
    cae.load_state_dict(state_from(fx, "cae0/"))
    new_enc = A.Enc3D(H, D, ch, 5, 1.0)
    new_enc.load_state_dict(state_from(fx, "enc0/"))
    cae, new_enc = cae.cuda().train(), new_enc.cuda().train()
    batch = {A.data.KEY_IMAGES: torch.from_numpy(fx["soft"].astype(np.float32)), A.data.KEY_LABELS: unpack_masks(fx),
             A.data.KEY_GLOBAL: torch.from_numpy(fx["clinical"])}
    opt = A.FusedAdam(new_enc.parameters(), lr=1e-3, weight_decay=1e-5)
    learner = A.CaePredictionLearner(None, None, cae, new_enc, opt, None, 1, None, "/tmp/x", A.BatchDiceLoss([1.0]))
    dto = learner.inference_step(batch)
    loss = learner.loss_step(dto, 0)
    assert abs(loss.item() - float(fx["loss"])) < 1e-5
    for k in ("core", "penu", "interpolation"):
        assert rel_l2(getattr(dto.latents.inputs, k), fx["lat_in/" + k]) < TOL_ACT
    loss.backward()
    # fp64 oracle of the same step
    sd_cae = O.clone_state(state_from(fx, "cae0/"), dtype=torch.float64)
    sd_enc = O.clone_state({"enc." + k: v for k, v in state_from(fx, "enc0/").items()}, requires_grad=True, dtype=torch.float64)
    labels64 = unpack_masks(fx).double()
    soft64 = torch.from_numpy(fx["soft"].astype(np.float32)).double()
    step64 = O.time_to_treatment(torch.from_numpy(fx["clinical"])).double()
    e64 = lambda x: O.encoder_pass(x, sd_enc, ch, 1.0, True, "enc.encoder")
    li = {"core": e64(soft64[:, 0:1]), "penu": e64(soft64[:, 1:2])}
    li["interpolation"] = O.interpolate(li["core"], li["penu"], step64)
    ri = {k: O.decoder_pass(li[k], sd_cae, ch, 1.0, True) for k in ("core", "penu", "interpolation")}
    lg, _ = O.cae_forward(sd_cae, ch, 1.0, True, labels64[:, 0:1], labels64[:, 1:2], labels64[:, 2:3], step64)
    signs = O.hinge_signs(dto.reconstructions.inputs)
    _check_signs(signs, ri)
    g64 = O.grads_of(O.cae_prediction_loss(li, ri, lg, labels64[:, 2:3], signs), sd_enc)
    _check_grads(new_enc, g64, {"enc." + k: v for k, v in _fixture_grads(fx).items()}, strip="enc.")
    assert all(p.grad is None for p in cae.parameters())
    for k, v in fx.items():
        if k.startswith("cae1/"):
            assert rel_max(cae.state_dict()[k[5:]], v) < 1e-5, k
    # the reference's literal flag bug is available behind the switch and fails like the reference does
    strict = A.CaeEncInference(cae, new_enc, 10, compat_reference_flag_bug=True)
    with pytest.raises(AssertionError):
        strict.inference_step(batch)


def test_unet_step_against_reference_fixture():
    A = _api()
    fx = load("unet_tiny")
    unet = A.Unet3D([int(c) for c in fx["channels"]])
    unet.load_state_dict(state_from(fx, "sd0/"))
    unet = unet.cuda().train()
    interior = torch.from_numpy(fx["images_interior"])
    B, _, D, H, W = interior.shape
    img = torch.zeros(B, 2, D + 40, H + 40, W + 40)
    img[:, :, 20:-20, 20:-20, 20:-20] = interior
    batch = {A.data.KEY_IMAGES: img, A.data.KEY_LABELS: unpack_masks(fx)}
    opt = torch.optim.Adam(unet.parameters(), lr=1e-3, weight_decay=1e-5, betas=(0.99, 0.999))
    learner = A.UnetSegmentationLearner(None, None, unet, opt, None, 1, A.BatchDiceLoss([1.0]))
    dto = learner.inference_step(batch)
    assert dto.outputs.core.shape == (B, 1, D, H, W)
    assert rel_l2(dto.outputs.core, fx["core"]) < TOL_ACT and rel_l2(dto.outputs.penu, fx["penu"]) < TOL_ACT
    assert abs(_dice_binary(dto.outputs.penu.cpu(), fx["penu"]) - 1.0) < TOL_DICE
    loss = learner.loss_step(dto, 0)
    assert abs(loss.item() - float(fx["loss"])) < 1e-5
    learner._optimizer.zero_grad()
    loss.backward()
    # U-Net gradients: judge against an fp64 oracle run with the CPU-fp32 noise floor (SURVEY fact 9 / §8c rule 3)
    sd64 = O.clone_state(state_from(fx, "sd0/"), requires_grad=True, dtype=torch.float64)
    labels = unpack_masks(fx).double()
    c64, p64 = O.unet_forward(sd64, img.double(), True)
    g64 = O.grads_of(O.unet_loss(c64, p64, labels[:, 0:1], labels[:, 1:2]), sd64)
    for n, p in unet.named_parameters():
        e_gpu, e_cpu = rel_l2(p.grad, g64[n]), rel_l2(fx["grad/" + n], g64[n])
        assert e_gpu <= max(TOL_GRAD, 2 * e_cpu), "%s: gpu %g cpu32 %g" % (n, e_gpu, e_cpu)
    learner._optimizer.step()
    for k, v in fx.items():
        if k.startswith("sd1/") and "running_" in k:
            assert rel_max(unet.state_dict()[k[4:]], v) < 1e-5, k
    # eval mode (running statistics) — Tester path
    unet.load_state_dict(state_from(fx, "sd0/"))
    unet.eval()
    with torch.no_grad():
        d2 = A.UnetInference(unet).inference_step(batch)
    assert rel_l2(d2.outputs.core, fx["eval_core"]) < TOL_ACT and rel_l2(d2.outputs.penu, fx["eval_penu"]) < TOL_ACT


def test_three_training_steps_track_the_oracle():
    """Loss trajectory of Learner.train_batch (forward, loss, backward, fused Adam, beta1 schedule) vs the oracle loop."""
    A = _api()
    ch = [1, 4, 6, 8, 10, 12, 1]
    torch.manual_seed(21)
    cae = A.Cae3D(A.Enc3D(56, 28, ch, 5, 1.0), A.Dec3D(56, 28, ch, 5, 1.0))
    sd = O.clone_state(cae.state_dict(), requires_grad=True)
    cae = cae.cuda().train()
    opt = torch.optim.Adam(cae.parameters(), lr=1e-3, weight_decay=1e-5, betas=(0.9, 0.999))
    learner = A.CaeReconstructionLearner(None, None, cae, opt, None, 10, None, "/tmp/x", A.BatchDiceLoss([1.0]))
    names = [k for k, v in sd.items() if v.requires_grad]
    m = {k: torch.zeros_like(sd[k]) for k in names}
    v = {k: torch.zeros_like(sd[k]) for k in names}
    for it in range(3):
        epoch = it          # adapt_betas: beta1 = 0.5, 0.6, 0.7 (CaeReconstructionLearner.py:28-40)
        learner.adapt_betas(epoch)
        batch = A.data.synthetic_cae_batch(2, size=(28, 56, 56), seed=40 + it)
        got = learner.train_batch(batch, epoch).loss
        labels = batch[A.data.KEY_LABELS]
        step = O.time_to_treatment(batch[A.data.KEY_GLOBAL])
        lat, rec = O.cae_forward(sd, ch, 1.0, True, labels[:, 0:1], labels[:, 1:2], labels[:, 2:3], step)
        loss = O.cae_reconstruction_loss(lat, rec, labels[:, 0:1], labels[:, 1:2], labels[:, 2:3], epoch)
        grads = O.grads_of(loss, sd)
        beta1 = 0.9 - 0.1 * (4 - epoch)
        with torch.no_grad():
            for k in names:
                newp, m[k], v[k] = O.adam_step(sd[k], grads[k], m[k], v[k], it + 1, beta1=beta1)
                sd[k].copy_(newp)
        # step 0 is a pure forward comparison; afterwards Adam's first updates are ~lr*sign(g), so parameters whose
        # gradient is at the fp32 noise level move in opposite directions on the two platforms and the trajectories
        # separate at the 1e-4 level (the same happens between two CPU runs with different thread counts)
        tol = 1e-5 if it == 0 else 2e-3
        assert abs(got - loss.item()) < tol * max(1.0, abs(loss.item())), (it, got, loss.item())


def test_overlapped_batch_metrics_equal_serial_ones():
    """train_batch launches the per-batch metrics (counts + exact-EDT surface distances) on a side stream before the backward
    pass and reads them back after the optimizer step: same metrics, loss and parameters as the serial order of Learner.py:116-130."""
    A = _api()
    ch = [1, 4, 6, 8, 10, 12, 1]
    runs = []
    for overlap in (True, False):
        torch.manual_seed(22)
        cae = A.Cae3D(A.Enc3D(56, 28, ch, 5, 1.0), A.Dec3D(56, 28, ch, 5, 1.0)).cuda().train()
        opt = torch.optim.Adam(cae.parameters(), lr=1e-3, weight_decay=1e-5, betas=(0.9, 0.999))
        learner = A.CaeReconstructionLearner(None, None, cae, opt, None, 10, None, "/tmp/x", A.BatchDiceLoss([1.0]))
        learner.OVERLAP_METRICS = overlap
        rows = []
        for it in range(3):
            batch = A.data.synthetic_cae_batch(2, size=(28, 56, 56), seed=60 + it)
            m = learner.train_batch(batch, it)
            rows.append([m.loss] + [getattr(part, f) for part in (m.lesion, m.core, m.penu)
                                    for f in ("dc", "hd", "assd", "precision", "sensitivity", "specificity")])
        torch.cuda.synchronize()
        runs.append((rows, [p.detach().clone() for p in cae.parameters()]))
    assert runs[0][0] == runs[1][0], (runs[0][0], runs[1][0])
    assert any(np.isfinite(r[2]) for r in runs[0][0]), "surface distances were expected to be computed"
    for a, b in zip(runs[0][1], runs[1][1]):
        assert torch.equal(a, b)


def test_prefetched_epoch_loop_equals_plain_train_batch():
    """Learner.train_batches (the epoch loop of run_training: host-to-device copy of batch i + 1 on a copy stream while batch i
    trains) gives the same metrics and parameters as calling train_batch on the host batches one after the other."""
    A = _api()
    ch = [1, 4, 6, 8, 10, 12, 1]
    batches = [{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in A.data.synthetic_cae_batch(2, size=(28, 56, 56), seed=80 + i).items()}
               for i in range(4)]
    runs = []
    for prefetch in (True, False):
        torch.manual_seed(24)
        cae = A.Cae3D(A.Enc3D(56, 28, ch, 5, 1.0), A.Dec3D(56, 28, ch, 5, 1.0)).cuda().train()
        opt = torch.optim.Adam(cae.parameters(), lr=1e-3, weight_decay=1e-5, betas=(0.9, 0.999))
        learner = A.CaeReconstructionLearner(None, None, cae, opt, None, 10, None, "/tmp/x", A.BatchDiceLoss([1.0]))
        ms = list(learner.train_batches(batches, 3)) if prefetch else [learner.train_batch(b, 3) for b in batches]
        torch.cuda.synchronize()
        runs.append(([(m.loss, m.lesion.dc, m.core.dc, m.penu.dc, m.lesion.hd) for m in ms], [p.detach().clone() for p in cae.parameters()]))
    assert runs[0][0] == runs[1][0], (runs[0][0], runs[1][0])
    for a, b in zip(runs[0][1], runs[1][1]):
        assert torch.equal(a, b)


def test_tester_eval_batch_one():
    """Tester path: eval mode, batch size 1 (breaks in the reference on current torch)."""
    A = _api()
    ch = [1, 4, 6, 8, 10, 12, 1]
    torch.manual_seed(23)
    cae = A.Cae3D(A.Enc3D(56, 28, ch, 5, 1.0), A.Dec3D(56, 28, ch, 5, 1.0))
    sd = O.clone_state(cae.state_dict())
    cae = cae.cuda()

    class T(A.Tester, A.CaeInference):
        def __init__(self, model):
            A.Tester.__init__(self, None, model, "/tmp/x")
            A.CaeInference.__init__(self, model, 10)

    t = T(cae)
    batch = A.data.synthetic_cae_batch(1, size=(28, 56, 56), seed=3)
    _, dto = t.infer_batch(batch)
    labels = batch[A.data.KEY_LABELS]
    step = O.time_to_treatment(batch[A.data.KEY_GLOBAL])
    lat, rec = O.cae_forward(sd, ch, 1.0, False, labels[:, 0:1], labels[:, 1:2], labels[:, 2:3], step)
    for k in ("core", "penu", "lesion", "interpolation"):
        assert rel_l2(getattr(dto.reconstructions.gtruth, k), rec[k]) < TOL_ACT
    assert not any(p.requires_grad for p in cae.parameters())


def test_enc3dctp_virtual_concat():
    A = _api()
    ch = [3, 4, 6, 8, 10, 12, 1]
    torch.manual_seed(25)
    enc = A.Enc3DCtp(56, 28, ch, 5, 1.0, [20, 20, 20])
    sd = O.clone_state({"enc." + k: v for k, v in enc.state_dict().items()})
    enc = enc.cuda().train()
    from stroke_prediction_b200.common.dto import CaeDto as U
    mask = (torch.rand(2, 3, 28, 56, 56) > 0.7).float()
    cbv, ttd = torch.rand(2, 1, 68, 96, 96), torch.rand(2, 1, 68, 96, 96)
    step = torch.rand(2, 1, 1, 1, 1)
    dto = U.init_dto(None, step.cuda(), None, None, cbv.cuda(), ttd.cuda(), mask[:, 0:1].cuda(), mask[:, 1:2].cuda(), mask[:, 2:3].cuda())
    dto = enc(dto)
    crop = lambda t: t[:, :, 20:-20, 20:-20, 20:-20]
    for j, k in enumerate(("core", "penu", "lesion")):
        ref = O.encoder_pass(torch.cat((mask[:, j:j + 1], crop(cbv), crop(ttd)), 1), sd, ch, 1.0, True, "enc.encoder")
        assert rel_l2(getattr(dto.latents.gtruth, k), ref) < TOL_ACT, k


@pytest.mark.parametrize("fc", [200, 800])
def test_cae_named_config_full_size_against_oracle(fc):
    """BASELINE configs 2/3: channels 1 16 24 32 100 fc 1 on 28 x 128 x 128, B = 2: forward, loss and gradients."""
    A = _api()
    ch = [1, 16, 24, 32, 100, fc, 1]
    torch.manual_seed(31)
    cae = A.Cae3D(A.Enc3D(128, 28, ch, 5, 1.0), A.Dec3D(128, 28, ch, 5, 1.0))
    sd = O.clone_state(cae.state_dict(), requires_grad=True)
    cae = cae.cuda().train()
    batch = A.data.synthetic_cae_batch(2, seed=4)
    opt = A.FusedAdam(cae.parameters(), lr=1e-3, weight_decay=1e-5)
    learner = A.CaeReconstructionLearner(None, None, cae, opt, None, 1, None, "/tmp/x", A.BatchDiceLoss([1.0]))
    dto = learner.inference_step(batch)
    loss = learner.loss_step(dto, 60)
    loss.backward()
    labels = batch[A.data.KEY_LABELS]
    step = O.time_to_treatment(batch[A.data.KEY_GLOBAL])
    lat, rec = O.cae_forward(sd, ch, 1.0, True, labels[:, 0:1], labels[:, 1:2], labels[:, 2:3], step)
    oloss = O.cae_reconstruction_loss(lat, rec, labels[:, 0:1], labels[:, 1:2], labels[:, 2:3], 60)
    grads = O.grads_of(oloss, sd)
    l64 = labels.double()

    def grads64_with(signs):
        sd64 = O.clone_state({k: v.detach() for k, v in sd.items()}, requires_grad=True, dtype=torch.float64)
        lat64, rec64 = O.cae_forward(sd64, ch, 1.0, True, l64[:, 0:1], l64[:, 1:2], l64[:, 2:3], step.double())
        return O.grads_of(O.cae_reconstruction_loss(lat64, rec64, l64[:, 0:1], l64[:, 1:2], l64[:, 2:3], 60, signs), sd64), rec64

    s_gpu = O.hinge_signs(dto.reconstructions.gtruth)
    grads64, rec64 = grads64_with(s_gpu)                 # hinge decisions of the CUDA forward (see _cae_oracle_grads)
    _check_signs(s_gpu, rec64)
    grads64_cpu, _ = grads64_with(O.hinge_signs(rec))    # noise floor: CPU fp32 against fp64 with the CPU run's decisions
    assert abs(loss.item() - oloss.item()) < 1e-5
    for k in ("core", "penu", "lesion", "interpolation"):
        assert rel_l2(getattr(dto.latents.gtruth, k), lat[k]) < TOL_ACT, k
        r = getattr(dto.reconstructions.gtruth, k)
        assert rel_l2(r, rec[k]) < TOL_ACT, k
        assert abs(_dice_binary(r.cpu(), rec[k]) - 1.0) < TOL_DICE or float((rec[k] > 0.5).sum()) == 0
    n = 0
    for name, p in cae.named_parameters():
        e_gpu, e_cpu = rel_l2(p.grad, grads64[name]), rel_l2(grads[name], grads64_cpu[name])
        assert e_gpu <= max(TOL_GRAD, 2 * e_cpu), "%s: gpu-vs-fp64 %g, cpu32-vs-fp64 %g" % (name, e_gpu, e_cpu)
        n += 1
    assert n > 0


def test_unet_named_config_patch_size_against_oracle():
    """BASELINE config 1 channels on the reference's training patch 2 x 68 x 104 x 104 -> 28 x 64 x 64, B = 2."""
    A = _api()
    torch.manual_seed(33)
    unet = A.Unet3D([2, 16, 32, 64, 32, 16, 32, 2])
    sd = O.clone_state(unet.state_dict())
    unet = unet.cuda().train()
    batch = A.data.synthetic_unet_batch(2, out_size=(28, 64, 64), seed=4)
    opt = A.FusedAdam(unet.parameters(), lr=1e-3, weight_decay=1e-5, betas=(0.99, 0.999))
    learner = A.UnetSegmentationLearner(None, None, unet, opt, None, 1, A.BatchDiceLoss([1.0]))
    dto = learner.inference_step(batch)
    loss = learner.loss_step(dto, 0)
    loss.backward()
    labels = batch[A.data.KEY_LABELS]
    x = batch[A.data.KEY_IMAGES]

    def oracle(sd0, dtype):
        s = O.clone_state(sd0, requires_grad=True, dtype=dtype)
        c, p_ = O.unet_forward(s, x.to(dtype), True)
        l = O.unet_loss(c, p_, labels[:, 0:1].to(dtype), labels[:, 1:2].to(dtype))
        return c, p_, l, O.grads_of(l, s)

    c32, p32, l32, g32 = oracle(sd, torch.float32)
    _, _, _, g64 = oracle(sd, torch.float64)
    assert rel_l2(dto.outputs.core, c32) < TOL_ACT and rel_l2(dto.outputs.penu, p32) < TOL_ACT
    assert abs(loss.item() - l32.item()) < 1e-5
    # sampled fp32 noise floor (see _fp32_noise_floor): the U-Net backward crosses two max-pools and ten LeakyReLU kinks
    # and ends in BatchNorm backward cancellations; the reference's own fp32 CPU gradient of block1's first BN weight is
    # anywhere between 1.6e-5 and 1.2e-3 from fp64 when the parameters move by one ulp, so one draw is not a bound.
    names = set(g64)
    floor = {n: rel_l2(g32[n], g64[n]) for n in names}
    gen = torch.Generator().manual_seed(7)
    for _ in range(3):
        sdk = O.ulp_perturbed(sd, names, gen)
        g32k, g64k = oracle(sdk, torch.float32)[3], oracle(sdk, torch.float64)[3]
        for n in names:
            floor[n] = max(floor[n], rel_l2(g32k[n], g64k[n]))
    bad = []
    for n, p in unet.named_parameters():
        e_gpu = rel_l2(p.grad, g64[n])
        print("%-36s gpu/f64 %.2e cpu32-floor/f64 %.2e" % (n, e_gpu, floor[n]))
        if e_gpu > max(TOL_GRAD, 2 * floor[n]):
            bad.append("%s: gpu %g cpu32 floor %g" % (n, e_gpu, floor[n]))
    assert not bad, bad


def test_stacked_passes_equal_separate_passes():
    """Running core / penumbra / lesion (/ interpolation) as one stacked pass with per-group BatchNorm statistics is
    the same computation as the reference's separate calls: compare the two execution modes of the drop-in."""
    A = _api()
    import stroke_prediction_b200.common.model.Cae3D as M
    ch = [1, 4, 6, 8, 10, 12, 1]
    results = []
    for grouped in (True, False):
        M.GROUP_PASSES = grouped
        try:
            torch.manual_seed(41)
            cae = A.Cae3D(A.Enc3D(56, 28, ch, 5, 1.0), A.Dec3D(56, 28, ch, 5, 1.0)).cuda().train()
            opt = A.FusedAdam(cae.parameters(), lr=1e-3, weight_decay=1e-5)
            learner = A.CaeReconstructionLearner(None, None, cae, opt, None, 1, None, "/tmp/x", A.BatchDiceLoss([1.0]))
            batch = A.data.synthetic_cae_batch(2, size=(28, 56, 56), seed=5)
            dto = learner.inference_step(batch)
            loss = learner.loss_step(dto, 60)
            opt.zero_grad()
            loss.backward()
            results.append((loss.item(), {n: p.grad.clone() for n, p in cae.named_parameters()},
                            {k: v.clone() for k, v in cae.state_dict().items() if "running" in k or "num_batches" in k},
                            dto.reconstructions.gtruth.interpolation.detach().clone()))
            opt.detach_grad_sink()
        finally:
            M.GROUP_PASSES = True
    (l1, g1, b1, r1), (l2, g2, b2, r2) = results
    assert abs(l1 - l2) < 1e-6
    assert rel_l2(r1, r2) < 1e-6
    for k in b1:
        assert rel_max(b1[k].double(), b2[k].double()) < 1e-6, k
    for n in g1:
        assert rel_l2(g1[n], g2[n]) < 2e-4, n      # fp32 reduction partition differs (one 3B-sample wgrad vs three B-sample ones)


# ===================================================================================================== round 2 additions
def _unet_oracle(sd0, x, labels, dtype):
    s = O.clone_state(sd0, requires_grad=True, dtype=dtype)
    c, p_ = O.unet_forward(s, x.to(dtype), True)
    l = O.unet_loss(c, p_, labels[:, 0:1].to(dtype), labels[:, 1:2].to(dtype))
    return c, p_, l, O.grads_of(l, s)


# Gradient bar at the U-Net's own shape.  The exact-fp32 FFMA tiers meet the noise-floor rule of every other test (2 x the
# reference arithmetic's own fp32 floor).  The default tcgen05 tiers are held to 4 x the floor here: the tensor core's fp32
# accumulation truncates aligned addends towards zero, a systematic shrink of 7e-8 rms on every activation (an IEEE FFMA chain:
# 2e-10, profiles/r02_tc_bias_probe.log) that — unlike rounding noise — does not average out in the longest cancelling sums of
# the backward pass (2 x 64 x 164 x 164 voxels): measured 1.8e-4 .. 5.3e-4 on four of the 44 tensors where the FFMA tiers have
# 0.4e-4 .. 2.1e-4 and the CPU 0.4e-4 .. 2.4e-4 (profiles/r02_diag_unet_full.log).  Negated accumulation on alternating planes and
# round-to-nearest operand splits were tried and change nothing (the datapath is sign-symmetric); DESIGN.md 3.3.
@pytest.mark.parametrize("tier,factor", [("tcgen05", 4.0), ("ffma", 2.0)])
def test_unet_named_config_full_size_against_oracle(tier, factor):
    """BASELINE configs[0] at its OWN shape: channels 2 16 32 64 32 16 32 2, input 2 x 68 x 168 x 168 -> 28 x 128 x 128.
    B = 2: forward, loss and every parameter gradient (noise-floor rule against fp64); B = 4 (the benchmarked batch, whose
    tile / grid wrap differs): forward + loss."""
    A = _api()
    from stroke_prediction_b200 import ops
    ops.set_tc_terms(5 if tier == "tcgen05" else 0)
    try:
        _unet_full_size_check(A, factor)
    finally:
        ops.set_tc_terms(5)


_UNET_FULL = {}


def _unet_full_size_oracle(A):
    """Seeded model / batch and the CPU oracle's fp32 + fp64 runs (with one more one-ulp draw of the fp32 floor), computed once
    for both tiers."""
    if not _UNET_FULL:
        torch.manual_seed(35)
        unet = A.Unet3D([2, 16, 32, 64, 32, 16, 32, 2])
        sd = O.clone_state(unet.state_dict())
        batch = A.data.synthetic_unet_batch(2, out_size=(28, 128, 128), seed=4)
        labels, x = batch[A.data.KEY_LABELS], batch[A.data.KEY_IMAGES]
        c32, p32, l32, g32 = _unet_oracle(sd, x, labels, torch.float32)
        g64 = _unet_oracle(sd, x, labels, torch.float64)[3]
        names = set(g64)
        floor = {n: rel_l2(g32[n], g64[n]) for n in names}
        sdk = O.ulp_perturbed(sd, names, torch.Generator().manual_seed(7))      # one more draw of the fp32 floor
        g32k, g64k = _unet_oracle(sdk, x, labels, torch.float32)[3], _unet_oracle(sdk, x, labels, torch.float64)[3]
        for n in names:
            floor[n] = max(floor[n], rel_l2(g32k[n], g64k[n]))
        batch4 = A.data.synthetic_unet_batch(4, out_size=(28, 128, 128), seed=5)
        with torch.no_grad():
            c4, p4 = O.unet_forward(O.clone_state(sd), batch4[A.data.KEY_IMAGES], True)
            l4 = O.unet_loss(c4, p4, batch4[A.data.KEY_LABELS][:, 0:1], batch4[A.data.KEY_LABELS][:, 1:2])
        _UNET_FULL.update(sd=sd, batch=batch, c32=c32, p32=p32, l32=l32, g64=g64, floor=floor, batch4=batch4, c4=c4, p4=p4, l4=l4)
    return _UNET_FULL


def _unet_full_size_check(A, factor):
    R = _unet_full_size_oracle(A)
    sd, batch, floor, g64 = R["sd"], R["batch"], R["floor"], R["g64"]
    unet = A.Unet3D([2, 16, 32, 64, 32, 16, 32, 2])
    unet.load_state_dict(sd)
    unet = unet.cuda().train()
    opt = A.FusedAdam(unet.parameters(), lr=1e-3, weight_decay=1e-5, betas=(0.99, 0.999))
    learner = A.UnetSegmentationLearner(None, None, unet, opt, None, 1, A.BatchDiceLoss([1.0]))
    assert tuple(batch[A.data.KEY_IMAGES].shape) == (2, 2, 68, 168, 168)
    dto = learner.inference_step(batch)
    assert tuple(dto.outputs.core.shape) == (2, 1, 28, 128, 128)
    loss = learner.loss_step(dto, 0)
    opt.zero_grad()
    loss.backward()
    assert rel_l2(dto.outputs.core, R["c32"]) < TOL_ACT and rel_l2(dto.outputs.penu, R["p32"]) < TOL_ACT
    assert abs(_dice_binary(dto.outputs.penu.cpu(), R["p32"]) - 1.0) < TOL_DICE or float((R["p32"] > 0.5).sum()) == 0
    assert abs(loss.item() - R["l32"].item()) < 1e-5
    bad = []
    for n, p in unet.named_parameters():
        e_gpu = rel_l2(p.grad, g64[n])
        print("%-36s gpu/f64 %.2e cpu32-floor/f64 %.2e" % (n, e_gpu, floor[n]))
        if e_gpu > max(TOL_GRAD, factor * floor[n]):
            bad.append("%s: gpu %g cpu32 floor %g" % (n, e_gpu, floor[n]))
    assert not bad, bad
    # the benchmarked batch: B = 4 forward + loss, running statistics reset so both sides start from the same state
    del dto, loss
    unet.load_state_dict(sd)
    with torch.no_grad():
        dto4 = learner.inference_step(R["batch4"])
        loss4 = learner.loss_step(dto4, 0)
    assert rel_l2(dto4.outputs.core, R["c4"]) < TOL_ACT and rel_l2(dto4.outputs.penu, R["p4"]) < TOL_ACT
    assert abs(loss4.item() - R["l4"].item()) < 1e-5
    opt.detach_grad_sink()


def test_cae_named_config_benchmark_batch_forward_loss():
    """BASELINE configs[1] at the benchmarked batch 8 (stacked passes N = 24 / 32): latents, reconstructions, loss."""
    A = _api()
    ch = [1, 16, 24, 32, 100, 200, 1]
    torch.manual_seed(37)
    cae = A.Cae3D(A.Enc3D(128, 28, ch, 5, 1.0), A.Dec3D(128, 28, ch, 5, 1.0))
    sd = O.clone_state(cae.state_dict())
    cae = cae.cuda().train()
    learner = A.CaeReconstructionLearner(None, None, cae, A.FusedAdam(cae.parameters(), lr=1e-3), None, 1, None, "/tmp/x",
                                         A.BatchDiceLoss([1.0]))
    batch = A.data.synthetic_cae_batch(8, seed=6)
    with torch.no_grad():
        dto = learner.inference_step(batch)
        loss = learner.loss_step(dto, 60)
        labels = batch[A.data.KEY_LABELS]
        step = O.time_to_treatment(batch[A.data.KEY_GLOBAL])
        lat, rec = O.cae_forward(sd, ch, 1.0, True, labels[:, 0:1], labels[:, 1:2], labels[:, 2:3], step)
        oloss = O.cae_reconstruction_loss(lat, rec, labels[:, 0:1], labels[:, 1:2], labels[:, 2:3], 60)
    assert abs(loss.item() - oloss.item()) < 1e-5
    for k in ("core", "penu", "lesion", "interpolation"):
        assert rel_l2(getattr(dto.latents.gtruth, k), lat[k]) < TOL_ACT, k
        r = getattr(dto.reconstructions.gtruth, k)
        assert tuple(r.shape) == (8, 1, 28, 128, 128)
        assert rel_l2(r, rec[k]) < TOL_ACT, k


def test_enc3dctp_training_step_gradients():
    """SURVEY a5 backward: Cae3DCtp (3-channel encoder on mask + cropped CBV / TTD, virtual concat) through
    CaeReconstructionLearner.loss_step — every encoder / decoder gradient against the fp64 oracle of the same graph."""
    A = _api()
    from stroke_prediction_b200.common.dto import CaeDto as U
    from stroke_prediction_b200.common.model.Cae3D import Cae3DCtp
    ch = [3, 4, 6, 8, 10, 12, 1]
    torch.manual_seed(27)
    cae = Cae3DCtp(A.Enc3DCtp(56, 28, ch, 5, 1.0, [20, 20, 20]), A.Dec3D(56, 28, ch, 5, 1.0))
    sd0 = O.clone_state(cae.state_dict())
    cae = cae.cuda().train()
    b = A.data.synthetic_cae_batch(2, size=(28, 56, 56), seed=9)
    labels = b[A.data.KEY_LABELS]
    g = torch.Generator().manual_seed(3)
    cbv = torch.zeros(2, 1, 68, 96, 96)
    ttd = torch.zeros(2, 1, 68, 96, 96)
    cbv[:, :, 20:-20, 20:-20, 20:-20] = 12 * torch.rand(2, 1, 28, 56, 56, generator=g)
    ttd[:, :, 20:-20, 20:-20, 20:-20] = 40 * torch.rand(2, 1, 28, 56, 56, generator=g)
    step = O.time_to_treatment(b[A.data.KEY_GLOBAL])
    opt = A.FusedAdam(cae.parameters(), lr=1e-3, weight_decay=1e-5)
    learner = A.CaeReconstructionLearner(None, None, cae, opt, None, 1, None, "/tmp/x", A.BatchDiceLoss([1.0]))
    dto = U.init_dto(None, step.float().cuda(), None, None, cbv.cuda(), ttd.cuda(), labels[:, 0:1].cuda(),
                     labels[:, 1:2].cuda(), labels[:, 2:3].cuda())
    dto = cae(dto)
    loss = learner.loss_step(dto, 60)
    opt.zero_grad()
    loss.backward()

    def oracle(dtype, signs=None):
        sd = O.clone_state(sd0, requires_grad=True, dtype=dtype)
        crop = lambda t: t[:, :, 20:-20, 20:-20, 20:-20]
        cat = lambda m: torch.cat((m, crop(cbv), crop(ttd)), 1).to(dtype)
        lat = {k: O.encoder_pass(cat(labels[:, j:j + 1]), sd, ch, 1.0, True) for j, k in enumerate(("core", "penu", "lesion"))}
        lat["interpolation"] = O.interpolate(lat["core"], lat["penu"], step.to(dtype))
        rec = {k: O.decoder_pass(v, sd, ch, 1.0, True) for k, v in lat.items()}
        l = O.cae_reconstruction_loss(lat, rec, labels[:, 0:1].to(dtype), labels[:, 1:2].to(dtype), labels[:, 2:3].to(dtype), 60, signs)
        return lat, rec, l, O.grads_of(l, sd)

    lat32, rec32, l32, g32 = oracle(torch.float32)
    assert abs(loss.item() - l32.item()) < 1e-5
    for k in ("core", "penu", "lesion", "interpolation"):
        assert rel_l2(getattr(dto.latents.gtruth, k), lat32[k]) < TOL_ACT, k
    s_gpu = O.hinge_signs(dto.reconstructions.gtruth)
    _, rec64, _, g64 = oracle(torch.float64, s_gpu)
    _check_signs(s_gpu, rec64)
    g64_cpu = oracle(torch.float64, O.hinge_signs(rec32))[3]
    n = 0
    for name, p in cae.named_parameters():
        e_gpu, e_cpu = rel_l2(p.grad, g64[name]), rel_l2(g32[name], g64_cpu[name])
        assert e_gpu <= max(TOL_GRAD, 4 * e_cpu), "%s: gpu-vs-fp64 %g, cpu32-vs-fp64 %g" % (name, e_gpu, e_cpu)
        n += 1
    assert n > 0
    # CBV / TTD that require a gradient take the autograd (torch.cat) form of the concat: the input gradient must arrive
    cbv_g = cbv.cuda().requires_grad_(True)
    dto2 = U.init_dto(None, step.float().cuda(), None, None, cbv_g, ttd.cuda(), labels[:, 0:1].cuda(), labels[:, 1:2].cuda(),
                      labels[:, 2:3].cuda())
    dto2 = cae.enc(dto2)
    dto2.latents.gtruth.core.sum().backward()
    sd = O.clone_state(sd0, dtype=torch.float64)
    cbv64 = cbv.double().requires_grad_(True)
    crop = lambda t: t[:, :, 20:-20, 20:-20, 20:-20]
    O.encoder_pass(torch.cat((labels[:, 0:1].double(), crop(cbv64), crop(ttd.double())), 1), sd, ch, 1.0, True).sum().backward()
    assert rel_l2(cbv_g.grad, cbv64.grad) < 1e-4


def _sink_aliases(opt):
    sink = opt._sink
    base = sink.flat.data_ptr()
    return all(p.grad is not None and base <= p.grad.data_ptr() < base + 4 * sink.flat.numel() for p in sink.params)


def test_save_model_keeps_the_gradient_sink_attached(tmp_path):
    """train -> save_model -> train: the checkpoint must not move the live module (Module._apply re-binds param.grad
    storage and would cut the gradients loose from the flat buffer that the all-reduce and the fused Adam read)."""
    A = _api()
    ch = [1, 4, 6, 8, 10, 12, 1]
    torch.manual_seed(51)
    cae = A.Cae3D(A.Enc3D(56, 28, ch, 5, 1.0), A.Dec3D(56, 28, ch, 5, 1.0)).cuda().train()
    opt = torch.optim.Adam(cae.parameters(), lr=1e-3, weight_decay=1e-5)
    learner = A.CaeReconstructionLearner(None, None, cae, opt, None, 1, None, str(tmp_path / "run"), A.BatchDiceLoss([1.0]))
    learner.enable_data_parallel()            # world size 1: the sync is a no-op, the wiring is the real one
    fused = learner._optimizer
    batch = A.data.synthetic_cae_batch(2, size=(28, 56, 56), seed=5)
    learner.train_batch(batch, 30)
    assert _sink_aliases(fused)
    learner.save_model()
    learner.save_training()
    assert all(p.is_cuda for p in cae.parameters()) and _sink_aliases(fused)
    before = [p.detach().clone() for p in cae.parameters()]
    # gradients of the next step must land in the flat buffer (seen through a fresh forward/backward without the step)
    dto = learner.inference_step(batch)
    loss = learner.loss_step(dto, 30)
    fused.zero_grad()
    loss.backward()
    assert _sink_aliases(fused) and float(fused._sink.flat.abs().sum()) > 0
    gsum = sum(float(p.grad.double().abs().sum()) for p in fused._sink.params)
    assert abs(gsum - float(fused._sink.flat.double().abs().sum())) <= 1e-6 * gsum
    learner.train_batch(batch, 30)
    assert any((p.detach() - b).abs().max().item() > 0 for p, b in zip(cae.parameters(), before))
    # even a caller that DOES move the module (reference-style scripts) is healed on the next backward
    cae.cpu()
    cae.cuda()
    learner.train_batch(batch, 30)
    assert _sink_aliases(fused)
    # the checkpoint is a CPU whole-module pickle like the reference's (Learner.py:112-114)
    m = torch.load(str(tmp_path / "run") + "_cae1.model", weights_only=False)
    assert all(not p.is_cuda and p.grad is None for p in m.parameters())


def test_second_learner_over_the_same_model_owns_the_gradients():
    """The reference trains the CAE, then builds CaeStepLearner on the same modules: the newest optimizer's sink must
    receive the gradients (and the old one must stop claiming them)."""
    A = _api()
    ch = [1, 4, 6, 8, 10, 12, 1]
    torch.manual_seed(53)
    cae = A.Cae3D(A.Enc3D(56, 28, ch, 5, 1.0), A.Dec3D(56, 28, ch, 5, 1.0)).cuda().train()
    batch = A.data.synthetic_cae_batch(2, size=(28, 56, 56), seed=5)
    l1 = A.CaeReconstructionLearner(None, None, cae, torch.optim.Adam(cae.parameters(), lr=1e-3), None, 1, None, "/tmp/x",
                                    A.BatchDiceLoss([1.0]))
    l1.train_batch(batch, 30)
    old = l1._optimizer._sink
    l2 = A.CaeReconstructionLearner(None, None, cae, torch.optim.Adam(cae.parameters(), lr=1e-3), None, 1, None, "/tmp/x",
                                    A.BatchDiceLoss([1.0]))
    new = l2._optimizer._sink
    assert new is not old and not old.params
    dto = l2.inference_step(batch)
    loss = l2.loss_step(dto, 30)
    l2._optimizer.zero_grad()
    loss.backward()
    assert float(new.flat.abs().sum()) > 0 and float(old.flat.abs().sum()) == 0
    assert _sink_aliases(l2._optimizer)


def test_reference_format_checkpoint_round_trip(tmp_path):
    """a reference-format checkpoint — whole-module pickle whose classes are named by the REFERENCE's module
    path (`common.model.Cae3D.*`, Learner.py:112-114), a torch.optim.Adam `.optim` state_dict and the history file —
    resumes training in a new learner: same loss trajectory as the uninterrupted run, and the script's LR scheduler
    still drives the optimizer that steps (its objects survive `load_state_dict`)."""
    import contextlib
    import stroke_prediction_b200 as pkg
    import stroke_prediction_b200.common.model.Cae3D as M
    A = _api()
    pkg.install_reference_aliases()
    ch = [1, 4, 6, 8, 10, 12, 1]
    batches = [A.data.synthetic_cae_batch(2, size=(28, 56, 56), seed=60 + i) for i in range(4)]

    @contextlib.contextmanager
    def reference_class_paths():
        """While active, the model classes pickle as `common.model.Cae3D.<name>` exactly like a reference-written file."""
        classes = [c for c in vars(M).values() if isinstance(c, type) and c.__module__ == M.__name__]
        for c in classes:
            c.__module__ = "common.model.Cae3D"
        try:
            yield
        finally:
            for c in classes:
                c.__module__ = M.__name__

    def fresh():
        torch.manual_seed(55)
        cae = A.Cae3D(A.Enc3D(56, 28, ch, 5, 1.0), A.Dec3D(56, 28, ch, 5, 1.0)).cuda().train()
        opt = torch.optim.Adam(cae.parameters(), lr=1e-3, weight_decay=1e-5)
        sched = torch.optim.lr_scheduler.MultiStepLR(opt, [1, 3])
        return cae, opt, sched

    def learner(cae, opt, sched, prev, out):
        return A.CaeReconstructionLearner(None, None, cae, opt, sched, 4, prev, out, A.BatchDiceLoss([1.0]))

    # uninterrupted: 4 steps, one "epoch" per step so the scheduler acts (lr 1e-3, 1e-4, 1e-4, 1e-5)
    cae, opt, sched = fresh()
    ln = learner(cae, opt, sched, None, str(tmp_path / "a"))
    ref_losses, ref_lrs = [], []
    for e, b in enumerate(batches):
        ref_lrs.append(ln._optimizer.param_groups[0]["lr"])
        ref_losses.append(ln.train_batch(b, 30).loss)
        ln.adapt_lr(e)
    assert ref_lrs[0] == 1e-3 and abs(ref_lrs[1] - 1e-4) < 1e-12 and abs(ref_lrs[3] - 1e-5) < 1e-12

    # interrupted after 2 steps, checkpointed in the reference's on-disk format
    cae, opt, sched = fresh()
    ln = learner(cae, opt, sched, None, str(tmp_path / "b"))
    for e, b in enumerate(batches[:2]):
        m = ln.train_batch(b, 30)
        assert abs(m.loss - ref_losses[e]) < 1e-6
        ln.adapt_lr(e)
        ln._metric_dtos['training'].append(m)
        ln._metric_dtos['validate'].append(m)
    with reference_class_paths():
        ln.save_model()
    ln.save_training()
    with open(str(tmp_path / "b") + "_cae1.model", "rb") as f:
        raw = f.read()
    assert b"common.model.Cae3D" in raw and b"stroke_prediction_b200" not in raw
    sd_opt = torch.load(str(tmp_path / "b") + "_cae1.optim", weights_only=False)
    assert set(sd_opt) == {"state", "param_groups"} and "exp_avg" in next(iter(sd_opt["state"].values()))

    # resumed in a new process-like setting: new model object, new Adam, new scheduler (as the scripts build them)
    cae2, opt2, sched2 = fresh()
    with torch.no_grad():
        for p in cae2.parameters():
            p.add_(1.0)                       # prove the weights really come from the checkpoint
    groups_before, state_before = opt2.param_groups, opt2.state
    ln2 = learner(cae2, opt2, sched2, str(tmp_path / "b"), str(tmp_path / "c"))
    assert ln2.get_start_epoch() == 2 and abs(ln2.get_start_min_loss() - min(ref_losses[:2])) < 1e-6
    assert ln2._model is cae2                 # adopted in place: the optimizer still owns the live parameters
    assert ln2._optimizer.param_groups is groups_before and ln2._optimizer.state is state_before
    assert opt2.param_groups is groups_before and sched2.optimizer is opt2
    assert abs(opt2.param_groups[0]["lr"] - 1e-4) < 1e-12         # the saved learning rate came back
    sched2.last_epoch = 2                     # the reference does not persist the scheduler; position it by hand
    for e, b in enumerate(batches[2:], start=2):
        assert abs(ln2._optimizer.param_groups[0]["lr"] - ref_lrs[e]) < 1e-12, e
        got = ln2.train_batch(b, 30).loss
        assert abs(got - ref_losses[e]) < 2e-5 * max(1.0, abs(ref_losses[e])), (e, got, ref_losses[e])
        ln2.adapt_lr(e)

    # Tester path: the reference-format pickle loads by path (Tester.py:17) and runs
    class T(A.Tester, A.CaeInference):
        def __init__(self, path):
            A.Tester.__init__(self, None, path, "/tmp/x")
            A.CaeInference.__init__(self, self._model, 10)

    t = T(str(tmp_path / "b") + "_cae1.model")
    t._model.cuda()
    _, dto = t.infer_batch(A.data.synthetic_cae_batch(1, size=(28, 56, 56), seed=3))
    assert tuple(dto.reconstructions.gtruth.core.shape) == (1, 1, 28, 56, 56)


def test_validate_batch_matches_oracle_eval_forward():
    """Learner.validate_batch (Learner.py:132-142): eval-mode forward + loss + metrics, no graph, no parameter change."""
    A = _api()
    ch = [1, 4, 6, 8, 10, 12, 1]
    torch.manual_seed(57)
    cae = A.Cae3D(A.Enc3D(56, 28, ch, 5, 1.0), A.Dec3D(56, 28, ch, 5, 1.0))
    sd = O.clone_state(cae.state_dict())
    cae = cae.cuda().eval()
    ln = A.CaeReconstructionLearner(None, None, cae, A.FusedAdam(cae.parameters(), lr=1e-3), None, 1, None, "/tmp/x", A.BatchDiceLoss([1.0]))
    batch = A.data.synthetic_cae_batch(2, size=(28, 56, 56), seed=8)
    before = {k: v.clone() for k, v in cae.state_dict().items()}
    m = ln.validate_batch(batch, 60)
    labels = batch[A.data.KEY_LABELS]
    step = O.time_to_treatment(batch[A.data.KEY_GLOBAL])
    with torch.no_grad():
        lat, rec = O.cae_forward(sd, ch, 1.0, False, labels[:, 0:1], labels[:, 1:2], labels[:, 2:3], step)
        oloss = O.cae_reconstruction_loss(lat, rec, labels[:, 0:1], labels[:, 1:2], labels[:, 2:3], 60)
    assert abs(m.loss - oloss.item()) < 1e-5
    want = O.binary_measures(rec["core"], labels[:, 0:1])
    assert abs(m.core.dc - want["dc"]) < TOL_DICE
    for k, v in cae.state_dict().items():
        assert torch.equal(v, before[k]), k


def test_packed_weight_cache_is_keyed_on_geometry():
    """Unchanged weights, changing volume size (validation / Tester / visualisation after training): each geometry selects
    its own tier and packed layout (FFMA pack below 4096 output voxels, tensor-core image above, GEMM layout for wide tiny
    planes), so a pack cached for one size must never be served to another."""
    import torch.nn as nn
    from stroke_prediction_b200 import engine
    torch.manual_seed(59)
    seq = nn.Sequential(nn.BatchNorm3d(16), nn.Conv3d(16, 16, 3, padding=(1, 0, 0)), nn.ELU(1.0),
                        nn.BatchNorm3d(16), nn.ConvTranspose3d(16, 16, 3), nn.ELU(1.0)).eval()
    for m in seq:
        if isinstance(m, nn.BatchNorm3d):
            m.running_mean.uniform_(-0.5, 0.5)
            m.running_var.uniform_(0.5, 2.0)
    plan = engine.SeqPlan(seq.cuda())
    for size in [(6, 10, 10), (20, 40, 40), (6, 10, 10), (8, 64, 64), (20, 40, 40)]:
        x = torch.randn(2, 16, *size)
        with torch.no_grad():
            y = engine.run_sequential(plan, x.cuda())
            want = seq.cpu()(x)
            seq.cuda()
        assert rel_l2(y, want) < TOL_ACT, size
    # gradients w.r.t. the input through both directions at alternating sizes as well
    for size in [(6, 10, 10), (20, 40, 40), (6, 10, 10)]:
        x = torch.randn(2, 16, *size)
        xg = x.cuda().requires_grad_(True)
        engine.run_sequential(plan, xg).square().sum().backward()
        xc = x.clone().requires_grad_(True)
        seq.cpu()(xc).square().sum().backward()
        seq.cuda()
        assert rel_l2(xg.grad, xc.grad) < 1e-4, size


def test_curve_tester_cached_sweep_equals_per_step_inference():
    """CaeReconstructionTesterCurve.infer_curve (latents cached, all time points decoded as one stacked pass, one
    D2H) against the reference's procedure — a full model run per time point (CaeReconstructionTesterCurve.py:18-42)."""
    from stroke_prediction_b200.tester.CaeReconstructionTesterCurve import CaeReconstructionTesterCurve
    A = _api()
    ch = [1, 4, 6, 8, 10, 12, 1]
    torch.manual_seed(61)
    cae = A.Cae3D(A.Enc3D(56, 28, ch, 5, 1.0), A.Dec3D(56, 28, ch, 5, 1.0))
    for m in cae.modules():                       # non-trivial running statistics (eval mode uses them)
        if isinstance(m, torch.nn.BatchNorm3d):
            m.running_mean.uniform_(-0.2, 0.2)
            m.running_var.uniform_(0.6, 1.5)
    sd = O.clone_state(cae.state_dict())
    t = CaeReconstructionTesterCurve(None, cae.cuda(), "/tmp/x", 10, ta_to_tr_fixed_hours=range(4))
    batch = A.data.synthetic_cae_batch(1, size=(28, 56, 56), seed=3)
    sched = t.curve_steps(batch)
    assert len(sched) == 1 + 4 + 9 + 11 and sched[0] == (None, '')
    steps = [s for s, _ in sched]
    fast = t.infer_curve(batch, steps)
    assert len(fast) == len(steps)
    labels = batch[A.data.KEY_LABELS]
    for (m_fast, d_fast), s in zip(fast, steps):
        m_ref, d_ref = t.infer_batch(batch, s)
        assert rel_max(d_fast.given_variables.time_to_treatment, d_ref.given_variables.time_to_treatment) == 0
        assert rel_l2(d_fast.reconstructions.gtruth.interpolation, d_ref.reconstructions.gtruth.interpolation) < 1e-6
        assert rel_l2(d_fast.latents.gtruth.interpolation, d_ref.latents.gtruth.interpolation) < 1e-6
        for part in ("lesion", "core", "penu"):
            a, b = getattr(m_fast, part), getattr(m_ref, part)
            for k in ("dc", "hd", "assd", "precision", "sensitivity", "specificity"):
                va, vb = getattr(a, k), getattr(b, k)
                assert va == vb or abs(va - vb) <= 1e-9 * abs(vb), (s, part, k, va, vb)
    # and one time point against the oracle (eval mode, batch 1)
    s = steps[3]
    step = O.time_to_treatment(batch[A.data.KEY_GLOBAL], 10.0, s)
    lat, rec = O.cae_forward(sd, ch, 1.0, False, labels[:, 0:1], labels[:, 1:2], labels[:, 2:3], step)
    assert rel_l2(fast[3][1].reconstructions.gtruth.interpolation, rec["interpolation"]) < TOL_ACT
    want = O.binary_measures(rec["interpolation"], labels[:, 2:3])
    assert abs(fast[3][0].lesion.dc - want["dc"]) < TOL_DICE
    import io, contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        batch[A.data.KEY_CASE_ID] = torch.tensor([7])
        t._dataloader = [batch]
        t.run_inference()
    assert buf.getvalue().count("Case Id=7") == len(steps) and "tr_to_penumbra=1.0" in buf.getvalue()


def test_sdm_baseline_matches_scipy_restatement():
    """test_sdm_resampling.py:15-52 with resample=False: signed distance maps + interpolation + sign segmentation."""
    from stroke_prediction_b200.common import sdm
    A = _api()
    b = A.data.synthetic_cae_batch(1, size=(28, 64, 64), seed=12)
    lab = b[A.data.KEY_LABELS]
    core, penu = lab[:, 0:1], lab[:, 1:2]
    tt = 0.37
    rc, ri, rp = sdm.sdm_interpolate(core.cuda(), penu.cuda(), tt)
    pc = O.signed_distance_map(core[0, 0].numpy(), 0.5, False, -1.0)
    pp = O.signed_distance_map(penu[0, 0].numpy(), 0.5, True, 1.0)
    want = pp * tt - pc * (1 - tt)
    assert rel_max(rc, pc) < 1e-6 and rel_max(rp, pp) < 1e-6 and rel_max(ri, want) < 1e-5
    seg = (ri > 0).float().cpu()
    assert abs(_dice_binary(seg, torch.from_numpy((want > 0).astype(np.float32))) - 1.0) < TOL_DICE
    # degenerate case: empty core -> artificial core at the penumbra's centre of mass (reference :24-29)
    rc2, _, _ = sdm.sdm_interpolate(torch.zeros_like(core).cuda(), penu.cuda(), tt)
    assert float(rc2.min()) == 0.0 and int((rc2 == 0).sum()) == 63      # 3 dilations of one voxel by the 3-D cross: |x|+|y|+|z| <= 3 -> 63 voxels


def test_cae_bf16_mode_against_oracle():
    """bf16 mode (BASELINE north_star "bf16 mode within a stated looser tolerance", configs[2]): tensor-core operands are one bf16
    term, accumulation / storage / BatchNorm statistics / loss sums / master weights stay fp32.  Stated tolerance, against the fp32
    oracle at the named shape: latents and reconstructions 2e-2 relative, loss 1e-2 relative, thresholded Dice within 1e-2; three
    training steps follow the oracle's loss trajectory within 2e-2."""
    from stroke_prediction_b200 import ops
    A = _api()
    ch = [1, 16, 24, 32, 100, 800, 1]
    torch.manual_seed(63)
    cae = A.Cae3D(A.Enc3D(128, 28, ch, 5, 1.0), A.Dec3D(128, 28, ch, 5, 1.0))
    sd = O.clone_state(cae.state_dict(), requires_grad=True)
    cae = cae.cuda().train()
    ops.set_tc_terms(1)
    try:
        opt = torch.optim.Adam(cae.parameters(), lr=1e-3, weight_decay=1e-5)
        learner = A.CaeReconstructionLearner(None, None, cae, opt, None, 10, None, "/tmp/x", A.BatchDiceLoss([1.0]))
        names = [k for k, v in sd.items() if v.requires_grad]
        m = {k: torch.zeros_like(sd[k]) for k in names}
        v = {k: torch.zeros_like(sd[k]) for k in names}
        for it in range(3):
            batch = A.data.synthetic_cae_batch(2, seed=70 + it)
            labels = batch[A.data.KEY_LABELS]
            step = O.time_to_treatment(batch[A.data.KEY_GLOBAL])
            if it == 0:
                with torch.no_grad():
                    dto = learner.inference_step(batch)
                    lat, rec = O.cae_forward(O.clone_state({k: t.detach() for k, t in sd.items()}), ch, 1.0, True, labels[:, 0:1], labels[:, 1:2],
                                             labels[:, 2:3], step)
                for k in ("core", "penu", "lesion", "interpolation"):
                    assert rel_l2(getattr(dto.latents.gtruth, k), lat[k]) < 2e-2, k
                    r = getattr(dto.reconstructions.gtruth, k)
                    assert rel_l2(r, rec[k]) < 2e-2, k
                    assert abs(_dice_binary(r.cpu(), rec[k]) - 1.0) < 1e-2 or float((rec[k] > 0.5).sum()) == 0
                cae.load_state_dict({k: t.detach() for k, t in sd.items()})      # undo the running-statistics update of the probe pass
            got = learner.train_batch(batch, 60).loss
            lat, rec = O.cae_forward(sd, ch, 1.0, True, labels[:, 0:1], labels[:, 1:2], labels[:, 2:3], step)
            loss = O.cae_reconstruction_loss(lat, rec, labels[:, 0:1], labels[:, 1:2], labels[:, 2:3], 60)
            grads = O.grads_of(loss, sd)
            with torch.no_grad():
                for k in names:
                    newp, m[k], v[k] = O.adam_step(sd[k], grads[k], m[k], v[k], it + 1)
                    sd[k].copy_(newp)
            assert abs(got - loss.item()) < (1e-2 if it == 0 else 2e-2) * max(1.0, abs(loss.item())), (it, got, loss.item())
    finally:
        ops.set_tc_terms(5)
