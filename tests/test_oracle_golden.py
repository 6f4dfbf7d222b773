"""CPU: the oracle restatement (oracle/stroke_oracle.py) against the fixtures produced by the reference's own modules
(oracle/make_golden.py).  Forward tensors were bit-identical when generated; across machines (different oneDNN ISA /
thread counts) a small tolerance is allowed."""
import numpy as np
import torch

import stroke_oracle as O
from util import load, rel_l2, rel_max, state_from, unpack_masks

FWD = 2e-5
GRAD = 1e-3   # accumulation-order noise of fp32 autograd, measured <= 1.3e-4 at generation time


def _cae_inputs(fx):
    labels = unpack_masks(fx)
    clinical = torch.from_numpy(fx["clinical"])
    return labels, clinical, [int(c) for c in fx["channels"]], float(fx["alpha"])


def test_cae_reconstruction_against_reference_fixture():
    fx = load("cae_rec_tiny")
    labels, clinical, ch, alpha = _cae_inputs(fx)
    sd = O.clone_state(state_from(fx, "sd0/"), requires_grad=True)
    core, penu, lesion = labels[:, 0:1], labels[:, 1:2], labels[:, 2:3]
    step = O.time_to_treatment(clinical)
    assert rel_max(step, fx["step"]) < 1e-6
    lat, rec = O.cae_forward(sd, ch, alpha, True, core, penu, lesion, step)
    loss = O.cae_reconstruction_loss(lat, rec, core, penu, lesion, int(fx["epoch"]))
    assert abs(loss.item() - float(fx["loss"])) < 1e-6
    for k in ("core", "penu", "lesion", "interpolation"):
        assert rel_l2(lat[k], fx["lat/" + k]) < FWD
        assert rel_l2(rec[k].detach().reshape(-1)[::13], fx["rec_sample/" + k]) < FWD
        m = fx["rec_moments/" + k]
        assert abs(rec[k].double().sum().item() - m[0]) <= 1e-5 * abs(m[0]) + 1e-3
    grads = O.grads_of(loss, sd)
    for k, v in fx.items():
        if k.startswith("grad/"):
            assert rel_max(grads[k[5:]], v) < GRAD, k
    # running statistics after the 3 encoder / 4 decoder calls
    for k, v in fx.items():
        if k.startswith("sd1/") and "running_" in k:
            assert rel_max(sd[k[4:]], v) < 1e-5, k
        if k.startswith("sd1/") and "num_batches" in k:
            assert int(sd[k[4:]]) == int(v)


def test_cae_step_learner_against_reference_fixture():
    fx = load("cae_step_tiny")
    labels, clinical, ch, alpha = _cae_inputs(fx)
    sd = O.clone_state(state_from(fx, "sd0/"))
    names = [k[5:] for k in fx if k.startswith("grad/")]
    assert sorted(names) == sorted(["enc.reduce.0.weight", "enc.reduce.0.bias", "enc.reduce.2.weight",
                                    "enc.reduce.2.bias", "enc.step.weight", "enc.step.bias"])
    for n in names:
        sd[n].requires_grad_(True)
    step = O.step_from_globals(clinical.float(), sd, alpha)
    lat, rec = O.cae_forward(sd, ch, alpha, True, labels[:, 0:1], labels[:, 1:2], labels[:, 2:3], step)
    loss = O.cae_step_loss(rec, labels[:, 2:3])
    assert abs(loss.item() - float(fx["loss"])) < 1e-6
    grads = O.grads_of(loss, sd)
    for n in names:
        assert rel_max(grads[n], fx["grad/" + n]) < GRAD, n


def test_cae_prediction_against_reference_fixture():
    fx = load("cae_pred_tiny")
    labels, clinical, ch, alpha = _cae_inputs(fx)
    soft = torch.from_numpy(fx["soft"].astype(np.float32))
    sd_cae = O.clone_state(state_from(fx, "cae0/"))
    sd_enc = O.clone_state({"enc." + k: v for k, v in state_from(fx, "enc0/").items()}, requires_grad=True)
    step = O.time_to_treatment(clinical)
    enc = lambda x: O.encoder_pass(x, sd_enc, ch, alpha, True, "enc.encoder")
    lat_in = {"core": enc(soft[:, 0:1]), "penu": enc(soft[:, 1:2])}
    lat_in["interpolation"] = O.interpolate(lat_in["core"], lat_in["penu"], step)
    rec_in = {k: O.decoder_pass(lat_in[k], sd_cae, ch, alpha, True) for k in ("core", "penu", "interpolation")}
    lat_gt, _ = O.cae_forward(sd_cae, ch, alpha, True, labels[:, 0:1], labels[:, 1:2], labels[:, 2:3], step)
    loss = O.cae_prediction_loss(lat_in, rec_in, lat_gt, labels[:, 2:3])
    assert abs(loss.item() - float(fx["loss"])) < 1e-6
    grads = O.grads_of(loss, sd_enc)
    for k, v in fx.items():
        if k.startswith("grad/"):
            assert rel_max(grads["enc." + k[5:]], v) < GRAD, k


def test_unet_against_reference_fixture():
    fx = load("unet_tiny")
    labels = unpack_masks(fx)
    interior = torch.from_numpy(fx["images_interior"])
    B, _, D, H, W = interior.shape
    img = torch.zeros(B, 2, D + 40, H + 40, W + 40)
    img[:, :, 20:-20, 20:-20, 20:-20] = interior
    sd = O.clone_state(state_from(fx, "sd0/"), requires_grad=True)
    core, penu = O.unet_forward(sd, img, True)
    assert rel_l2(core, fx["core"]) < FWD and rel_l2(penu, fx["penu"]) < FWD
    loss = O.unet_loss(core, penu, labels[:, 0:1], labels[:, 1:2])
    assert abs(loss.item() - float(fx["loss"])) < 1e-6
    grads = O.grads_of(loss, sd)
    for k, v in fx.items():
        if k.startswith("grad/"):
            assert rel_max(grads[k[5:]], v) < 5e-3, k     # U-Net fp32 gradient noise floor
    ec, ep = O.unet_forward(O.clone_state(state_from(fx, "sd0/")), img, False)
    assert rel_l2(ec, fx["eval_core"]) < FWD and rel_l2(ep, fx["eval_penu"]) < FWD


def test_adam_restatement_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(257)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-5)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for t in range(1, 4):
        g = torch.randn(257)
        ref.grad = g.clone()
        opt.step()
        p, m, v = O.adam_step(p, g, m, v, t)
        assert rel_max(p, ref) < 1e-6


def test_oracle_binary_measures_definitions():
    """medpy.metric.binary dc / precision / sensitivity / specificity on a hand-computed case (metrics.py:31-47)."""
    import torch
    import stroke_oracle as O
    r = torch.tensor([0.9, 0.8, 0.2, 0.6, 0.1, 0.5])      # > 0.5 -> 1 1 0 1 0 0
    t = torch.tensor([1.0, 0.0, 1.0, 1.0, 0.0, 0.0])
    m = O.binary_measures(r, t)
    assert m["counts"] == (2, 1, 1, 2)
    assert m["dc"] == 2 * 2 / (3 + 3) and m["precision"] == 2 / 3 and m["sensitivity"] == 2 / 3 and m["specificity"] == 2 / 3
    z = torch.zeros(4)
    e = O.binary_measures(z, z)
    assert e["dc"] == 0.0 and e["precision"] == 0.0 and e["sensitivity"] == 0.0 and e["specificity"] == 1.0
