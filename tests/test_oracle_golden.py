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


def test_oracle_surface_measures_known_answers():
    """MedPy 0.3.0 hd / assd restatement (oracle.surface_measures) on hand-computed cases (metrics.py:31-47)."""
    import numpy as np
    import stroke_oracle as O
    # two single voxels in a 3-D array: both are their own border; distance sqrt(1 + 4 + 9)
    a = np.zeros((5, 6, 7), np.float32)
    b = np.zeros((5, 6, 7), np.float32)
    a[1, 1, 1] = 1
    b[2, 3, 4] = 1
    m = O.surface_measures(a, b)
    assert m["hd"] == 14 ** 0.5 and m["assd"] == 14 ** 0.5
    # a 3x3x3 cube against its centre voxel: the cube's border is its 26 shell voxels (connectivity-1 erosion leaves the centre),
    # their distances to the centre are 6 x 1, 12 x sqrt(2), 8 x sqrt(3); the centre voxel is 1 away from the shell
    a[:] = 0
    b[:] = 0
    a[1:4, 1:4, 1:4] = 1
    b[2, 2, 2] = 1
    m = O.surface_measures(a, b)
    want_rt = (6 * 1 + 12 * 2 ** 0.5 + 8 * 3 ** 0.5) / 26
    assert abs(m["asd_rt"] - want_rt) < 1e-12 and m["asd_tr"] == 1.0 and m["hd"] == 3 ** 0.5
    assert abs(m["assd"] - 0.5 * (want_rt + 1.0)) < 1e-12
    # the reference hands in B x 1 x D x H x W: the extent-1 channel axis empties the erosion -> every voxel is border, and the
    # batch axis is a lattice axis (a voxel in sample 0 and the same position in sample 1 are 1 apart)
    a5 = np.zeros((2, 1, 5, 6, 7), np.float32)
    b5 = np.zeros((2, 1, 5, 6, 7), np.float32)
    a5[0, 0, 1:4, 1:4, 1:4] = 1
    b5[1, 0, 2, 2, 2] = 1
    m = O.surface_measures(a5, b5)
    assert m["n_r"] == 27 and m["n_t"] == 1
    assert m["hd"] == (1 + 3) ** 0.5                      # corner voxel: sqrt(3) in-plane, 1 along the batch axis
    assert O.surface_measures(np.zeros((2, 1, 3, 3, 3)), b5[:, :, :3, :3, :3])["hd"] == float("inf")
    # signed distance map: 1-D-like check inside a volume
    v = np.zeros((1, 1, 9), np.float32)
    v[0, 0, 3:6] = 1
    s = O.signed_distance_map(v, 0.5, True, 1.0)
    assert list(s[0, 0]) == [-3, -2, -1, 1, 2, 1, -1, -2, -3]


def test_oracle_transforms_against_reference_fixture():
    """oracle.elastic_transform / resample_plane_xy / pad_images / to_tensor reproduce the reference's own transform classes
    (tests/golden/transforms_tiny.npz, common/data.py:280-380) bit for bit."""
    import numpy as np
    import stroke_oracle as O
    from util import load
    fx = load("transforms_tiny")
    labels, images = fx["labels"], fx["images"]
    rs = np.random.RandomState(int(fx["elastic_seed"]))
    for c in range(labels.shape[3]):
        noise = [rs.rand(*labels.shape[:3]) for _ in range(3)]
        assert np.array_equal(O.elastic_transform(labels[:, :, :, c], noise, 100, 4), fx["elastic_labels"][:, :, :, c])
    for c in range(images.shape[3]):
        noise = [rs.rand(*images.shape[:3]) for _ in range(3)]
        assert np.array_equal(O.elastic_transform(images[:, :, :, c], noise, 100, 4), fx["elastic_images"][:, :, :, c])
    for sf, tag in ((0.5, "0p5"), (0.75, "0p75")):
        for order, mode in ((0, "nearest"), (1, "bilinear")):
            assert np.array_equal(O.resample_plane_xy(images, sf, order), fx["zoom_%s_%s_images" % (tag, mode)])
    assert np.array_equal(O.pad_images(images, 3, 2, 1, 0.5), fx["pad_images"])
    assert np.array_equal(O.to_tensor(labels), fx["to_tensor_labels"])
