"""Diagnostic (GPU box): per-parameter gradient error of the CUDA path and of the fp32 CPU oracle, both against an
fp64 CPU oracle run, for a CAE configuration.  usage: python tools/diag_grads.py [tiny|full] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402
import stroke_oracle as O  # noqa: E402
from stroke_prediction_b200.common import data  # noqa: E402
from stroke_prediction_b200.common.metrics import BatchDiceLoss  # noqa: E402
from stroke_prediction_b200.common.model.Cae3D import Cae3D, Dec3D, Enc3D  # noqa: E402
from stroke_prediction_b200.learner.CaeReconstructionLearner import CaeReconstructionLearner  # noqa: E402
from stroke_prediction_b200.optim import FusedAdam  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "tiny"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ch, size = ([1, 4, 6, 8, 10, 12, 1], (28, 56, 56)) if mode == "tiny" else ([1, 16, 24, 32, 100, 200, 1], (28, 128, 128))
torch.manual_seed(31)
cae = Cae3D(Enc3D(size[1], size[0], ch, 5, 1.0), Dec3D(size[1], size[0], ch, 5, 1.0))
sd0 = {k: v.clone() for k, v in cae.state_dict().items()}
cae = cae.cuda().train()
batch = data.synthetic_cae_batch(B, size=size, seed=4)
opt = FusedAdam(cae.parameters(), lr=1e-3, weight_decay=1e-5)
learner = CaeReconstructionLearner(None, None, cae, opt, None, 1, None, "/tmp/x", BatchDiceLoss([1.0]))
TERMS = os.environ.get("TERMS", "all")
from stroke_prediction_b200 import functions as Fn  # noqa: E402


def pick_gpu(dto):
    rec, giv, lat = dto.reconstructions.gtruth, dto.given_variables.gtruth, dto.latents.gtruth
    if TERMS == "dice":
        return Fn.dice_term(rec.core, giv.core) + Fn.dice_term(rec.penu, giv.penu) + Fn.dice_term(rec.lesion, giv.lesion)
    if TERMS == "hinge":
        return Fn.hinge_mean(rec.penu, rec.interpolation) + Fn.hinge_mean(rec.penu, rec.core)
    if TERMS == "l1":
        return Fn.l1_mean(lat.interpolation, lat.lesion)
    if TERMS == "hinge_a":
        return Fn.hinge_mean(rec.penu, rec.core.detach())
    if TERMS == "hinge_b":
        return Fn.hinge_mean(rec.penu.detach(), rec.core)
    if TERMS == "hinge_ab":
        return Fn.hinge_mean(rec.penu, rec.core)
    return learner.loss_step(dto, 60)


def pick_cpu(lat, rec, lab):
    if TERMS == "dice":
        return O.dice_loss(rec["core"], lab[:, 0:1]) + O.dice_loss(rec["penu"], lab[:, 1:2]) + O.dice_loss(rec["lesion"], lab[:, 2:3])
    if TERMS == "hinge":
        return O.hinge(rec["penu"], rec["interpolation"]) + O.hinge(rec["penu"], rec["core"])
    if TERMS == "l1":
        return O.l1(lat["interpolation"], lat["lesion"])
    if TERMS == "hinge_a":
        return O.hinge(rec["penu"], rec["core"].detach())
    if TERMS == "hinge_b":
        return O.hinge(rec["penu"].detach(), rec["core"])
    if TERMS == "hinge_ab":
        return O.hinge(rec["penu"], rec["core"])
    return O.cae_reconstruction_loss(lat, rec, lab[:, 0:1], lab[:, 1:2], lab[:, 2:3], 60)


dto = learner.inference_step(batch)
loss = pick_gpu(dto)
loss.backward()
labels = batch[data.KEY_LABELS]
res = {}
for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
    sd = O.clone_state(sd0, requires_grad=True, dtype=dt)
    lab = labels.to(dt)
    step = O.time_to_treatment(batch[data.KEY_GLOBAL]).to(dt)
    lat, rec = O.cae_forward(sd, ch, 1.0, True, lab[:, 0:1], lab[:, 1:2], lab[:, 2:3], step)
    l = pick_cpu(lat, rec, lab)
    res[name] = (l.item(), O.grads_of(l, sd), lat, rec)
print("loss gpu %.9f f32 %.9f f64 %.9f" % (loss.item(), res["f32"][0], res["f64"][0]))
for k in ("core", "penu", "lesion", "interpolation"):
    print("act %-14s lat gpu/f64 %.2e cpu32/f64 %.2e | rec gpu/f64 %.2e cpu32/f64 %.2e" % (
        k, O.rel_l2(getattr(dto.latents.gtruth, k).cpu(), res["f64"][2][k]), O.rel_l2(res["f32"][2][k], res["f64"][2][k]),
        O.rel_l2(getattr(dto.reconstructions.gtruth, k).cpu(), res["f64"][3][k]), O.rel_l2(res["f32"][3][k], res["f64"][3][k])))
print("%-28s %10s %10s %10s" % ("param", "gpu/f64", "cpu32/f64", "gpu/cpu32"))
for n, p in cae.named_parameters():
    if p.grad is None or res["f64"][1][n] is None or not n.endswith("weight") or n.split(".")[-2] not in ("1", "4", "13", "16", "28", "31", "34"):
        continue
    g64, g32 = res["f64"][1][n], res["f32"][1][n]
    print("%-28s %10.2e %10.2e %10.2e   |g| %.3e" % (n, O.rel_l2(p.grad.cpu(), g64), O.rel_l2(g32, g64), O.rel_l2(p.grad.cpu(), g32),
                                                   g64.norm().item()))
