"""Diagnostic (GPU box): where do hinge seeds sign(d) of the CUDA path / fp32 CPU disagree with fp64?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import stroke_oracle as O
from stroke_prediction_b200.common import data
from stroke_prediction_b200.common.metrics import BatchDiceLoss
from stroke_prediction_b200.common.model.Cae3D import Cae3D, Dec3D, Enc3D
from stroke_prediction_b200.learner.CaeReconstructionLearner import CaeReconstructionLearner
from stroke_prediction_b200.optim import FusedAdam
ch, size, B = [1, 16, 24, 32, 100, 200, 1], (28, 128, 128), 2
torch.manual_seed(31)
cae = Cae3D(Enc3D(size[1], size[0], ch, 5, 1.0), Dec3D(size[1], size[0], ch, 5, 1.0))
sd0 = {k: v.clone() for k, v in cae.state_dict().items()}
cae = cae.cuda().train()
batch = data.synthetic_cae_batch(B, size=size, seed=4)
learner = CaeReconstructionLearner(None, None, cae, FusedAdam(cae.parameters()), None, 1, None, "/tmp/x", BatchDiceLoss([1.0]))
with torch.no_grad():
    dto = learner.inference_step(batch)
recs = {}
for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
    sd = O.clone_state(sd0, dtype=dt)
    lab = batch[data.KEY_LABELS].to(dt)
    step = O.time_to_treatment(batch[data.KEY_GLOBAL]).to(dt)
    with torch.no_grad():
        recs[name] = O.cae_forward(sd, ch, 1.0, True, lab[:, 0:1], lab[:, 1:2], lab[:, 2:3], step)[1]
g = dto.reconstructions.gtruth
for other in ("interpolation", "core"):
    d_gpu = (g.penu - getattr(g, other)).cpu().double()
    d32 = (recs["f32"]["penu"] - recs["f32"][other]).double()
    d64 = recs["f64"]["penu"] - recs["f64"][other]
    n = d64.numel()
    print("d = penu - %s: |d64| quantiles" % other, [float(d64.abs().flatten().kthvalue(max(1, int(q * n)))[0]) for q in (1e-5, 1e-4, 1e-3, 1e-2, 0.1, 0.5)])
    print("   exact zeros: gpu %d cpu32 %d f64 %d" % ((d_gpu == 0).sum(), (d32 == 0).sum(), (d64 == 0).sum()))
    print("   sign mismatches vs f64: gpu %d cpu32 %d of %d" % ((torch.sign(d_gpu) != torch.sign(d64)).sum(), (torch.sign(d32) != torch.sign(d64)).sum(), n))
    print("   max |d_gpu - d64| %.2e  max |d32 - d64| %.2e   rms %.2e / %.2e" % ((d_gpu - d64).abs().max(), (d32 - d64).abs().max(), (d_gpu - d64).pow(2).mean().sqrt(), (d32 - d64).pow(2).mean().sqrt()))
    print("   rec range", float(recs["f64"]["penu"].min()), float(recs["f64"]["penu"].max()))
