"""Diagnostics: per-tensor gradient error of the U-Net at BASELINE configs[0]'s own shape (2 x 68 x 168 x 168, B = 2) against an fp64
oracle run, for the tcgen05 tiers (default) and the exact-fp32 FFMA tiers (SP tc terms 0), next to several draws of the reference
arithmetic's own fp32 noise floor (CPU fp32 vs fp64 with the parameters moved by one ulp)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch

import stroke_oracle as O
from stroke_prediction_b200 import ops
from stroke_prediction_b200.common import data
from stroke_prediction_b200.common.metrics import BatchDiceLoss
from stroke_prediction_b200.common.model.Unet3D import Unet3D
from stroke_prediction_b200.learner.UnetSegmentationLearner import UnetSegmentationLearner
from stroke_prediction_b200.optim import FusedAdam
from util import rel_l2

draws = int(sys.argv[1]) if len(sys.argv) > 1 else 4
out_size = (28, 128, 128) if (len(sys.argv) < 3 or sys.argv[2] == "full") else (28, 64, 64)
torch.manual_seed(35)
unet = Unet3D([2, 16, 32, 64, 32, 16, 32, 2])
sd = O.clone_state(unet.state_dict())
batch = data.synthetic_unet_batch(2, out_size=out_size, seed=4)
labels, x = batch[data.KEY_LABELS], batch[data.KEY_IMAGES]


def oracle(sd0, dtype):
    s = O.clone_state(sd0, requires_grad=True, dtype=dtype)
    c, p_ = O.unet_forward(s, x.to(dtype), True)
    l = O.unet_loss(c, p_, labels[:, 0:1].to(dtype), labels[:, 1:2].to(dtype))
    return O.grads_of(l, s)


g64 = oracle(sd, torch.float64)
names = list(g64)
floors = [{n: rel_l2(oracle(sd, torch.float32)[n], g64[n]) for n in names}]
gen = torch.Generator().manual_seed(7)
for _ in range(draws - 1):
    sdk = O.ulp_perturbed(sd, set(names), gen)
    a, b = oracle(sdk, torch.float32), oracle(sdk, torch.float64)
    floors.append({n: rel_l2(a[n], b[n]) for n in names})

res = {}
for mode in (5, 4, 0):
    ops.set_tc_terms(mode)
    torch.manual_seed(35)
    m = Unet3D([2, 16, 32, 64, 32, 16, 32, 2])
    m.load_state_dict(sd)
    m = m.cuda().train()
    opt = FusedAdam(m.parameters(), lr=1e-3)
    ln = UnetSegmentationLearner(None, None, m, opt, None, 1, BatchDiceLoss([1.0]))
    dto = ln.inference_step(batch)
    loss = ln.loss_step(dto, 0)
    opt.zero_grad()
    loss.backward()
    res[mode] = {n: rel_l2(p.grad, g64[n]) for n, p in m.named_parameters()}
    opt.detach_grad_sink()
ops.set_tc_terms(5)
print("%-34s %9s %9s %9s | floor draws" % ("tensor", "tc3", "tc2", "ffma"))
for n in names:
    fl = [f[n] for f in floors]
    flag = "  <-- tc3 > 2 x max floor" if res[5][n] > max(1e-4, 2 * max(fl)) else ""
    print("%-34s %.2e %.2e %.2e | %s%s" % (n, res[5][n], res[4][n], res[0][n], " ".join("%.1e" % v for v in fl), flag))
