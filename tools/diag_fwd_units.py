"""Diagnostic (GPU box): per-unit FORWARD error of the encoder on the committed fixture (random BN affine parameters).
For every unit i and every pass (core / penu / lesion) the unit's OWN input activation (as produced by that
implementation) is pushed through an fp64 copy of the unit; the distance of the implementation's output from that is
the error the unit itself adds.  Printed for the CUDA path and for the fp32 CPU oracle.
usage: python tools/diag_fwd_units.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
import stroke_oracle as O  # noqa: E402
from stroke_prediction_b200 import engine  # noqa: E402
from stroke_prediction_b200.common import data  # noqa: E402
from stroke_prediction_b200.common.model.Cae3D import Cae3D, Dec3D, Enc3D  # noqa: E402
from util import load, state_from, unpack_masks  # noqa: E402

fx = load("cae_rec_tiny")
ch = [int(c) for c in fx["channels"]]
size = tuple(int(v) for v in fx["size"])
cae = Cae3D(Enc3D(size[1], size[0], ch, 5, 1.0), Dec3D(size[1], size[0], ch, 5, 1.0))
cae.load_state_dict(state_from(fx, "sd0/"))
sd = {k: v.clone() for k, v in cae.state_dict().items()}
labels = unpack_masks(fx)
B = labels.shape[0]


def unit(x, sd, i, dtype):
    cin, cout, stride, pad = O.enc_table(ch)[i]
    p = "enc.encoder.%d" % (3 * i)
    q = "enc.encoder.%d" % (3 * i + 1)
    y = F.batch_norm(x.to(dtype), None, None, sd[p + ".weight"].to(dtype), sd[p + ".bias"].to(dtype), True, 0.1, 1e-5)
    bn = y
    y = F.conv3d(y, sd[q + ".weight"].to(dtype), sd[q + ".bias"].to(dtype), stride, pad)
    return bn, y, F.elu(y, 1.0)


# GPU chain
cae = cae.cuda().train()
engine.DEBUG_ACTS = []
from stroke_prediction_b200.learner.CaeReconstructionLearner import CaeReconstructionLearner  # noqa: E402
from stroke_prediction_b200.common.metrics import BatchDiceLoss  # noqa: E402
from stroke_prediction_b200.optim import FusedAdam  # noqa: E402
learner = CaeReconstructionLearner(None, None, cae, FusedAdam(cae.parameters(), lr=1e-3), None, 1, None, "/tmp/x", BatchDiceLoss([1.0]))
batch = {data.KEY_IMAGES: torch.zeros(B, 2, 1, 1, 1), data.KEY_LABELS: labels, data.KEY_GLOBAL: torch.from_numpy(fx["clinical"])}
dto = learner.inference_step(batch)
torch.cuda.synchronize()
enc_acts = None
for plan, acts in engine.DEBUG_ACTS:
    if len(plan.units) == 10 and enc_acts is None:
        enc_acts = [a.detach().float().cpu().contiguous() for a in acts]
print("encoder acts captured:", [tuple(a.shape) for a in enc_acts])
G = enc_acts[0].shape[0] // B

rel = O.rel_l2
names = ["core", "penu", "lesion"]
for g in range(G):
    x32 = labels[:, g:g + 1].clone()
    x64c = labels[:, g:g + 1].double()
    print("pass %s" % names[g])
    print(" unit   gpu:unit-err   cpu:unit-err | gpu:chain-err cpu:chain-err | |mean|/std of unit input (max over ch)   bg-fraction")
    for i in range(10):
        xin_gpu = enc_acts[i][g * B:(g + 1) * B]
        out_gpu = enc_acts[i + 1][g * B:(g + 1) * B]
        _, _, ref_gpu = unit(xin_gpu, sd, i, torch.float64)
        _, _, out_cpu = unit(x32, sd, i, torch.float32)
        _, _, ref_cpu = unit(x32, sd, i, torch.float64)
        _, _, x64c = unit(x64c, sd, i, torch.float64)
        xm = x32.double()
        m = xm.mean(dim=(0, 2, 3, 4)).abs()
        s = xm.std(dim=(0, 2, 3, 4))
        print("  %2d    %.2e       %.2e    |  %.2e      %.2e     |   %.1f" % (
            i, rel(out_gpu, ref_gpu), rel(out_cpu, ref_cpu), rel(out_gpu, x64c), rel(out_cpu, x64c), float((m / s).max())))
        x32 = out_cpu


# ---------------------------------------------------------------------------------------------- decoder
def dunit(x, sd, i, dtype):
    kind, cin, cout, k, stride, pad = O.dec_table(ch)[i]
    p = "dec.decoder.%d" % (3 * i)
    q = "dec.decoder.%d" % (3 * i + 1)
    y = F.batch_norm(x.to(dtype), None, None, sd[p + ".weight"].to(dtype), sd[p + ".bias"].to(dtype), True, 0.1, 1e-5)
    w, b = sd[q + ".weight"].to(dtype), sd[q + ".bias"].to(dtype)
    y = F.conv_transpose3d(y, w, b, stride, pad) if kind == "T" else F.conv3d(y, w, b, stride, pad)
    return torch.sigmoid(y) if i == 11 else F.elu(y, 1.0)


dec_acts = None
for plan, acts in engine.DEBUG_ACTS:
    if len(plan.units) == 12 and dec_acts is None:
        dec_acts = [a.detach().float().cpu().contiguous() for a in acts]
print("decoder acts captured:", [tuple(a.shape) for a in dec_acts])
G = dec_acts[0].shape[0] // B
for g in range(G):
    x32 = dec_acts[0][g * B:(g + 1) * B].clone()      # the CPU chain starts from the CUDA latent: isolates the decoder
    x64c = x32.double()
    print("decoder pass %d" % g)
    print(" unit   gpu:unit-err   cpu:unit-err | gpu:chain-err cpu:chain-err | |mean|/std of unit input (max over ch), count per channel")
    for i in range(12):
        xin_gpu = dec_acts[i][g * B:(g + 1) * B]
        out_gpu = dec_acts[i + 1][g * B:(g + 1) * B]
        ref_gpu = dunit(xin_gpu, sd, i, torch.float64)
        out_cpu = dunit(x32, sd, i, torch.float32)
        ref_cpu = dunit(x32, sd, i, torch.float64)
        x64c = dunit(x64c, sd, i, torch.float64)
        xm = x32.double()
        m = xm.mean(dim=(0, 2, 3, 4)).abs()
        s = xm.std(dim=(0, 2, 3, 4))
        print("  %2d    %.2e       %.2e    |  %.2e      %.2e     |   %.1f   %d" % (
            i, rel(out_gpu, ref_gpu), rel(out_cpu, ref_cpu), rel(out_gpu, x64c), rel(out_cpu, x64c), float((m / s).max()),
            xm.numel() // xm.shape[1]))
        x32 = out_cpu
