"""Diagnostic (GPU box): where does the data-gradient error enter the decoder backward chain?  For a hinge-only loss
(the ill-conditioned term) compare dL/d(conv output) of every decoder unit — CUDA path and fp32 CPU
oracle, both against the fp64 oracle.  usage: python tools/diag_chain.py [tiny|full] [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
import stroke_oracle as O  # noqa: E402
from stroke_prediction_b200 import engine, functions as Fn  # noqa: E402
from stroke_prediction_b200.common import data  # noqa: E402
from stroke_prediction_b200.common.metrics import BatchDiceLoss  # noqa: E402
from stroke_prediction_b200.common.model.Cae3D import Cae3D, Dec3D, Enc3D  # noqa: E402
from stroke_prediction_b200.learner.CaeReconstructionLearner import CaeReconstructionLearner  # noqa: E402
from stroke_prediction_b200.optim import FusedAdam  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "tiny"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ch, size = ([1, 16, 24, 32, 100, 200, 1], (28, 128, 128)) if mode == "full" else ([1, 4, 6, 8, 10, 12, 1], (28, 56, 56))
torch.manual_seed(31)
EPOCH = 60
if mode == "fixture":      # the committed reference fixture (random BN affine parameters, epoch 30)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import load, state_from, unpack_masks
    fx = load("cae_rec_tiny")
    ch = [int(c) for c in fx["channels"]]
    size = tuple(int(v) for v in fx["size"])
    EPOCH = int(fx["epoch"])
cae = Cae3D(Enc3D(size[1], size[0], ch, 5, 1.0), Dec3D(size[1], size[0], ch, 5, 1.0))
if mode == "fixture":
    cae.load_state_dict(state_from(fx, "sd0/"))
    batch = {data.KEY_IMAGES: torch.zeros(2, 2, 1, 1, 1), data.KEY_LABELS: unpack_masks(fx), data.KEY_GLOBAL: torch.from_numpy(fx["clinical"])}
else:
    batch = data.synthetic_cae_batch(B, size=size, seed=4)
if os.environ.get("PLAIN_BN") == "1":    # what-if: neutral BN affine parameters
    for m in cae.modules():
        if isinstance(m, torch.nn.BatchNorm3d):
            torch.nn.init.ones_(m.weight); torch.nn.init.zeros_(m.bias)
sd0 = {k: v.clone() for k, v in cae.state_dict().items()}
cae = cae.cuda().train()
opt = FusedAdam(cae.parameters(), lr=1e-3, weight_decay=1e-5)
learner = CaeReconstructionLearner(None, None, cae, opt, None, 1, None, "/tmp/x", BatchDiceLoss([1.0]))

engine.DEBUG_GZ = []
engine.DEBUG_ACTS = []
dto = learner.inference_step(batch)
rec = dto.reconstructions.gtruth
FULL = os.environ.get("TERMS", "hinge") == "all"
loss = learner.loss_step(dto, EPOCH) if FULL else Fn.hinge_mean(rec.penu, rec.interpolation) + Fn.hinge_mean(rec.penu, rec.core)
loss.backward()
signs_gpu = O.hinge_signs(rec)
torch.cuda.synchronize()
dec_plan = cae.dec._plan if hasattr(cae.dec, "_plan") else None
gpu, gpu_enc = {}, {}
for plan, i, gz in engine.DEBUG_GZ:
    nunits = len(plan.units)
    if nunits == 12:
        gpu[i] = gz.detach().cpu()
    elif nunits == 10:
        gpu_enc[i] = gz.detach().cpu()
print("captured decoder units:", sorted(gpu), "encoder units:", sorted(gpu_enc))


def enc_capture(x, sd, cap):
    for i, (cin, cout, stride, pad) in enumerate(O.enc_table(ch)):
        x = O._bn(x, sd, 'enc.encoder.%d' % (3 * i), True)
        x = F.conv3d(x, sd['enc.encoder.%d.weight' % (3 * i + 1)], sd['enc.encoder.%d.bias' % (3 * i + 1)], stride, pad)
        x.retain_grad()
        cap.setdefault(i, []).append(x)
        x = F.elu(x, 1.0)
        cap.setdefault(("act", i), []).append(x.detach())
    return x


def dec_capture(z, sd, cap):
    table = O.dec_table(ch)
    for i, (kind, cin, cout, k, stride, pad) in enumerate(table):
        z = O._bn(z, sd, 'dec.decoder.%d' % (3 * i), True)
        w, b = sd['dec.decoder.%d.weight' % (3 * i + 1)], sd['dec.decoder.%d.bias' % (3 * i + 1)]
        z = F.conv_transpose3d(z, w, b, stride, pad) if kind == 'T' else F.conv3d(z, w, b, stride, pad)
        z.retain_grad()
        cap.setdefault(i, []).append(z)
        z = torch.sigmoid(z) if i == len(table) - 1 else F.elu(z, 1.0)
        cap.setdefault(("act", i), []).append(z.detach())
    return z


labels = batch[data.KEY_LABELS]
res = {}
for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
    sd = O.clone_state(sd0, requires_grad=True, dtype=dt)
    lab = labels.to(dt)
    step = O.time_to_treatment(batch[data.KEY_GLOBAL]).to(dt)
    ecap = {}
    e = lambda x: enc_capture(x, sd, ecap)
    lat = {'core': e(lab[:, 0:1]), 'penu': e(lab[:, 1:2]), 'lesion': e(lab[:, 2:3])}
    lat['interpolation'] = O.interpolate(lat['core'], lat['penu'], step)
    cap = {}
    r = {k: dec_capture(lat[k], sd, cap) for k in ('core', 'penu', 'lesion', 'interpolation')}
    if FULL:
        l = O.cae_reconstruction_loss(lat, r, lab[:, 0:1], lab[:, 1:2], lab[:, 2:3], EPOCH, signs_gpu)
    else:
        l = O.hinge(r['penu'], r['interpolation'], signs_gpu[0]) + O.hinge(r['penu'], r['core'], signs_gpu[1])
    l.backward()
    res[name + "_act"] = {k[1]: torch.cat(v, 0) for k, v in cap.items() if isinstance(k, tuple)}
    res[name + "_eact"] = {k[1]: torch.cat(v, 0) for k, v in ecap.items() if isinstance(k, tuple)}
    ecap = {k: v for k, v in ecap.items() if not isinstance(k, tuple)}
    cap = {k: v for k, v in cap.items() if not isinstance(k, tuple)}
    cat = lambda caps: {i: torch.cat([(t.grad if t.grad is not None else torch.zeros_like(t)) for t in ts], 0) for i, ts in caps.items()}
    res[name] = cat(cap)
    res[name + "_enc"] = cat(ecap)
    res[name + "_w"] = {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}
print("loss gpu %.9f" % loss.item())
print("%-6s %12s %12s %12s" % ("unit", "gpu/f64", "cpu32/f64", "|g64|"))
for i in sorted(gpu, reverse=True):
    g64, g32 = res["f64"][i], res["f32"][i]
    g = gpu[i]
    if g.shape != g64.shape:
        print(i, "shape mismatch", tuple(g.shape), tuple(g64.shape))
        continue
    # per pass (core, penu, lesion, interpolation)
    nb = g64.shape[0] // 4
    per = " ".join("%.1e/%.1e" % (O.rel_l2(g[j * nb:(j + 1) * nb], g64[j * nb:(j + 1) * nb]),
                                  O.rel_l2(g32[j * nb:(j + 1) * nb], g64[j * nb:(j + 1) * nb])) for j in (0, 1, 3))
    print("dec.%-3d %12.3e %12.3e %12.3e   %s" % (3 * i + 1, O.rel_l2(g, g64), O.rel_l2(g32, g64), g64.double().norm().item(), per))

print("encoder (stacked core, penu, lesion)")
for i in sorted(gpu_enc, reverse=True):
    g64, g32, g = res["f64_enc"][i], res["f32_enc"][i], gpu_enc[i]
    print("enc.%-3d %12.3e %12.3e %12.3e" % (3 * i + 1, O.rel_l2(g, g64), O.rel_l2(g32, g64), g64.double().norm().item()))
print("weight gradients (decisions aligned to the CUDA forward)")
for n, p in cae.named_parameters():
    if p.grad is None or n not in res["f64_w"] or not n.endswith("weight"):
        continue
    g64, g32 = res["f64_w"][n], res["f32_w"][n]
    print("%-28s %10.2e %10.2e   |g| %.3e" % (n, O.rel_l2(p.grad.cpu(), g64), O.rel_l2(g32, g64), g64.norm().item()))

print("decoder forward activations (post-activation outputs), per pass core/penu/interpolation: gpu/f64 | cpu32/f64")
for plan, acts in engine.DEBUG_ACTS:
    if len(plan.units) != 12:
        continue
    for i in range(12):
        a, a64, a32 = acts[i + 1].detach().cpu(), res["f64_act"][i], res["f32_act"][i]
        nb = a64.shape[0] // 4
        per = " ".join("%.1e/%.1e" % (O.rel_l2(a[j * nb:(j + 1) * nb], a64[j * nb:(j + 1) * nb]),
                                      O.rel_l2(a32[j * nb:(j + 1) * nb], a64[j * nb:(j + 1) * nb])) for j in (0, 1, 3))
        print("dec.%-3d %s" % (3 * i + 1, per))

print("encoder forward activations, per pass core/penu/lesion: gpu/f64 | cpu32/f64 ; then channel stats of the f64 core-pass output: min over channels of std, max |mean|/std")
for plan, acts in engine.DEBUG_ACTS:
    if len(plan.units) != 10:
        continue
    for i in range(10):
        a, a64, a32 = acts[i + 1].detach().cpu(), res["f64_eact"][i], res["f32_eact"][i]
        nb = a64.shape[0] // 3
        per = " ".join("%.1e/%.1e" % (O.rel_l2(a[j * nb:(j + 1) * nb], a64[j * nb:(j + 1) * nb]),
                                      O.rel_l2(a32[j * nb:(j + 1) * nb], a64[j * nb:(j + 1) * nb])) for j in (0, 1, 2))
        c = a64[:nb].transpose(0, 1).reshape(a64.shape[1], -1)
        sd_, mu_ = c.std(dim=1), c.mean(dim=1)
        print("enc.%-3d %s   min std %.2e  max |mean|/std %.1f" % (3 * i + 1, per, sd_.min().item(), (mu_.abs() / sd_).max().item()))
