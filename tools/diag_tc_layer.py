"""Diagnostics: per-layer accuracy (forward, dgrad, wgrad) of the tcgen05 tiers vs the exact-fp32 FFMA tiers vs the CPU's fp32, all
against an fp64 CPU run, at the U-Net's / CAE's real layer shapes and with upstream gradients of realistic dynamic range."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import torch.nn as nn

from stroke_prediction_b200 import engine, ops
from util import rel_l2

modes = [int(m) for m in (sys.argv[1].split(",") if len(sys.argv) > 1 else "5,4,0".split(","))]
cases = [("unet block5 conv b", 2, 16, 16, (30, 130, 130), 0), ("unet block5 conv b (patch)", 2, 16, 16, (30, 66, 66), 0),
         ("cae dec.28", 4, 16, 16, (28, 126, 126), (1, 2, 2)), ("cae enc.10", 4, 24, 24, (14, 62, 62), (1, 0, 0)),
         ("unet block5 conv a", 2, 48, 16, (32, 132, 132), 0), ("unet block2 conv b", 2, 32, 32, (30, 80, 80), 0)]
if os.environ.get("DIAG_CENTER"):    # zero-mean input (what BatchNorm hands the convolution in the real networks)
    cases = [c[:1] + c[1:] for c in cases[:3]]
for name, N, ci, co, size, pad in cases:
    torch.manual_seed(1)
    conv = nn.Conv3d(ci, co, 3, padding=pad)
    x = torch.randn(N, ci, *size)
    x = torch.where(x > 0, x, 0.01 * x) * 1.3 + 0.2          # post-LeakyReLU-like input with a mean
    if os.environ.get("DIAG_CENTER"):
        x = x - x.mean(dim=(0, 2, 3, 4), keepdim=True)
    y64 = conv.double()(x.double())
    for gkind in os.environ.get("DIAG_GRADS", "randn,sparse,sparse0,tail").split(","):
        g = torch.randn(y64.shape)
        if gkind == "sparse":        # a few huge values + small ones with a common offset
            g = g * (torch.rand(y64.shape) < 0.02).float() * 50 + 1e-3 * torch.randn(y64.shape) + 3e-3
        elif gkind == "sparse0":     # the same without the offset
            g = g * (torch.rand(y64.shape) < 0.02).float() * 50 + 1e-3 * torch.randn(y64.shape)
        elif gkind == "tail":        # log-normal magnitudes (4 decades), random signs
            g = g.sign() * torch.exp(2.3 * torch.randn(y64.shape))
        res = {}
        for label, dtype in (("cpu64", torch.float64), ("cpu32", torch.float32)):
            c = conv.to(dtype)
            c.zero_grad()
            xi = x.detach().clone().to(dtype).requires_grad_(True)
            y = c(xi)
            y.backward(g.to(dtype))
            res[label] = (y.detach(), xi.grad, c.weight.grad.clone())
        conv.float()
        for mode in modes:
            ops.set_tc_terms(mode)
            seq = nn.Sequential(nn.Conv3d(ci, co, 3, padding=pad)).cuda()
            seq[0].load_state_dict(conv.state_dict())
            plan = engine.SeqPlan(seq)
            xi = x.detach().clone().cuda().requires_grad_(True)
            y = engine.run_sequential(plan, xi)
            y.backward(g.cuda())
            res["gpu tc=%d" % mode] = (y.detach(), xi.grad, seq[0].weight.grad.clone())
        ref = res["cpu64"]
        print("%-28s grad=%-6s" % (name, gkind), " | ".join(
            "%s: y %.1e dx %.1e dw %.1e" % (k, rel_l2(v[0], ref[0]), rel_l2(v[1], ref[1]), rel_l2(v[2], ref[2]))
            for k, v in res.items() if k != "cpu64"))
ops.set_tc_terms(5)
