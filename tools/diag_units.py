"""Diagnostic (GPU box): per-unit backward precision at the real CAE layer shapes.  For every [BN, conv, act] unit of
the encoder/decoder (channels 1 16 24 32 100 200 1, B=2) feed random x / upstream gradient and compare input-grad and
weight-grad of the CUDA path and of fp32 CPU torch against fp64 CPU torch."""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402
from stroke_prediction_b200 import engine  # noqa: E402
from stroke_prediction_b200.common.model.Cae3D import Dec3D, Enc3D  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return ((a - b).norm() / b.norm()).item()


ch = [1, 16, 24, 32, 100, 200, 1]
B = 2
torch.manual_seed(0)
for name, seq, shape in (("enc", Enc3D(128, 28, ch, 5, 1.0).encoder, (B, 1, 28, 128, 128)),
                         ("dec", Dec3D(128, 28, ch, 5, 1.0).decoder, (B, 200, 1, 10, 10))):
    mods = list(seq.children())
    x = torch.randn(shape)
    for i in range(0, len(mods), 3):
        unit = nn.Sequential(*mods[i:i + 3]).train()
        xin = torch.nn.functional.elu(x * 0.8) if i > 0 else (x > 0.5).float()
        u32, u64, ug = copy.deepcopy(unit), copy.deepcopy(unit).double(), copy.deepcopy(unit).cuda()
        res = {}
        for tag, m, xi in (("f32", u32, xin.clone()), ("f64", u64, xin.double())):
            xi.requires_grad_(True)
            y = m(xi)
            gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(5)).to(y.dtype)
            y.backward(gy)
            res[tag] = (y.detach(), xi.grad, m[1].weight.grad, m[0].weight.grad)
        xg = xin.clone().cuda().requires_grad_(True)
        yg = engine.run_sequential(engine.SeqPlan(ug), xg)
        yg.backward(torch.randn(yg.shape, generator=torch.Generator().manual_seed(5)).cuda())
        y64, gx64, gw64, gg64 = res["f64"]
        y32, gx32, gw32, gg32 = res["f32"]
        conv = mods[i + 1]
        print("%s.%-2d %-15s %3d->%3d k%d s%d out %-18s | y %.1e/%.1e  gx %.1e/%.1e  gw %.1e/%.1e  ggamma %.1e/%.1e" % (
            name, i + 1, type(conv).__name__, conv.in_channels, conv.out_channels, conv.kernel_size[0], conv.stride[0],
            tuple(y64.shape[2:]), rel(yg.detach(), y64), rel(y32, y64), rel(xg.grad, gx64), rel(gx32, gx64),
            rel(ug[1].weight.grad, gw64), rel(gw32, gw64), rel(ug[0].weight.grad, gg64), rel(gg32, gg64)))
        x = torch.randn(y64.shape)
