"""GPU parity, op level: every kernel family behind the C-ABI against the CPU oracle (plain torch ops on the host)
on identical seeded inputs.  fp32 bar: 1e-4 relative (BASELINE.json north_star); gradients are additionally judged
against an fp64 CPU run with `err_gpu <= max(1e-4, 2 * err_cpu32)` (noise-floor rule)."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from util import TOL_ACT, TOL_GRAD, rel_l2, rel_max

pytestmark = pytest.mark.gpu


def _mods():
    from stroke_prediction_b200 import engine, functions, ops
    return engine, functions, ops


def _randomize_bn(seq, gen):
    for m in seq.modules():
        if isinstance(m, nn.BatchNorm3d):
            with torch.no_grad():
                m.weight.copy_(0.5 + torch.rand(m.weight.shape, generator=gen))
                m.bias.copy_(0.2 * torch.randn(m.bias.shape, generator=gen))
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=gen))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=gen))


def _check_sequential(seq, x, training=True, G=1, check_input_grad=True, tol_act=TOL_ACT, tol_grad=TOL_GRAD):
    """Run `seq` (a plain torch nn.Sequential) on the CPU in fp32 and fp64 and through the fused engine on the GPU."""
    engine, _, ops = _mods()
    gen = torch.Generator().manual_seed(1234)
    _randomize_bn(seq, gen)
    seq.train(training)
    ref32 = copy.deepcopy(seq)
    ref64 = copy.deepcopy(seq).double()
    mine = copy.deepcopy(seq).cuda()

    B = x.shape[0] // G
    outs = {}
    for name, mod, xin in (("f32", ref32, x.clone()), ("f64", ref64, x.double())):
        xin.requires_grad_(check_input_grad)
        ys = [mod(xin[g * B:(g + 1) * B]) for g in range(G)]   # G separate calls: separate statistics, sequential running stats
        y = torch.cat(ys, 0)
        gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(99)).to(y.dtype)
        y.backward(gy)
        outs[name] = (y.detach(), xin.grad, {n: p.grad for n, p in mod.named_parameters()}, dict(mod.named_buffers()))
    y64, gx64, gp64, _ = outs["f64"]
    y32, gx32, gp32, buf32 = outs["f32"]

    xg = x.clone().cuda().requires_grad_(check_input_grad)
    plan = engine.SeqPlan(mine)
    y = engine.run_sequential(plan, xg, G)
    gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(99)).cuda()
    y.backward(gy)
    torch.cuda.synchronize()

    assert tuple(y.shape) == tuple(y32.shape)
    assert rel_l2(y, y32) < tol_act, "forward: %g" % rel_l2(y, y32)

    def judge(mine_t, t32, t64, what):
        e_gpu, e_cpu = rel_l2(mine_t, t64), rel_l2(t32, t64)
        assert e_gpu <= max(tol_grad, 2 * e_cpu), "%s: gpu-vs-fp64 %g, cpu32-vs-fp64 %g" % (what, e_gpu, e_cpu)

    if check_input_grad:
        judge(xg.grad, gx32, gx64, "input grad")
    for n, p in mine.named_parameters():
        assert p.grad is not None, n
        judge(p.grad, gp32[n], gp64[n], "grad " + n)
    for n, b in mine.named_buffers():
        if "running" in n:
            assert rel_max(b, buf32[n]) < 1e-5, n
        if "num_batches" in n:
            assert int(b) == int(buf32[n]), n


# ---------------------------------------------------------------------------------------------------- layout
def test_layout_roundtrip():
    _, _, ops = _mods()
    x = torch.randn(2, 3, 5, 7, 9).cuda()
    v = ops.as_vol(x)
    assert ops.is_ndhwc(v) and torch.equal(v, x)
    assert torch.equal(v.permute(0, 2, 3, 4, 1).contiguous(), x.permute(0, 2, 3, 4, 1).contiguous())
    back = ops.to_ncdhw(v)
    assert back.is_contiguous() and torch.equal(back, x)
    big = torch.randn(1, 2, 40, 70, 50).cuda()   # > 65535 tiles-of-32 rows would break a 2-D grid
    assert torch.equal(ops.to_ncdhw(ops.as_vol(big)), big)


# ---------------------------------------------------------------------------------------------------- conv units
CONV_CASES = [
    # (kind, cin, cout, k, stride, padding, act, in_size)      reference call site
    ("C", 1, 16, 3, 1, (1, 0, 0), "elu", (6, 12, 11)),         # Cae3D.py:41   first encoder conv, Cin = 1
    ("C", 16, 16, 3, 1, (1, 0, 0), "elu", (5, 10, 13)),        # Cae3D.py:44
    ("C", 16, 24, 3, 2, 1, "elu", (8, 12, 10)),                # Cae3D.py:48   stride 2 pad 1
    ("C", 24, 24, 3, 1, (1, 0, 0), "elu", (4, 9, 9)),          # Cae3D.py:52
    ("C", 24, 32, 3, 2, 1, "elu", (7, 9, 11)),                 # Cae3D.py:59   odd extents, stride 2
    ("C", 32, 100, 3, 2, 0, "elu", (7, 9, 9)),                 # Cae3D.py:70   stride 2 pad 0
    ("C", 100, 40, 3, 1, 0, "elu", (3, 5, 4)),                 # Cae3D.py:74   bottleneck (fc reduced to 40 for speed)
    ("T", 40, 100, 3, 1, 0, "elu", (1, 3, 2)),                 # Cae3D.py:178  convT k3 s1
    ("T", 100, 32, 3, 2, 0, "elu", (3, 4, 5)),                 # Cae3D.py:182  convT k3 s2
    ("C", 32, 24, 3, 1, (1, 2, 2), "elu", (4, 6, 7)),          # Cae3D.py:189  pad > (k-1)/2
    ("T", 24, 24, 2, 2, 0, "elu", (3, 5, 4)),                  # Cae3D.py:193  convT k2 s2
    ("C", 16, 16, 1, 1, 0, "elu", (3, 6, 5)),                  # Cae3D.py:215  1x1
    ("C", 16, 1, 1, 1, 0, "sigmoid", (3, 6, 5)),               # Cae3D.py:218  1x1 + sigmoid
    ("C", 2, 16, 3, 1, 0, "leaky", (7, 9, 10)),                # Unet3D.py:19  valid conv, Cin = 2
    ("C", 96, 32, 3, 1, 0, "leaky", (5, 6, 7)),                # Unet3D.py:19  block4 on the concat
    ("C", 3, 8, 3, 1, (1, 0, 0), "elu", (4, 8, 8)),            # Enc3DCtp first conv, Cin = 3
    ("C", 5, 5, 1, 1, 0, "elu", (1, 1, 1)),                    # Cae3D.py:126  step MLP on B x 5 x 1 x 1 x 1
]


def _act(name):
    return {"elu": nn.ELU(1.0, True), "leaky": nn.LeakyReLU(0.01, True), "sigmoid": nn.Sigmoid(), "none": None}[name]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "%s%d-%d_k%ds%d" % (c[0], c[1], c[2], c[3], c[4]))
@pytest.mark.parametrize("with_bn", [True, False])
def test_fused_unit_forward_backward(case, with_bn):
    kind, cin, cout, k, s, p, act, size = case
    torch.manual_seed(100 + CONV_CASES.index(case))
    conv = nn.ConvTranspose3d(cin, cout, k, stride=s, padding=p) if kind == "T" else nn.Conv3d(cin, cout, k, stride=s, padding=p)
    mods = ([nn.BatchNorm3d(cin)] if with_bn else []) + [conv] + ([_act(act)] if _act(act) is not None else [])
    x = torch.randn(3, cin, *size) * 2.0 + 0.5
    _check_sequential(nn.Sequential(*mods), x)


@pytest.mark.parametrize("kind", ["C", "T"])
def test_paper_width_bottleneck_at_batch_8(kind):
    """Cae3D.py:74 / :178 at fc = 800 and the stacked batch of a training step (8 samples): 800 rows x 800 columns of GEMM output —
    more elements than the capped element-wise grids of the GEMM tier hold threads (their kernels must grid-stride), split-K forward
    with several slices, 128-bit im2col / col2im."""
    torch.manual_seed(150)
    if kind == "C":
        conv, x = nn.Conv3d(100, 800, 3), torch.randn(8, 100, 3, 12, 12) * 1.5 + 0.3
        mods = [nn.BatchNorm3d(100), conv, nn.ELU(1.0, True)]
    else:
        conv, x = nn.ConvTranspose3d(800, 100, 3), torch.randn(8, 800, 1, 10, 10) * 1.5 + 0.3
        mods = [nn.BatchNorm3d(800), conv, nn.ELU(1.0, True)]
    _check_sequential(nn.Sequential(*mods), x, G=2)


TILED_CASES = [
    # big enough (>= 2048 output voxels, Wo >= 8) to take the shared-memory tiled 3x3x3 path, with ragged tile edges
    ("C", 16, 16, 3, 1, (1, 0, 0), "elu", (9, 20, 37)),        # Cae3D.py:44
    ("C", 1, 16, 3, 1, (1, 0, 0), "elu", (9, 18, 35)),         # Cae3D.py:41   Cin = 1 -> CK = 4 kernel
    ("C", 2, 16, 3, 1, 0, "leaky", (12, 19, 33)),              # Unet3D.py:19  Cin = 2
    ("C", 3, 8, 3, 1, (1, 0, 0), "elu", (8, 18, 20)),          # Enc3DCtp, Cout = 8 < 16
    ("C", 24, 24, 3, 1, (1, 0, 0), "elu", (10, 17, 19)),       # Cae3D.py:52   Cout = 24 -> two passes, second ragged
    ("C", 32, 24, 3, 1, (1, 2, 2), "elu", (7, 14, 15)),        # Cae3D.py:189  pad 2: transposed pass has pad 0
    ("C", 96, 32, 3, 1, 0, "leaky", (10, 18, 18)),             # Unet3D.py:19  block4 conv a
    ("C", 16, 100, 3, 1, 0, "elu", (10, 14, 20)),              # Cout = 100 (not a multiple of 16)
    ("T", 12, 20, 3, 1, 0, "elu", (8, 14, 16)),                # convT k3 s1 forward = flipped correlation, pad 2
]


@pytest.mark.parametrize("case", TILED_CASES, ids=lambda c: "%s%d-%d_p%s" % (c[0], c[1], c[2], str(c[5]).replace(" ", "")))
def test_tiled_corr_paths(case):
    kind, cin, cout, k, s, p, act, size = case
    torch.manual_seed(200 + TILED_CASES.index(case))
    conv = nn.ConvTranspose3d(cin, cout, k, stride=s, padding=p) if kind == "T" else nn.Conv3d(cin, cout, k, stride=s, padding=p)
    x = torch.randn(2, cin, *size) * 1.5 + 0.3
    _check_sequential(nn.Sequential(nn.BatchNorm3d(cin), conv, _act(act)), x)


STRIDED_TILED_CASES = [
    # stride-2 layers big enough for the parity-split tiled tier (sp_conv_tiled.cuh, S = 2) and the parity-class transposed
    # kernel (sp_conv_tiledT.cuh): forward, dgrad and wgrad of each, ragged tiles on every axis
    ("C", 16, 24, 3, 2, 1, "elu", (12, 40, 44)),               # Cae3D.py:48   16 -> 24, pad 1
    ("C", 24, 32, 3, 2, 1, "elu", (13, 37, 41)),               # Cae3D.py:59   odd extents, two I-side channel passes in dgrad
    ("C", 32, 40, 3, 2, 0, "elu", (13, 37, 41)),               # Cae3D.py:70   pad 0; Cout = 40 > 32: dgrad stays generic
    ("C", 6, 10, 3, 2, 1, "elu", (12, 36, 40)),                # channel counts that are not multiples of 4
    ("T", 24, 24, 2, 2, 0, "elu", (7, 20, 22)),                # Cae3D.py:193  convT k2 s2
    ("T", 16, 16, 2, 2, 0, "elu", (6, 19, 23)),                # Cae3D.py:204  odd extents
    ("T", 12, 10, 3, 2, 0, "elu", (5, 14, 18)),                # Cae3D.py:182  convT k3 s2 (small channel counts)
    ("T", 5, 7, 2, 2, 0, "leaky", (6, 17, 21)),                # k2 s2, ragged channels
]


@pytest.mark.parametrize("wgrad_generation", [2, 1])
@pytest.mark.parametrize("case", STRIDED_TILED_CASES, ids=lambda c: "%s%d-%d_k%dp%s" % (c[0], c[1], c[2], c[3], str(c[5]).replace(" ", "")))
def test_strided_tiled_paths(case, wgrad_generation):
    """wgrad_generation 2: the weight gradients of the 9..32-channel k3 s2 p1 / k2 s2 layers come from the stride-2 tcgen05 kernel
    (sp_wgrad_tc4s2.cuh); 1: from the FFMA tiers (tiled / k2s2 kernels), which stay the fallback for every other geometry."""
    _, _, ops = _mods()
    kind, cin, cout, k, s, p, act, size = case
    torch.manual_seed(400 + STRIDED_TILED_CASES.index(case))
    conv = nn.ConvTranspose3d(cin, cout, k, stride=s, padding=p) if kind == "T" else nn.Conv3d(cin, cout, k, stride=s, padding=p)
    x = torch.randn(2, cin, *size) * 1.5 + 0.3
    ops.set_wgrad_tc_options(wgrad_generation, 0)
    try:
        _check_sequential(nn.Sequential(nn.BatchNorm3d(cin), conv, _act(act)), x, G=2)
    finally:
        ops.set_wgrad_tc_options(2, 0)


TC_CASES = [
    # >= 8192 output voxels, 8..16 channels on both sides, Ho >= 16: the tcgen05 / TMEM tier (sp_conv_tc.cuh) serves the
    # forward and — through flipped taps — the dgrad; ragged tiles in w, h and d
    ("C", 16, 16, 3, 1, (1, 0, 0), "elu", (9, 40, 29)),        # Cae3D.py:44   encoder 16->16
    ("C", 16, 16, 3, 1, (1, 2, 2), "elu", (9, 30, 30)),        # Cae3D.py:208  decoder 16->16, pad 2 (dgrad has pad 0)
    ("C", 16, 16, 3, 1, 0, "leaky", (12, 34, 30)),             # Unet3D.py:22  valid conv (dgrad has pad 2)
    ("C", 12, 10, 3, 1, (1, 1, 1), "elu", (10, 33, 31)),       # channel counts that are not multiples of 8
    ("T", 16, 16, 3, 1, 0, "elu", (8, 30, 36)),                # convT k3 s1 forward = flipped correlation
    # 17..24 channels: output width 24 and / or two input-channel passes (16 + 8) of the pipelined kernel
    ("C", 24, 24, 3, 1, (1, 0, 0), "elu", (9, 40, 29)),        # Cae3D.py:52,55
    ("C", 24, 16, 3, 1, (1, 2, 2), "elu", (9, 30, 30)),        # Cae3D.py:200  two passes in, 16 out (dgrad: 16 in, 24 out)
    ("C", 20, 18, 3, 1, (1, 1, 1), "leaky", (10, 33, 31)),     # ragged second pass (4 channels), ragged 24-wide output
    ("T", 24, 24, 3, 1, 0, "elu", (8, 30, 36)),                # convT k3 s1 with 24 channels
    ("C", 48, 16, 3, 1, 0, "leaky", (10, 30, 34)),             # Unet3D.py:19 block5 on the concat: three input passes; dgrad: three output slices
    ("C", 32, 32, 3, 1, 0, "leaky", (10, 30, 34)),             # Unet3D.py:22 block2/4: two passes x two 16-wide output slices
    ("C", 96, 32, 3, 1, 0, "leaky", (10, 30, 26)),             # Unet3D.py:19 block4 on the concat: six passes; dgrad: 2 passes x 6 slices
    ("C", 40, 28, 3, 1, (1, 1, 1), "elu", (9, 28, 22)),        # ragged last pass (8 channels) and ragged last slice (12 channels)
]


@pytest.mark.parametrize("case", TC_CASES, ids=lambda c: "%s%d-%d_p%s" % (c[0], c[1], c[2], str(c[5]).replace(" ", "")))
def test_tensor_core_corr_paths(case):
    """Default tcgen05 tier (mode 5: three bf16 terms, kw-stacked N, leading products and corrections in separate accumulators):
    held to the same bars as the exact-fp32 FFMA tier, including the CPU noise-floor rule for gradients."""
    _, _, ops = _mods()
    kind, cin, cout, k, s, p, act, size = case
    torch.manual_seed(300 + TC_CASES.index(case))
    conv = nn.ConvTranspose3d(cin, cout, k, stride=s, padding=p) if kind == "T" else nn.Conv3d(cin, cout, k, stride=s, padding=p)
    x = torch.randn(2, cin, *size) * 1.5 + 0.3
    assert ops.get_tc_terms() == 5
    _check_sequential(nn.Sequential(nn.BatchNorm3d(cin), conv, _act(act)), x, G=2)


WGRAD_TC_CASES = [
    # the tcgen05 weight-gradient tier (sp_wgrad_tc.cuh): 3x3x3 stride 1, 9..16 channels on both sides, Do >= 8, Wo >= 24;
    # several 32-wide column tiles with a ragged last one, ragged 4-row tiles, zero padding on every axis
    ("C", 16, 16, 3, 1, (1, 2, 2), "elu", (8, 17, 67)),        # Cae3D.py:208  Wo = 69: three column tiles
    ("C", 16, 16, 3, 1, (1, 0, 0), "elu", (10, 23, 44)),       # Cae3D.py:44   Wo = 42, Ho = 21
    ("C", 10, 14, 3, 1, 1, "leaky", (9, 18, 40)),              # ragged channel halves on both sides
    # 17..24 channels on either side: (16 + 8)-channel slices of the second-generation kernel (sp_wgrad_tc24.cuh when switched off)
    ("C", 24, 24, 3, 1, (1, 0, 0), "elu", (9, 19, 44)),        # Cae3D.py:52,55
    ("C", 24, 16, 3, 1, (1, 2, 2), "elu", (8, 14, 35)),        # Cae3D.py:200  24 -> 16: the third O-side group is empty
    ("C", 12, 20, 3, 1, 1, "leaky", (10, 17, 33)),             # ragged groups on both sides
    ("C", 48, 16, 3, 1, 0, "leaky", (10, 18, 40)),             # Unet3D.py:19 block5: three 16-channel I-side slices, scatter-reduce
    ("C", 40, 12, 3, 1, (1, 1, 1), "elu", (8, 16, 30)),        # ragged last slice (8 channels)
    ("C", 16, 40, 3, 1, (1, 1, 1), "elu", (8, 16, 30)),        # three O-side slices, the last one ragged
    ("C", 32, 32, 3, 1, 0, "leaky", (10, 18, 40)),             # 2 x 2 slice pairs (second generation: up to twelve pairs)
    ("C", 96, 32, 3, 1, 0, "leaky", (10, 18, 40)),             # Unet3D.py:19 block4: 6 x 2 = 12 slice pairs
    ("C", 64, 64, 3, 1, 0, "leaky", (10, 18, 40)),             # Unet3D.py:22 block3: 16 pairs -> stays on the FFMA tier
    ("C", 32, 32, 3, 1, (1, 2, 2), "elu", (5, 25, 25)),        # Cae3D.py:186: shallow volume (Do = 5), one column tile, 4 pairs
    ("C", 2, 16, 3, 1, 0, "leaky", (10, 24, 44)),              # Unet3D.py:19 block1: two input channels (upper channel half = zeros)
    ("C", 3, 16, 3, 1, (1, 0, 0), "elu", (10, 24, 44)),        # Enc3DCtp: three input channels, floats per voxel not a multiple of 4
]


@pytest.mark.parametrize("case", WGRAD_TC_CASES, ids=lambda c: "%s%d-%d_p%s" % (c[0], c[1], c[2], str(c[5]).replace(" ", "")))
def test_tensor_core_wgrad_paths(case):
    """dW from tcgen05 MMAs with MN-major operands (voxels = contraction index), walked along the depth axis through a
    ring of input planes; G = 2 stacked statistics groups.  Same bars as the FFMA tier."""
    kind, cin, cout, k, s, p, act, size = case
    torch.manual_seed(500 + WGRAD_TC_CASES.index(case))
    conv = nn.Conv3d(cin, cout, k, stride=s, padding=p)
    x = torch.randn(4, cin, *size) * 1.5 + 0.3
    _check_sequential(nn.Sequential(nn.BatchNorm3d(cin), conv, _act(act)), x, G=2)


WGRAD_TC_S2_CASES = [
    # the stride-2 tcgen05 weight gradient (sp_wgrad_tc4s2.cuh): 3x3x3 stride 2 padding 1, 9..16 input channels; tile columns of
    # both input-column parities in one launch per 16-channel output slice
    ("C", 16, 24, 3, 2, 1, "elu", (16, 36, 70)),          # Cae3D.py:48   even extents, two column tiles per parity
    ("C", 16, 24, 3, 2, 1, "elu", (17, 23, 67)),          # odd extents on every axis: ragged last plane / row / column
    ("C", 12, 16, 3, 2, 1, "leaky", (16, 20, 66)),        # ragged channel half on the I-side, one output slice
    ("C", 16, 32, 3, 2, 1, "elu", (16, 20, 66)),          # two full output slices
    ("C", 24, 32, 3, 2, 1, "elu", (14, 30, 58)),          # Cae3D.py:59   two input slices (16 + 8) x two output slices
]


@pytest.mark.parametrize("max_ctas", [0, 3])
@pytest.mark.parametrize("case", WGRAD_TC_S2_CASES, ids=lambda c: "%s%d-%d_%s" % (c[0], c[1], c[2], "x".join(map(str, c[7]))))
def test_tensor_core_wgrad_stride2(case, max_ctas):
    """dW of the stride-2 layer from tcgen05 MMAs: input columns split by parity (tap slots along M), rows / planes read at stride
    two through the N-group base and a ring that advances two planes per step; also with three CTAs (many columns per CTA)."""
    _, _, ops = _mods()
    kind, cin, cout, k, s, p, act, size = case
    torch.manual_seed(560 + WGRAD_TC_S2_CASES.index(case))
    conv = nn.Conv3d(cin, cout, k, stride=s, padding=p)
    x = torch.randn(4, cin, *size) * 1.5 + 0.3
    ops.set_wgrad_tc_options(2, max_ctas)
    try:
        _check_sequential(nn.Sequential(nn.BatchNorm3d(cin), conv, _act(act)), x, G=2)
    finally:
        ops.set_wgrad_tc_options(2, 0)


@pytest.mark.parametrize("max_ctas", [3, 1])
@pytest.mark.parametrize("case", WGRAD_TC_CASES[:7] + [WGRAD_TC_CASES[9]],
                         ids=lambda c: "%s%d-%d_p%s" % (c[0], c[1], c[2], str(c[5]).replace(" ", "")))
def test_tensor_core_wgrad_several_columns_per_cta(case, max_ctas):
    """Second-generation kernel with the persistent grid capped at 3 / 1 CTAs: every CTA walks many tile columns, so the
    first two planes of the next column are staged during the last two steps of the current one (ring of seven planes),
    statistics groups change between columns of one CTA, and the A buffers / ring slots wrap many times."""
    _, _, ops = _mods()
    kind, cin, cout, k, s, p, act, size = case
    torch.manual_seed(520 + WGRAD_TC_CASES.index(case))
    conv = nn.Conv3d(cin, cout, k, stride=s, padding=p)
    x = torch.randn(4, cin, *size) * 1.5 + 0.3
    ops.set_wgrad_tc_options(2, max_ctas)
    try:
        _check_sequential(nn.Sequential(nn.BatchNorm3d(cin), conv, _act(act)), x, G=2)
    finally:
        ops.set_wgrad_tc_options(2, 0)


@pytest.mark.parametrize("case", WGRAD_TC_CASES[:7], ids=lambda c: "%s%d-%d_p%s" % (c[0], c[1], c[2], str(c[5]).replace(" ", "")))
def test_tensor_core_wgrad_generation1(case):
    """First-generation kernels (three exact bf16 terms: M 64 x N 48 MMAs for 16 channels, sp_wgrad_tc24.cuh for 17..24, the
    16-channel kernel on slice pairs for 48 -> 16), kept for A/B measurements."""
    _, _, ops = _mods()
    kind, cin, cout, k, s, p, act, size = case
    torch.manual_seed(540 + WGRAD_TC_CASES.index(case))
    conv = nn.Conv3d(cin, cout, k, stride=s, padding=p)
    x = torch.randn(4, cin, *size) * 1.5 + 0.3
    ops.set_wgrad_tc_options(1, 0)
    try:
        _check_sequential(nn.Sequential(nn.BatchNorm3d(cin), conv, _act(act)), x, G=2)
    finally:
        ops.set_wgrad_tc_options(2, 0)


POINTWISE_CASES = [
    # 1x1x1 mixes on enough rows for several CTAs / grid-stride passes with a ragged tail (sp_conv_pw.cuh)
    (16, 16, "elu", (7, 33, 35)),          # Cae3D.py:215  pw16_fwd<16>, pw_wgrad16<16>
    (16, 1, "sigmoid", (7, 33, 35)),       # Cae3D.py:218  pw16_fwd<1>, pw_wgrad16<1>; dgrad 1 -> 16
    (16, 32, "leaky", (5, 21, 30)),        # Unet3D.py:50  two 16-wide output passes
    (32, 2, "sigmoid", (5, 21, 30)),       # Unet3D.py:52  runtime-Cs kernels
]


@pytest.mark.parametrize("case", POINTWISE_CASES, ids=lambda c: "%d-%d" % (c[0], c[1]))
def test_pointwise_paths(case):
    cin, cout, act, size = case
    torch.manual_seed(600 + POINTWISE_CASES.index(case))
    x = torch.randn(3, cin, *size) * 1.5 + 0.3
    _check_sequential(nn.Sequential(nn.BatchNorm3d(cin), nn.Conv3d(cin, cout, 1), _act(act)), x)


@pytest.mark.parametrize("case", TC_CASES, ids=lambda c: "%s%d-%d_p%s" % (c[0], c[1], c[2], str(c[5]).replace(" ", "")))
def test_tensor_core_generation2_paths(case):
    """Generation-2 kernel (mode 4), kept for A/B measurements: same arithmetic and bars as the default."""
    _, _, ops = _mods()
    kind, cin, cout, k, s, p, act, size = case
    torch.manual_seed(300 + TC_CASES.index(case))
    conv = nn.ConvTranspose3d(cin, cout, k, stride=s, padding=p) if kind == "T" else nn.Conv3d(cin, cout, k, stride=s, padding=p)
    x = torch.randn(2, cin, *size) * 1.5 + 0.3
    ops.set_tc_terms(4)
    try:
        _check_sequential(nn.Sequential(nn.BatchNorm3d(cin), conv, _act(act)), x, G=2)
    finally:
        ops.set_tc_terms(5)


# bf16 mode (BASELINE north_star: "bf16 mode within a stated looser tolerance"; SURVEY §8c suggests 2e-2 relative on activations):
# tensor-core operands are ONE bf16 term (8 significand bits, round to nearest) with fp32 accumulation; everything else stays fp32.
# Per layer that is a relative error of ~2^-9 per product: 2.0e-3 .. 2.4e-3 rel-L2 measured on outputs, input gradients and (the
# split-arithmetic) weight gradients alike (profiles/r02_diag_tc_layer.log); the activation bound leaves a factor 4.  The gradient
# bound covers the BatchNorm backward behind the dgrad, which subtracts the mean and the projection on x-hat and so amplifies the
# relative error of what is left (measured up to 4.2e-2 on the 32- and 96-channel units of this list).
TOL_BF16_ACT = 1e-2
TOL_BF16_GRAD = 6e-2


@pytest.mark.parametrize("case", [TC_CASES[0], TC_CASES[1], TC_CASES[3]] + TC_CASES[-3:-1],
                         ids=lambda c: "%s%d-%d_p%s" % (c[0], c[1], c[2], str(c[5]).replace(" ", "")))
def test_bf16_mode_paths(case):
    _, _, ops = _mods()
    kind, cin, cout, k, s, p, act, size = case
    torch.manual_seed(300 + TC_CASES.index(case))
    conv = nn.ConvTranspose3d(cin, cout, k, stride=s, padding=p) if kind == "T" else nn.Conv3d(cin, cout, k, stride=s, padding=p)
    x = torch.randn(2, cin, *size) * 1.5 + 0.3
    ops.set_tc_terms(1)
    try:
        assert ops.get_tc_terms() == 1
        _check_sequential(nn.Sequential(nn.BatchNorm3d(cin), conv, _act(act)), x, G=2, tol_act=TOL_BF16_ACT, tol_grad=TOL_BF16_GRAD)
    finally:
        ops.set_tc_terms(5)


@pytest.mark.parametrize("terms", [3, 2, 0])
@pytest.mark.parametrize("case", TC_CASES[:2], ids=lambda c: "%s%d-%d_p%s" % (c[0], c[1], c[2], str(c[5]).replace(" ", "")))
def test_tensor_core_legacy_modes(case, terms):
    """First-generation kernels (one accumulator per output; the truncating fp32 accumulator makes them 5-20x noisier than
    IEEE FFMA chains, so gradients get 1e-4 / 3e-4 bounds instead of the CPU noise floor) and the tier switched off."""
    _, _, ops = _mods()
    kind, cin, cout, k, s, p, act, size = case
    torch.manual_seed(300 + TC_CASES.index(case))
    conv = nn.ConvTranspose3d(cin, cout, k, stride=s, padding=p) if kind == "T" else nn.Conv3d(cin, cout, k, stride=s, padding=p)
    x = torch.randn(2, cin, *size) * 1.5 + 0.3
    ops.set_tc_terms(terms)
    try:
        assert ops.get_tc_terms() == terms
        _check_sequential(nn.Sequential(nn.BatchNorm3d(cin), conv, _act(act)), x, G=2, tol_grad=3e-4 if terms == 2 else 1e-4)
    finally:
        ops.set_tc_terms(5)


def test_tiled_chain_grouped():
    seq = nn.Sequential(nn.BatchNorm3d(1), nn.Conv3d(1, 16, 3, padding=(1, 0, 0)), nn.ELU(1.0, True),
                        nn.BatchNorm3d(16), nn.Conv3d(16, 16, 3, padding=(1, 0, 0)), nn.ELU(1.0, True),
                        nn.BatchNorm3d(16), nn.Conv3d(16, 16, 3, padding=(1, 2, 2)), nn.ELU(1.0, True))
    x = (torch.rand(4, 1, 8, 24, 24) > 0.6).float()
    _check_sequential(seq, x, G=2, check_input_grad=False)


def test_unit_eval_mode_uses_running_stats():
    seq = nn.Sequential(nn.BatchNorm3d(6), nn.Conv3d(6, 10, 3, padding=(1, 0, 0)), nn.ELU(1.0, True),
                        nn.BatchNorm3d(10), nn.Conv3d(10, 4, 3, padding=(1, 2, 2)), nn.ELU(1.0, True))
    _check_sequential(seq, torch.randn(2, 6, 4, 7, 8), training=False)


def test_chain_with_grouped_statistics():
    """G = 3 passes stacked along the batch == 3 separate calls (separate batch statistics, sequential running-stat
    updates, summed parameter gradients) — the CAE's core / penumbra / lesion encoder passes (Cae3D.py:105-107)."""
    seq = nn.Sequential(nn.BatchNorm3d(1), nn.Conv3d(1, 8, 3, padding=(1, 0, 0)), nn.ELU(1.0, True),
                        nn.BatchNorm3d(8), nn.Conv3d(8, 12, 3, stride=2, padding=1), nn.ELU(1.0, True),
                        nn.BatchNorm3d(12), nn.ConvTranspose3d(12, 6, 2, stride=2), nn.ELU(1.0, True),
                        nn.BatchNorm3d(6), nn.Conv3d(6, 1, 1), nn.Sigmoid())
    x = (torch.rand(6, 1, 6, 10, 10) > 0.6).float()
    _check_sequential(seq, x, G=3, check_input_grad=False)


def test_frozen_parameters_get_no_gradient():
    engine, _, _ = _mods()
    seq = nn.Sequential(nn.BatchNorm3d(4), nn.Conv3d(4, 6, 3), nn.ELU(1.0, True), nn.BatchNorm3d(6), nn.Conv3d(6, 2, 1), nn.Sigmoid()).cuda()
    for p in seq.parameters():
        p.requires_grad = False
    x = torch.randn(2, 4, 5, 6, 7).cuda().requires_grad_(True)
    y = engine.run_sequential(engine.SeqPlan(seq), x)
    y.sum().backward()
    assert all(p.grad is None for p in seq.parameters()) and x.grad is not None
    ref = copy.deepcopy(seq).cpu()
    xr = x.detach().cpu().requires_grad_(True)
    ref(xr).sum().backward()
    assert rel_l2(x.grad, xr.grad) < TOL_GRAD


def test_large_channel_count_statistics():
    """C = 800 (paper config fc) exercises the channel loop of the statistics kernels."""
    seq = nn.Sequential(nn.BatchNorm3d(800), nn.ConvTranspose3d(800, 12, 3), nn.ELU(1.0, True))
    _check_sequential(seq, torch.randn(2, 800, 1, 3, 3))


def test_raw_perfusion_statistics_do_not_cancel():
    """First U-Net BN sees raw CBV/TTD (mean far from 0, 20-voxel zero border): variance must not lose digits."""
    seq = nn.Sequential(nn.BatchNorm3d(2), nn.Conv3d(2, 4, 3), nn.LeakyReLU(0.01, True))
    x = torch.zeros(2, 2, 12, 14, 16)
    x[:, 0, 4:-4, 4:-4, 4:-4] = 1000.0 + torch.rand(2, 4, 6, 8)
    x[:, 1, 4:-4, 4:-4, 4:-4] = 40 * torch.rand(2, 4, 6, 8)
    _check_sequential(seq, x, check_input_grad=False)


@pytest.mark.parametrize("cin,size,training,pad", [
    (2, (12, 40, 44), True, 0),                 # Unet3D.py:16-18 block1 (tcgen05 weight gradient on 2 channels)
    (1, (6, 12, 14), True, 0), (3, (10, 20, 36), True, 0), (2, (8, 14, 16), False, 0),
    (1, (9, 30, 34), True, (1, 0, 0)),          # Cae3D.py:40-41: depth padded after BatchNorm -> per-tap border sums
    (2, (8, 14, 16), True, (1, 1, 1)), (1, (6, 9, 11), True, (1, 2, 2)), (3, (3, 1, 5), True, (1, 1, 2)),
    (1, (8, 14, 16), False, (1, 0, 0))])
def test_first_unit_bn_gradients_from_the_weight_gradient(cin, size, training, pad):
    """First unit of a network (Unet3D.py:16-18 block1: BatchNorm3d on the data -> Conv3d without padding): no input gradient is
    needed, so dgamma / dbeta / dW come from the weight gradient against the normalised input (sp_bn_grads_from_wgrad) instead
    of a dgrad + reduction.  Same bars as the dgrad path, which is run next to it; G = 2 statistics groups, data with a mean far
    from zero (raw CBV / TTD), tcgen05 (2 channels, large planes) and FFMA weight-gradient tiers."""
    engine, _, _ = _mods()
    torch.manual_seed(77 + cin)
    seq = nn.Sequential(nn.BatchNorm3d(cin), nn.Conv3d(cin, 16, 3, padding=pad), nn.LeakyReLU(0.01, True),
                        nn.BatchNorm3d(16), nn.Conv3d(16, 16, 3, padding=1), nn.LeakyReLU(0.01, True))
    x = torch.randn(4, cin, *size) * 3.0 + 5.0
    assert engine.FIRST_UNIT_SHORTCUT
    _check_sequential(copy.deepcopy(seq), x, training=training, G=2, check_input_grad=False)
    engine.FIRST_UNIT_SHORTCUT = False
    try:
        _check_sequential(copy.deepcopy(seq), x, training=training, G=2, check_input_grad=False)
    finally:
        engine.FIRST_UNIT_SHORTCUT = True


# ---------------------------------------------------------------------------------------------------- resampling
@pytest.mark.parametrize("size", [(4, 6, 8), (5, 7, 9), (2, 2, 2)])
@pytest.mark.parametrize("C", [5, 8])            # 8: the 128-bit kernels
def test_maxpool_with_ties(size, C):
    _, _, ops = _mods()
    torch.manual_seed(3)
    x = torch.randn(2, C, *size)
    x[:, :, :, : size[1] // 2] = 0.25            # constant region -> every window there is an 8-way tie
    x[0, 0] = torch.round(x[0, 0])               # many partial ties
    xr = x.clone().requires_grad_(True)
    yr = F.max_pool3d(xr, 2, 2)
    gy = torch.randn(yr.shape)
    yr.backward(gy)
    xv = ops.as_vol(x.cuda())
    y = ops.maxpool2_fwd(xv)
    gx = ops.maxpool2_bwd(xv, y, ops.as_vol(gy.cuda()))
    assert torch.equal(y.cpu(), yr.detach())
    assert torch.equal(gx.cpu(), xr.grad)      # bit exact: pure routing


@pytest.mark.parametrize("align", [False, True])
@pytest.mark.parametrize("size", [(3, 4, 5), (1, 2, 7), (10, 35, 35)])
@pytest.mark.parametrize("C,off,extra", [(6, 2, 3), (8, 4, 8)])       # (8, 4, 8): 16-byte aligned slice -> the 128-bit kernels
def test_trilinear_upsample(align, size, C, off, extra):
    _, _, ops = _mods()
    torch.manual_seed(5)
    x = torch.randn(2, C, *size)
    xr = x.clone().requires_grad_(True)
    yr = F.interpolate(xr, scale_factor=2, mode="trilinear", align_corners=align)
    gy = torch.randn(yr.shape)
    yr.backward(gy)
    xv = ops.as_vol(x.cuda())
    N, _, D, H, W = x.shape
    cat = ops.zeros_vol(N, C + extra, 2 * D, 2 * H, 2 * W, "cuda")     # write into a channel slice of a wider buffer
    ops.upsample2_fwd(xv, cat, off, align)
    assert rel_l2(cat[:, off:off + C], yr) < 1e-6
    assert float(cat[:, :off].abs().max()) == 0.0 and float(cat[:, off + C:].abs().max()) == 0.0
    gcat = ops.zeros_vol(N, C + extra, 2 * D, 2 * H, 2 * W, "cuda")
    gcat[:, off:off + C] = gy.cuda()
    gx = ops.upsample2_bwd(gcat, off, C, align)
    assert rel_l2(gx, xr.grad) < 1e-5


def test_crop_concat_and_channel_helpers():
    _, _, ops = _mods()
    torch.manual_seed(7)
    src = torch.randn(2, 4, 9, 10, 11).cuda()
    out = ops.zeros_vol(2, 7, 5, 6, 7, "cuda")
    offs = (2, 2, 2)
    ops.crop_into(ops.as_vol(src), out, 3, offs)
    assert torch.equal(out[:, 3:], src[:, :, 2:7, 2:8, 2:9]) and float(out[:, :3].abs().max()) == 0.0
    big = ops.as_vol(torch.randn(2, 4, 9, 10, 11).cuda())
    before = big.clone()
    g = ops.as_vol(torch.randn(2, 7, 5, 6, 7).cuda())
    ops.crop_add(big, g, 3, offs)
    want = before.clone()
    want[:, :, 2:7, 2:8, 2:9] += g[:, 3:]
    assert torch.allclose(big, want, atol=0, rtol=0)
    ch = ops.extract_channel(g, 5)
    assert ch.shape == (2, 1, 5, 6, 7) and ch.is_contiguous() and torch.equal(ch[:, 0], g[:, 5])
    tgt = ops.zeros_vol(2, 7, 5, 6, 7, "cuda")
    ops.insert_channel(ch, tgt, 1)
    assert torch.equal(tgt[:, 1], g[:, 5]) and float(tgt[:, 2:].abs().max()) == 0.0


# ---------------------------------------------------------------------------------------------------- losses
def test_dice_term():
    _, functions, _ = _mods()
    torch.manual_seed(11)
    for n in (1, 5, 4096, 100003):
        o = torch.rand(n)
        t = (torch.rand(n) > 0.7).float()
        orf = o.clone().double().requires_grad_(True)
        lr = 1.0 - (2.0 * (orf * t.double()).sum() + 1e-7) / ((orf * orf).sum() + (t.double() * t.double()).sum() + 1e-7)
        (lr * 0.37).backward()
        og = o.clone().cuda().requires_grad_(True)
        lg = functions.dice_term(og, t.cuda(), 1.0, 1e-7)
        (lg * 0.37).backward()
        assert abs(lg.item() - lr.item()) < 2e-6
        assert rel_l2(og.grad, orf.grad) < 1e-5


def test_dice_all_zero_prediction_and_target():
    _, functions, _ = _mods()
    o = torch.zeros(1000).cuda().requires_grad_(True)
    t = torch.zeros(1000).cuda()
    l = functions.dice_term(o, t, 1.0, 1e-7)
    l.backward()
    assert abs(l.item() - 0.0) < 1e-6 and torch.isfinite(o.grad).all()   # (0 + eps) / (0 + eps) = 1 -> loss 0


def test_binary_measures_match_oracle():
    """Thresholded overlap metrics (metrics.py:31-47) from the device confusion counts vs the oracle's restatement of medpy's
    definitions: counts are exact integers, the derived measures agree to the last bit; empty masks give 0.0."""
    _, _, ops = _mods()
    import stroke_oracle as O
    from stroke_prediction_b200.common import metrics
    torch.manual_seed(11)
    shape = (3, 1, 9, 31, 33)            # odd element count: tail handling
    res = torch.rand(shape)
    tgt = (torch.rand(shape) > 0.7).float()
    res[0, 0, 0, 0, :5] = 0.5            # exactly the threshold: '>' is strict (metrics.py:32)
    cases = [(res, tgt), (torch.zeros(shape), tgt), (res, torch.zeros(shape)), (torch.zeros(shape), torch.zeros(shape)),
             (torch.ones(shape), torch.ones(shape))]
    got = metrics.binary_measures_many([(r.cuda(), t.cuda()) for r, t in cases])
    for (r, t), g in zip(cases, got):
        ref = O.binary_measures(r, t)
        cnt = ops.binary_counts(r.cuda(), t.cuda(), 0.5).cpu().tolist()
        assert tuple(int(c) for c in cnt) == ref["counts"]
        assert sum(ref["counts"]) == r.numel()
        for k in ("dc", "precision", "sensitivity", "specificity"):
            assert getattr(g, k) == ref[k], (k, getattr(g, k), ref[k])
        sm = O.surface_measures(r.numpy(), t.numpy())
        assert g.hd == sm["hd"], (g.hd, sm["hd"])                       # sqrt of an exact integer squared distance
        assert g.assd == sm["assd"] or abs(g.assd - sm["assd"]) <= 1e-12 * sm["assd"], (g.assd, sm["assd"])
    single = metrics.binary_measures_torch(res.cuda(), tgt.cuda(), True)
    assert single.dc == O.binary_measures(res, tgt)["dc"]
    off = metrics.binary_measures_many([(res.cuda(), tgt.cuda())], surface_distances=False)[0]
    assert off.hd == float("inf") and off.assd == float("inf") and off.dc == single.dc


def _blobs(shape, seed, n=3):
    """A few random ellipsoidal blobs in an array of `shape` (last three axes spatial), soft values in [0, 1]."""
    import numpy as np
    rng = np.random.RandomState(seed)
    out = np.zeros(shape, dtype=np.float32)
    sp = shape[-3:]
    zz, yy, xx = np.meshgrid(*[np.arange(s) for s in sp], indexing="ij")
    flat = out.reshape((-1,) + tuple(sp))
    for i in range(flat.shape[0]):
        for _ in range(n):
            c = [rng.uniform(0.2, 0.8) * s for s in sp]
            r = [max(1.0, rng.uniform(0.08, 0.3) * s) for s in sp]
            d = ((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2
            flat[i] = np.maximum(flat[i], np.clip(1.2 - d, 0, 1))
    return torch.from_numpy(out)


@pytest.mark.parametrize("shape", [(2, 1, 12, 40, 36), (1, 1, 28, 64, 64), (3, 1, 7, 33, 29), (10, 30, 28), (3, 9, 20, 24),
                                   (4, 1, 28, 128, 128)])
def test_surface_distances_match_medpy_restatement(shape):
    """hd / assd (metrics.py:43-45 -> MedPy 0.3.0 `__surface_distances`) on the device against the scipy restatement, on the
    reference's B x 1 x D x H x W batches (extent-1 channel axis: the erosion is empty, every object voxel is "border", the batch
    axis is a lattice axis) and on arrays without extent-1 axes (true connectivity-1 erosion)."""
    _, _, ops = _mods()
    import stroke_oracle as O
    r, t = _blobs(shape, 3), _blobs(shape, 5)
    got = ops.surface_distances(r.cuda(), t.cuda(), 0.5).cpu().tolist()
    ref = O.surface_measures(r.numpy(), t.numpy())
    assert got[6] == ref["n_r"] and got[7] == ref["n_t"]
    assert got[0] == ref["hd"], (got[0], ref["hd"])
    assert abs(got[1] - ref["assd"]) <= 1e-12 * ref["assd"]
    assert abs(got[2] - ref["asd_rt"]) <= 1e-12 * ref["asd_rt"] and abs(got[3] - ref["asd_tr"]) <= 1e-12 * ref["asd_tr"]
    # the same volume as an NDHWC-strided single-channel tensor (what the models return)
    if len(shape) == 5:
        rv = ops.as_vol(r.cuda())
        got2 = ops.surface_distances(rv, t.cuda(), 0.5).cpu().tolist()
        assert got2 == got


def test_surface_distances_edge_cases():
    _, _, ops = _mods()
    import stroke_oracle as O
    shape = (2, 1, 6, 20, 20)
    z = torch.zeros(shape)
    one = torch.zeros(shape)
    one[1, 0, 3, 4, 5] = 1.0
    other = torch.zeros(shape)
    other[0, 0, 1, 10, 17] = 1.0
    full = torch.ones(shape)
    inf = float("inf")
    for r, t in [(z, one), (one, z), (z, z)]:
        got = ops.surface_distances(r.cuda(), t.cuda(), 0.5).cpu().tolist()
        assert got[0] == inf and got[1] == inf
    for r, t in [(one, one), (one, other), (full, one), (full, full)]:
        got = ops.surface_distances(r.cuda(), t.cuda(), 0.5).cpu().tolist()
        ref = O.surface_measures(r.numpy(), t.numpy())
        assert got[0] == ref["hd"] and abs(got[1] - ref["assd"]) <= 1e-12 * max(ref["assd"], 1e-300), (got, ref)
    # two single voxels in different samples: the batch axis counts as a lattice axis, sqrt(1 + 4 + 36 + 144)
    got = ops.surface_distances(one.cuda(), other.cuda(), 0.5).cpu().tolist()
    assert got[0] == (1 + 2 ** 2 + 6 ** 2 + 12 ** 2) ** 0.5
    with pytest.raises(RuntimeError):
        ops.surface_distances(torch.zeros(2, 2, 2, 2, 2, 2).cuda(), torch.zeros(2, 2, 2, 2, 2, 2).cuda())


@pytest.mark.parametrize("form", ["penu", "core"])
def test_signed_distance_map_matches_scipy(form):
    """SDM baseline (test_sdm_resampling.py:16-33): edt(inside) - edt(outside) per case volume."""
    _, _, ops = _mods()
    import stroke_oracle as O
    v = _blobs((28, 64, 64), 9)
    v[3, 10, 10:14] = 0.5                   # exactly the threshold: neither inside nor (for the '<' form) outside
    lt, sign = (True, 1.0) if form == "penu" else (False, -1.0)
    got = ops.signed_distance(v.cuda(), 0.5, lt, sign).cpu()
    ref = torch.from_numpy(O.signed_distance_map(v.numpy(), 0.5, lt, sign)).float()
    assert torch.equal(got, ref) or rel_max(got, ref) < 1e-6


def test_hinge_and_l1_including_exact_zeros():
    _, functions, _ = _mods()
    torch.manual_seed(13)
    a = torch.rand(3, 1, 4, 5, 6)
    b = torch.rand(3, 1, 4, 5, 6)
    b.view(-1)[::3] = a.view(-1)[::3]            # d == 0 exactly where both sigmoids saturate
    for mode, fn in ((0, lambda d: torch.mean(torch.abs(d) - d)), (1, lambda d: torch.mean(torch.abs(d)))):
        ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        lr = fn(ar - br)
        (2.5 * lr).backward()
        ag, bg = a.clone().cuda().requires_grad_(True), b.clone().cuda().requires_grad_(True)
        lg = functions.hinge_mean(ag, bg) if mode == 0 else functions.l1_mean(ag, bg)
        (2.5 * lg).backward()
        assert abs(lg.item() - lr.item()) < 1e-6
        assert rel_max(ag.grad, ar.grad) < 1e-6 and rel_max(bg.grad, br.grad) < 1e-6


def test_latent_interpolation():
    _, functions, _ = _mods()
    torch.manual_seed(17)
    zc, zp = torch.randn(4, 12, 1, 3, 3), torch.randn(4, 12, 1, 3, 3)
    s = torch.tensor([-1.0, 0.0, 0.4, 2.0]).reshape(4, 1, 1, 1, 1)     # no clamping of the step (Cae3D.py:78-89)
    zcr, zpr, sr = zc.clone().requires_grad_(True), zp.clone().requires_grad_(True), s.clone().requires_grad_(True)
    out_r = zcr + sr * (zpr - zcr)
    g = torch.randn(out_r.shape)
    out_r.backward(g)
    zcg, zpg, sg = (t.clone().cuda().requires_grad_(True) for t in (zc, zp, s))
    out_g = functions.latent_interp(zcg, zpg, sg)
    out_g.backward(g.cuda())
    assert rel_max(out_g, out_r) < 1e-6
    assert rel_max(zcg.grad, zcr.grad) < 1e-6 and rel_max(zpg.grad, zpr.grad) < 1e-6
    assert sg.grad.shape == s.shape and rel_max(sg.grad, sr.grad) < 1e-5


# ---------------------------------------------------------------------------------------------------- optimizer
def test_fused_adam_matches_torch_adam():
    from stroke_prediction_b200.optim import FusedAdam
    torch.manual_seed(19)
    shapes = [(16, 1, 3, 3, 3), (16,), (5000,), (1,), (100, 40, 3, 3, 3)]
    ref = [torch.randn(s).requires_grad_(True) for s in shapes]
    mine = [p.detach().clone().cuda().requires_grad_(True) for p in ref]
    o_ref = torch.optim.Adam(ref, lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-5)
    o_mine = FusedAdam(mine, lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-5)
    for t in range(5):
        beta1 = 0.5 + 0.1 * t if t < 4 else 0.9          # adapt_betas schedule, CaeReconstructionLearner.py:28-40
        for o in (o_ref, o_mine):
            for g in o.param_groups:
                g["betas"] = (beta1, 0.999)
        for p, q in zip(ref, mine):
            gr = torch.randn(p.shape)
            p.grad = gr.clone()
            q.grad = gr.clone().cuda()
        o_ref.step()
        o_mine.step()
    for p, q in zip(ref, mine):
        assert rel_max(q, p) < 2e-6
    sd = o_mine.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    assert rel_max(sd["state"][2]["exp_avg_sq"], o_ref.state_dict()["state"][2]["exp_avg_sq"]) < 1e-6


def test_fused_adam_adopts_torch_adam():
    from stroke_prediction_b200.optim import FusedAdam
    p = torch.randn(300).cuda().requires_grad_(True)
    ref = p.detach().clone().cpu().requires_grad_(True)
    o = torch.optim.Adam([p], lr=1e-3, betas=(0.99, 0.999), weight_decay=1e-5)
    f = FusedAdam.from_torch(o)
    assert f.param_groups is o.param_groups
    o2 = torch.optim.Adam([ref], lr=1e-3, betas=(0.99, 0.999), weight_decay=1e-5)
    for _ in range(2):
        g = torch.randn(300)
        p.grad, ref.grad = g.cuda(), g.clone()
        f.step()
        o2.step()
    assert rel_max(p, ref) < 2e-6


def test_no_cpu_fallback():
    _, functions, ops = _mods()
    with pytest.raises(RuntimeError):
        ops.as_vol(torch.zeros(1, 1, 2, 2, 2))
    with pytest.raises(RuntimeError):
        functions.dice_term(torch.rand(10), torch.rand(10))


# ---------------------------------------------------------------------------------------------------- data transforms (n2)
def _np_to_batch(v):
    """reference numpy sample [x][y][z][c] -> torch B x C x Z x Y x X on the device"""
    return torch.from_numpy(v.copy()).permute(3, 2, 1, 0).unsqueeze(0).contiguous().cuda()


def _batch_to_np(t):
    return t[0].permute(3, 2, 1, 0).contiguous().cpu().numpy()


def test_device_transforms_against_reference_fixture():
    """common/data.py:215-380 on the device vs outputs of the reference's own transform classes (tests/golden/transforms_tiny,
    generated by oracle/make_golden.py): elastic deformation replayed with the reference's numpy noise stream, per-slice zoom
    (nearest / bilinear), constant padding, hemispheric flip, random patch, ToTensor."""
    import random
    import numpy as np
    from util import load
    from stroke_prediction_b200.common import data, gpu_transforms as T
    fx = load("transforms_tiny")
    labels, images = fx["labels"], fx["images"]
    X, Y, Z, C = labels.shape
    batch = {data.KEY_LABELS: _np_to_batch(labels), data.KEY_IMAGES: _np_to_batch(images), data.KEY_CASE_ID: [3]}
    assert np.array_equal(batch[data.KEY_LABELS][0].cpu().numpy(), fx["to_tensor_labels"])
    tt = T.ToTensor()({data.KEY_LABELS: labels, data.KEY_IMAGES: images})
    assert torch.equal(tt[data.KEY_LABELS], batch[data.KEY_LABELS])
    # elastic: the reference draws dx, dy, dz per channel from ONE running RandomState, labels first, then images
    rs = np.random.RandomState(int(fx["elastic_seed"]))
    def draw(n_ch):
        noise = np.zeros((3, 1, n_ch, Z, Y, X))
        for c in range(n_ch):
            for f in range(3):
                noise[f, 0, c] = rs.rand(X, Y, Z).transpose(2, 1, 0)
        return torch.from_numpy(noise)
    n_l, n_i = draw(C), draw(images.shape[3])
    out = T.ElasticDeform(100, 4, apply_to_images=True)(batch, noise=n_l, image_noise=n_i)
    got_l, got_i = _batch_to_np(out[data.KEY_LABELS]), _batch_to_np(out[data.KEY_IMAGES])
    assert np.abs(got_l - fx["elastic_labels"]).max() < 2e-6, np.abs(got_l - fx["elastic_labels"]).max()
    assert np.abs(got_i - fx["elastic_images"]).max() < 2e-5 and rel_l2(got_i, fx["elastic_images"]) < 1e-6
    assert float(np.abs(got_l - labels).max()) > 0.5           # it really deformed something
    # zoom
    for sf, tag in ((0.5, "0p5"), (0.75, "0p75")):
        for mode in ("nearest", "bilinear"):
            z = T.ResamplePlaneXY(sf, mode)(batch)
            for key, name in ((data.KEY_IMAGES, "images"), (data.KEY_LABELS, "labels")):
                want = fx["zoom_%s_%s_%s" % (tag, mode, name)]
                got = _batch_to_np(z[key])
                assert got.shape == want.shape
                if mode == "nearest":
                    assert np.array_equal(got, want), (sf, mode, name)
                else:
                    assert np.abs(got - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), (sf, mode, name)
    up = T.ResamplePlaneXY(1.5, "bilinear")(batch)                           # factors > 1 (the reference's class cannot)
    import stroke_oracle as O
    assert np.abs(_batch_to_np(up[data.KEY_IMAGES]) - O.resample_plane_xy(images, 1.5, 1)).max() < 2e-5
    # pad, flip, patch
    p = T.PadImages(3, 2, 1, pad_value=0.5)(batch)
    assert np.array_equal(_batch_to_np(p[data.KEY_IMAGES]), fx["pad_images"])
    random.seed(5)
    for want_flip in fx["flip_decisions_seed5"]:
        f = T.HemisphericFlip()(batch)
        assert np.array_equal(_batch_to_np(f[data.KEY_LABELS]), fx["flipped_labels"] if want_flip else labels)
    f = T.HemisphericFlipFixedToCaseId(split_id=2)(batch)
    assert np.array_equal(_batch_to_np(f[data.KEY_LABELS]), fx["flipped_labels"])
    f = T.HemisphericFlipFixedToCaseId(split_id=5)(batch)
    assert np.array_equal(_batch_to_np(f[data.KEY_LABELS]), labels)
    random.seed(9)
    rp = T.RandomPatch(16, 12, 6, 2, 1, 1)(batch)
    assert np.array_equal(_batch_to_np(rp[data.KEY_IMAGES]), fx["patch_images_seed9"])
    assert np.array_equal(_batch_to_np(rp[data.KEY_LABELS]), fx["patch_labels_seed9"])


def test_elastic_deform_named_shape_against_scipy_restatement():
    """ElasticDeform at the CAE's volume size (128 x 128 x 28, batch 2, 3 label channels) vs the scipy restatement."""
    import numpy as np
    import stroke_oracle as O
    from stroke_prediction_b200.common import data, gpu_transforms as T
    b = data.synthetic_cae_batch(2, seed=4)
    lab = b[data.KEY_LABELS]
    rs = np.random.RandomState(123)
    B, C, D, H, W = lab.shape
    noise = np.zeros((3, B, C, D, H, W))
    want = np.zeros((B, C, D, H, W), np.float32)
    for n in range(B):
        for c in range(C):
            fields = [rs.rand(W, H, D) for _ in range(3)]
            for f in range(3):
                noise[f, n, c] = fields[f].transpose(2, 1, 0)
            vol = lab[n, c].permute(2, 1, 0).contiguous().numpy()                       # [x][y][z]
            want[n, c] = O.elastic_transform(vol, fields, 100, 4).transpose(2, 1, 0)
    out = T.ElasticDeform()({data.KEY_LABELS: lab.cuda()}, noise=torch.from_numpy(noise))[data.KEY_LABELS]
    assert np.abs(out.cpu().numpy() - want).max() < 5e-6
    # device-generated noise: same statistics (values stay in [0, 1], mass roughly preserved), deterministic per generator
    g = torch.Generator(device="cuda").manual_seed(3)
    o1 = T.ElasticDeform(generator=g)({data.KEY_LABELS: lab.cuda()})[data.KEY_LABELS]
    g = torch.Generator(device="cuda").manual_seed(3)
    o2 = T.ElasticDeform(generator=g)({data.KEY_LABELS: lab.cuda()})[data.KEY_LABELS]
    assert torch.equal(o1, o2) and float(o1.min()) >= 0.0 and float(o1.max()) <= 1.0
    assert abs(float(o1.sum()) / float(lab.sum()) - 1.0) < 0.2


@pytest.mark.parametrize("mode", [5, 1])
def test_tensor_core_tier_is_deterministic(mode):
    """The persistent, warp-specialised tcgen05 kernel (TMA producer, staging warps, MMA issuers, epilogue groups linked by
    mbarriers) must give bit-identical results launch after launch: several work items per CTA, ragged tiles, padding.  (A missing
    cross-proxy fence between the staging warps' reads and the next TMA write showed up as a few stale voxels in ~4 % of launches.)"""
    engine, _, ops = _mods()
    torch.manual_seed(77)
    seq = nn.Sequential(nn.BatchNorm3d(16), nn.Conv3d(16, 16, 3, padding=(1, 2, 2)), nn.ELU(1.0)).cuda().eval()
    x = (torch.randn(3, 16, 14, 45, 61) * 1.5 + 0.3).cuda()
    ops.set_tc_terms(mode)
    try:
        plan = engine.SeqPlan(seq)
        with torch.no_grad():
            first = engine.run_sequential(plan, x).clone()
            for _ in range(150):
                assert torch.equal(engine.run_sequential(plan, x), first)
        xg = x.clone().requires_grad_(True)
        engine.run_sequential(plan, xg).square().sum().backward()
        g0 = xg.grad.clone()
        for _ in range(30):
            xg.grad = None
            engine.run_sequential(plan, xg).square().sum().backward()
            assert torch.equal(xg.grad, g0)
    finally:
        ops.set_tc_terms(5)


@pytest.mark.parametrize("layer", ["s1", "s2", "k2s2"])
@pytest.mark.parametrize("max_ctas", [0, 5])
def test_tensor_core_weight_gradient_is_deterministic(layer, max_ctas):
    """The weight-gradient kernels (three staging groups running ahead through a ring of planes, elected MMA issue, periodic TMEM
    drains, fixed-order reduction of the per-CTA partials) must give bit-identical dW launch after launch, also when every CTA
    walks many tile columns."""
    engine, _, ops = _mods()
    torch.manual_seed(78)
    conv = {"s1": nn.Conv3d(16, 16, 3, padding=(1, 2, 2)), "s2": nn.Conv3d(16, 24, 3, stride=2, padding=1),
            "k2s2": nn.ConvTranspose3d(16, 16, 2, stride=2)}[layer]
    seq = nn.Sequential(nn.BatchNorm3d(16), conv, nn.ELU(1.0)).cuda().train()
    size = (7, 20, 23) if layer == "k2s2" else (14, 45, 61)
    x = (torch.randn(4, 16, *size) * 1.5 + 0.3).cuda()
    ops.set_wgrad_tc_options(2, max_ctas)
    try:
        plan = engine.SeqPlan(seq)
        first = None
        for _ in range(40):
            for p in seq.parameters():
                p.grad = None
            engine.run_sequential(plan, x).square().sum().backward()
            g = conv.weight.grad.clone()
            if first is None:
                first = g
                assert torch.isfinite(first).all() and float(first.abs().max()) > 0
            assert torch.equal(g, first)
    finally:
        ops.set_wgrad_tc_options(2, 0)
