"""CPU: host-side drop-in contract — module constructors, state_dict names/shapes, DTO behaviour,
plan parsing, C-ABI library exports.  No compute call reaches the GPU library here."""
import ctypes
import os
import re

import pytest
import torch

import stroke_prediction_b200 as sp
from stroke_prediction_b200 import _lib, engine
from stroke_prediction_b200.common.dto import CaeDto as CaeDtoUtil
from stroke_prediction_b200.common.dto import MetricMeasuresDto as MM
from stroke_prediction_b200.common.dto import UnetDto as UnetDtoUtil
from stroke_prediction_b200.common.dto.Dto import Dto
from stroke_prediction_b200.common.model.Cae3D import Cae3D, Dec3D, Enc3D, Enc3DCtp, Enc3DStep
from stroke_prediction_b200.common.model.Unet3D import Unet3D, crop
from util import load, state_from

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cae_state_dict_matches_reference_layout():
    fx = load("cae_step_tiny")
    ref = state_from(fx, "sd0/")
    ch = [int(c) for c in fx["channels"]]
    cae = Cae3D(Enc3DStep(56, 28, ch, 5, 1.0), Dec3D(56, 28, ch, 5, 1.0))
    mine = cae.state_dict()
    assert list(mine.keys()) == list(ref.keys())
    for k in ref:
        assert tuple(mine[k].shape) == tuple(ref[k].shape), k
    cae.load_state_dict(ref)   # reference weights load as they are


def test_unet_state_dict_matches_reference_layout():
    fx = load("unet_tiny")
    ref = state_from(fx, "sd0/")
    unet = Unet3D([int(c) for c in fx["channels"]])
    mine = unet.state_dict()
    assert list(mine.keys()) == list(ref.keys())
    for k in ref:
        assert tuple(mine[k].shape) == tuple(ref[k].shape), k


def test_parameter_counts_of_named_configs():
    n = lambda m: sum(p.numel() for p in m.parameters())
    assert n(Unet3D([2, 16, 32, 64, 32, 16, 32, 2])) == 355014
    ch = [1, 16, 24, 32, 100, 200, 1]
    assert n(Enc3D(128, 28, ch, 5, 1.0)) == 752350 and n(Dec3D(128, 28, ch, 5, 1.0)) == 722805
    ch = [1, 16, 24, 32, 100, 800, 1]
    assert n(Enc3D(128, 28, ch, 5, 1.0)) == 2372950 and n(Dec3D(128, 28, ch, 5, 1.0)) == 2344005
    assert n(Enc3DStep(128, 28, ch, 5, 1.0)) - n(Enc3D(128, 28, ch, 5, 1.0)) == 45


def test_constructor_asserts():
    with pytest.raises(AssertionError):
        Enc3D(126, 28, [1, 16, 24, 32, 100, 200, 1], 5, 1.0)
    with pytest.raises(AssertionError):
        Enc3DCtp(128, 28, [2, 16, 24, 32, 100, 200, 1], 5, 1.0, [20, 20, 20])


def test_freeze_toggles_requires_grad():
    cae = Cae3D(Enc3D(56, 28, [1, 4, 6, 8, 10, 12, 1], 5, 1.0), Dec3D(56, 28, [1, 4, 6, 8, 10, 12, 1], 5, 1.0))
    cae.freeze(True)
    assert not any(p.requires_grad for p in cae.parameters())
    cae.freeze(False)
    assert all(p.requires_grad for p in cae.parameters())


def test_plan_shapes():
    ch = [1, 16, 24, 32, 100, 200, 1]
    enc, dec = Enc3D(128, 28, ch, 5, 1.0), Dec3D(128, 28, ch, 5, 1.0)
    pe, pd = engine.SeqPlan(enc.encoder), engine.SeqPlan(dec.decoder)
    assert len(pe.units) == 10 and len(pd.units) == 12
    assert pe.out_shape((8, 1, 28, 128, 128)) == (8, 200, 1, 10, 10)
    assert pd.out_shape((8, 200, 1, 10, 10)) == (8, 1, 28, 128, 128)
    # 64 x 256 x 256 does not round-trip, 60 and 68 do
    assert pd.out_shape(pe.out_shape((1, 1, 60, 256, 256))) == (1, 1, 60, 256, 256)
    assert pd.out_shape(pe.out_shape((1, 1, 64, 256, 256))) != (1, 1, 64, 256, 256)
    step = Enc3DStep(128, 28, ch, 5, 1.0)
    ps = step._step_plan()
    assert [u.act for u in ps.units] == [engine.ACT_ELU, engine.ACT_ELU, engine.ACT_SIGMOID]


def test_models_refuse_cpu_tensors():
    ch = [1, 4, 6, 8, 10, 12, 1]
    cae = Cae3D(Enc3D(56, 28, ch, 5, 1.0), Dec3D(56, 28, ch, 5, 1.0))
    dto = CaeDtoUtil.init_dto(None, torch.zeros(2, 1, 1, 1, 1), None, None, None, None,
                              torch.zeros(2, 1, 28, 56, 56), None, None)
    with pytest.raises(RuntimeError, match="no CPU path"):
        cae(dto)
    unet = Unet3D([2, 4, 6, 8, 6, 4, 6, 2])
    with pytest.raises(RuntimeError, match="no CPU path"):
        unet(UnetDtoUtil.init_dto(torch.zeros(1, 2, 44, 44, 44)))


def test_dto_semantics():
    dto = CaeDtoUtil.init_dto(1, 2, 3, 4, None, None, None, None, None)
    assert dto.flag == CaeDtoUtil.FLAG_DEFAULT
    assert dto.latents.gtruth._is_empty() and dto.reconstructions.inputs._is_empty()
    dto.latents.gtruth.core = 5
    assert not dto.latents.gtruth._is_empty()
    assert dto.latents._is_empty()          # nested results are dropped, like the reference (Dto.py:41)
    text = str(dto)
    assert "[x] given_variables" in text and "[ ] core" in text
    assert dict(Dto(a=1, b=None)) == {"a": 1, "b": None}
    u = UnetDtoUtil.init_dto(1, 2, 3)
    assert u.given_variables.core == 2 and u.outputs.core is None


def test_metric_accumulators():
    a = MM.init_dto(loss=1.0, core_dc=0.5)
    b = MM.init_dto(loss=3.0, core_dc=0.7)
    a.add(b)
    a.div(2)
    assert a.loss == 2.0 and abs(a.core.dc - 0.6) < 1e-12 and a.penu.dc is None
    with pytest.raises(Exception):
        a.add(a.core)


def test_crop_view():
    t = torch.arange(2 * 1 * 6 * 8 * 10).reshape(2, 1, 6, 8, 10)
    like = torch.zeros(2, 1, 2, 4, 4)
    c = crop(t, like, dims=[2, 3, 4])
    assert c.shape == like.shape and c[0, 0, 0, 0, 0] == t[0, 0, 2, 2, 3]


def test_reference_aliases():
    sp.install_reference_aliases()
    import common.model.Cae3D as alias
    assert alias.Cae3D is Cae3D


def test_header_symbols_are_bound_and_exported():
    header = open(os.path.join(ROOT, "include", "stroke_b200.h")).read()
    declared = set(re.findall(r"\b(sp_[a-z0-9_]+)\s*\(", header, flags=re.I))
    declared = {d for d in declared if not d.isupper()}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib.sp_version.restype = ctypes.c_int
    assert lib.sp_version() == 100
    assert ctypes.sizeof(_lib.SpConvDesc) == 18 * 4 and ctypes.sizeof(_lib.SpAdamTensor) == 48


def test_argument_errors_do_not_launch():
    """Negative return + message for bad descriptors; nothing touches a GPU."""
    lib = _lib.load()
    d = _lib.SpConvDesc()
    rc = lib.sp_corr(ctypes.byref(d), None, None, None, None, None, 1, None, None, 0, None)
    assert rc < 0 and b"conv" in lib.sp_last_error()
    rc = lib.sp_adam_multi(None, 0, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 1.0, 0, None)
    assert rc < 0


def test_fused_adam_load_state_dict_keeps_the_scheduler_link():
    """torch's load_state_dict replaces param_groups / state by new objects; a scheduler built on the script's
    torch.optim.Adam (whose groups FusedAdam.from_torch shares) must keep driving the optimizer that steps after a resume."""
    from stroke_prediction_b200.optim import FusedAdam
    w = [torch.nn.Parameter(torch.randn(3, 3)), torch.nn.Parameter(torch.randn(5))]
    opt = torch.optim.Adam(w, lr=1e-3, weight_decay=1e-5)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, [1])
    fused = FusedAdam.from_torch(opt)
    assert fused.param_groups is opt.param_groups and fused.state is opt.state
    # a checkpoint written by another torch.optim.Adam run
    src = torch.optim.Adam([torch.nn.Parameter(t.detach().clone()) for t in w], lr=5e-4, weight_decay=1e-5, betas=(0.7, 0.999))
    for p in src.param_groups[0]['params']:
        p.grad = torch.ones_like(p)
    src.step()
    groups, state = fused.param_groups, fused.state
    fused.load_state_dict(src.state_dict())
    assert fused.param_groups is groups and fused.state is state and opt.param_groups is groups and opt.state is state
    assert fused.param_groups[0]['lr'] == 5e-4 and fused.param_groups[0]['betas'] == (0.7, 0.999)
    assert fused.param_groups[0]['params'][0] is w[0]
    assert float(fused.state[w[0]]['step']) == 1.0 and torch.equal(fused.state[w[1]]['exp_avg'], src.state[src.param_groups[0]['params'][1]]['exp_avg'])
    sched.step()                                   # milestone 1: lr *= 0.1 on the groups the fused optimizer reads
    assert abs(fused.param_groups[0]['lr'] - 5e-5) < 1e-15


def test_legacy_upsample_pickle_defaults_to_align_corners_true():
    """nn.Upsample pickled by torch 0.3.1 has no align_corners attribute and interpolated with corner alignment."""
    import warnings
    unet = Unet3D([2, 4, 8, 16, 8, 4, 8, 2])
    assert engine._align(unet, "upsa34") is False                    # installed torch: None -> False
    del unet.upsa34.__dict__["align_corners"]
    engine._warned_legacy_upsample = False
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert engine._align(unet, "upsa34") is True
    assert any("align_corners" in str(x.message) for x in w)
    unet.align_corners = False
    assert engine._align(unet, "upsa34") is False and engine._align(unet, "upsa45") is False
    unet.align_corners = True
    assert engine._align(unet, "upsa45") is True


def test_cpu_snapshot_leaves_the_live_module_alone():
    from stroke_prediction_b200.learner.Learner import Learner
    m = Cae3D(Enc3D(56, 28, [1, 4, 6, 8, 10, 12, 1], 5, 1.0), Dec3D(56, 28, [1, 4, 6, 8, 10, 12, 1], 5, 1.0))
    grads = []
    for p in m.parameters():
        p.grad = torch.zeros_like(p)
        grads.append(p.grad)
    snap = Learner.cpu_snapshot(m)
    assert all(p.grad is g for p, g in zip(m.parameters(), grads))
    assert all(q.grad is None for q in snap.parameters())
    assert all(torch.equal(p, q) and p is not q for p, q in zip(m.parameters(), snap.parameters()))
    assert '_sp_plans' not in vars(snap.enc)


def test_adopt_checkpoint_copies_into_the_live_module():
    from stroke_prediction_b200.learner.Learner import Learner
    ch = [1, 4, 6, 8, 10, 12, 1]
    live = Cae3D(Enc3D(56, 28, ch, 5, 1.0), Dec3D(56, 28, ch, 5, 1.0))
    loaded = Cae3D(Enc3D(56, 28, ch, 5, 1.0), Dec3D(56, 28, ch, 5, 1.0))
    params = list(live.parameters())
    out = Learner.adopt_checkpoint(live, loaded, cuda=False)
    assert out is live and all(p is q for p, q in zip(live.parameters(), params))
    assert all(torch.equal(a, b) for a, b in zip(live.state_dict().values(), loaded.state_dict().values()))
    other = Cae3D(Enc3D(56, 28, [1, 4, 6, 8, 10, 16, 1], 5, 1.0), Dec3D(56, 28, [1, 4, 6, 8, 10, 16, 1], 5, 1.0))
    assert Learner.adopt_checkpoint(live, other, cuda=False) is other          # different structure: replaced, like the reference
