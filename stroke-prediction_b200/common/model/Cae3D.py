"""Drop-in CAE modules: constructor signatures, attribute names, ``forward(dto) -> dto`` and state_dict layout of the
reference's common/model/Cae3D.py, with every arithmetic op executed by libstroke_b200.so.

The ``nn.Sequential`` containers only *own* the parameters/buffers (so ``enc.encoder.1.weight`` etc. keep the
reference's names and shapes and reference checkpoints load); the forward never calls them — it hands the Sequential
to the fused engine (engine.SeqPlan), one autograd node per pass.
"""
import torch
import torch.nn as nn

from ... import engine, functions, ops
from ..dto import CaeDto as CaeDtoUtil
from ..dto.CaeDto import CaeDto
from .. import data


def _require_cuda(t, what):
    if t is not None and not t.is_cuda:
        raise RuntimeError("%s: stroke_prediction_b200 runs on CUDA tensors only — there is no CPU path" % what)


class _PlanCache:
    """Lazily built SeqPlan for a Sequential attribute; rebuilt when the attribute is re-assigned
    (train_interpolationstep_after_reconstruction.py:25 swaps ``enc.encoder``)."""

    def get(self, owner, attr):
        seq = getattr(owner, attr)
        cache = owner.__dict__.setdefault('_sp_plans', {})
        hit = cache.get(attr)
        if hit is None or hit[0] is not seq:
            hit = (seq, engine.SeqPlan(seq))
            cache[attr] = hit
        return hit[1]


_plans = _PlanCache()


def _run_grouped(owner, attr, inputs, what):
    """Run owner.<attr> (a Sequential) on every non-None input; same-shaped ones share one stacked pass."""
    present = [(i, t) for i, t in enumerate(inputs) if t is not None]
    outs = [None] * len(inputs)
    if not present:
        return outs
    for _, t in present:
        _require_cuda(t, what)
    plan = _plans.get(owner, attr)
    if GROUP_PASSES and len({tuple(t.shape) for _, t in present}) == 1:
        ys = engine.run_sequential_grouped(plan, [t for _, t in present])
    else:
        ys = [engine.run_sequential(plan, t) for _, t in present]
    for (i, _), y in zip(present, ys):
        outs[i] = y
    return outs


GROUP_PASSES = True   # False: one launch sequence per pass, literally like the reference (used by A/B tests)


class CaeBase(nn.Module):

    def __init__(self, size_input_xy=128, size_input_z=28, channels=[1, 16, 32, 64, 128, 1024, 128, 1], n_ch_global=2,
                 alpha=0.01, inner_xy=12, inner_z=3):
        super().__init__()
        assert size_input_xy % 4 == 0 and size_input_z % 4 == 0
        self.n_ch_origin = channels[1]
        self.n_ch_down2x = channels[2]
        self.n_ch_down4x = channels[3]
        self.n_ch_down8x = channels[4]
        self.n_ch_fc = channels[5]

        self._inner_ch = self.n_ch_down8x
        self._inner_xy = inner_xy
        self._inner_z = inner_z

        self.n_ch_global = n_ch_global
        self.n_input = channels[0]
        self.n_classes = channels[-1]
        self.alpha = alpha

    def freeze(self, freeze=False):
        requires_grad = not freeze
        for param in self.parameters():
            param.requires_grad = requires_grad

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop('_sp_plans', None)   # kernel-side caches (packed weights) never enter a checkpoint
        return state


def _unit(kind, cin, cout, k, stride, padding, act):
    """[BatchNorm3d, conv, activation] triple in the reference's pre-norm order."""
    conv = (nn.ConvTranspose3d(cin, cout, k, stride=stride, padding=padding, output_padding=0) if kind == 'T'
            else nn.Conv3d(cin, cout, k, stride=stride, padding=padding))
    return [nn.BatchNorm3d(cin), conv, act]


class Enc3D(CaeBase):
    def __init__(self, size_input_xy, size_input_z, channels, n_ch_global, alpha):
        super().__init__(size_input_xy, size_input_z, channels, n_ch_global, alpha, inner_xy=10, inner_z=3)
        c0, c1, c2, c4, c8, cf = (self.n_input, self.n_ch_origin, self.n_ch_down2x, self.n_ch_down4x,
                                  self.n_ch_down8x, self.n_ch_fc)
        same_z = (1, 0, 0)
        # (cin, cout, stride, padding) of the ten 3x3x3 convolutions, Cae3D.py:40-75
        table = [(c0, c1, 1, same_z), (c1, c1, 1, same_z),
                 (c1, c2, 2, 1),
                 (c2, c2, 1, same_z), (c2, c2, 1, same_z),
                 (c2, c4, 2, 1),
                 (c4, c4, 1, same_z), (c4, c4, 1, same_z),
                 (c4, c8, 2, 0),
                 (c8, cf, 1, 0)]
        mods = []
        for cin, cout, stride, padding in table:
            mods += _unit('C', cin, cout, 3, stride, padding, nn.ELU(self.alpha, True))
        self.encoder = nn.Sequential(*mods)

    def _interpolate(self, latent_core, latent_penu, step):
        assert step is not None, 'Step must be given for interpolation!'
        if latent_core is None or latent_penu is None:
            return None
        return functions.latent_interp(latent_core, latent_penu, step)

    def _forward_single(self, input_image):
        if input_image is None:
            return None
        _require_cuda(input_image, 'Enc3D')
        return engine.run_sequential(_plans.get(self, 'encoder'), input_image)

    def _get_step(self, dto: CaeDto):
        return dto.given_variables.time_to_treatment

    def _forward_group(self, inputs):
        """The reference calls the encoder once per input, in order (Cae3D.py:105-107,113-114).  Same-shaped inputs
        are stacked along the batch and run as ONE pass with per-call BatchNorm statistics and sequential
        running-stat updates, which is numerically the same thing with a third of the launches."""
        return _run_grouped(self, 'encoder', inputs, 'Enc3D')

    def forward(self, dto: CaeDto):
        step = self._get_step(dto)
        given, latents = dto.given_variables, dto.latents

        if dto.flag == CaeDtoUtil.FLAG_GTRUTH or dto.flag == CaeDtoUtil.FLAG_DEFAULT:
            assert latents.gtruth._is_empty()  # do not overwrite earlier results
            latents.gtruth.core, latents.gtruth.penu, latents.gtruth.lesion = self._forward_group(
                [given.gtruth.core, given.gtruth.penu, given.gtruth.lesion])
            latents.gtruth.interpolation = self._interpolate(latents.gtruth.core, latents.gtruth.penu, step)
        if dto.flag == CaeDtoUtil.FLAG_INPUTS or dto.flag == CaeDtoUtil.FLAG_DEFAULT:
            assert latents.inputs._is_empty()
            latents.inputs.core, latents.inputs.penu = self._forward_group([given.inputs.core, given.inputs.penu])
            latents.inputs.interpolation = self._interpolate(latents.inputs.core, latents.inputs.penu, step)
        return dto


class Enc3DStep(Enc3D):
    def __init__(self, size_input_xy, size_input_z, channels, n_ch_global, alpha):
        super().__init__(size_input_xy, size_input_z, channels, n_ch_global, alpha)
        g = self.n_ch_global
        self.reduce = nn.Sequential(
            nn.Conv3d(g, g, 1), nn.ELU(self.alpha, True),
            nn.Conv3d(g, g // 2, 1), nn.ELU(self.alpha, True),
        )
        self.step = nn.Conv3d(g // 2, 1, 1)
        torch.nn.init.normal_(self.step.weight, 0, 0.001)   # Cae3D.py:133-134
        torch.nn.init.normal_(self.step.bias, 0.5, 0.01)
        self.sigmoid = nn.Sigmoid()

    def _step_plan(self):
        cache = self.__dict__.setdefault('_sp_plans', {})
        key = (self.reduce, self.step)
        hit = cache.get('_step')
        if hit is None or hit[0][0] is not key[0] or hit[0][1] is not key[1]:
            hit = (key, engine.SeqPlan(list(self.reduce.children()) + [self.step, self.sigmoid]))
            cache['_step'] = hit
        return hit[1]

    def _get_step(self, dto: CaeDto):
        step = dto.given_variables.time_to_treatment
        if step is None:
            glob = dto.given_variables.globals
            _require_cuda(glob, 'Enc3DStep')
            step = engine.run_sequential(self._step_plan(), glob)
        return step


class Enc3DCtp(Enc3D):
    def __init__(self, size_input_xy, size_input_z, channels, n_ch_global, alpha, padding):
        Enc3D.__init__(self, size_input_xy, size_input_z, channels, n_ch_global, alpha)
        assert channels[0] > 2, 'At least 3 channels required to process input'
        self._padding = padding

    def _stack(self, mask, cbv_padded, ttd_padded):
        """cat((mask, crop(cbv), crop(ttd)), dim=1) written straight into one NDHWC volume (Cae3D.py:153-162)."""
        if any(t.requires_grad for t in (mask, cbv_padded, ttd_padded)):
            p = self._padding
            crop = lambda t: t[:, :, p[0]:-p[0], p[1]:-p[1], p[2]:-p[2]]
            return torch.cat((mask, crop(cbv_padded), crop(ttd_padded)), dim=data.DIM_CHANNEL_TORCH3D_5)
        mask, cbv_padded, ttd_padded = ops.as_vol(mask), ops.as_vol(cbv_padded), ops.as_vol(ttd_padded)
        N, c_m, D, H, W = mask.shape
        c_i = cbv_padded.shape[1]
        out = ops.new_vol(N, c_m + 2 * c_i, D, H, W, mask.device)
        ops.crop_into(mask, out, 0, (0, 0, 0))
        ops.crop_into(cbv_padded, out, c_m, tuple(self._padding))
        ops.crop_into(ttd_padded, out, c_m + c_i, tuple(self._padding))
        return out

    def forward(self, dto: CaeDto):
        step = self._get_step(dto)
        given = dto.given_variables
        cbv, ttd = given.inputs.core, given.inputs.penu
        if dto.flag == CaeDtoUtil.FLAG_GTRUTH or dto.flag == CaeDtoUtil.FLAG_DEFAULT:
            lat = dto.latents.gtruth
            lat.core, lat.penu, lat.lesion = self._forward_group(
                [self._stack(given.gtruth.core, cbv, ttd), self._stack(given.gtruth.penu, cbv, ttd),
                 self._stack(given.gtruth.lesion, cbv, ttd)])
            lat.interpolation = self._interpolate(lat.core, lat.penu, step)
        return dto


class Dec3D(CaeBase):
    def __init__(self, size_input_xy, size_input_z, channels, n_ch_global, alpha):
        super().__init__(size_input_xy, size_input_z, channels, n_ch_global, alpha, inner_xy=10, inner_z=3)
        c1, c2, c4, c8, cf = self.n_ch_origin, self.n_ch_down2x, self.n_ch_down4x, self.n_ch_down8x, self.n_ch_fc
        grow = (1, 2, 2)
        # (kind, cin, cout, k, stride, padding), Cae3D.py:177-218
        table = [('T', cf, c8, 3, 1, 0), ('T', c8, c4, 3, 2, 0),
                 ('C', c4, c4, 3, 1, grow), ('C', c4, c2, 3, 1, grow),
                 ('T', c2, c2, 2, 2, 0),
                 ('C', c2, c2, 3, 1, grow), ('C', c2, c1, 3, 1, grow),
                 ('T', c1, c1, 2, 2, 0),
                 ('C', c1, c1, 3, 1, grow), ('C', c1, c1, 3, 1, grow),
                 ('C', c1, c1, 1, 1, 0)]
        mods = []
        for kind, cin, cout, k, stride, padding in table:
            mods += _unit(kind, cin, cout, k, stride, padding, nn.ELU(alpha, True))
        mods += _unit('C', c1, self.n_classes, 1, 1, 0, nn.Sigmoid())
        self.decoder = nn.Sequential(*mods)

    def _forward_single(self, input_latent):
        if input_latent is None:
            return None
        _require_cuda(input_latent, 'Dec3D')
        return engine.run_sequential(_plans.get(self, 'decoder'), input_latent)

    def _forward_group(self, latents):
        return _run_grouped(self, 'decoder', latents, 'Dec3D')

    def forward(self, dto: CaeDto):
        lat, rec = dto.latents, dto.reconstructions
        if dto.flag == CaeDtoUtil.FLAG_GTRUTH or dto.flag == CaeDtoUtil.FLAG_DEFAULT:
            assert rec.gtruth._is_empty()  # do not overwrite earlier results
            rec.gtruth.core, rec.gtruth.penu, rec.gtruth.lesion, rec.gtruth.interpolation = self._forward_group(
                [lat.gtruth.core, lat.gtruth.penu, lat.gtruth.lesion, lat.gtruth.interpolation])
        if dto.flag == CaeDtoUtil.FLAG_INPUTS or dto.flag == CaeDtoUtil.FLAG_DEFAULT:
            assert rec.inputs._is_empty()
            rec.inputs.core, rec.inputs.penu, rec.inputs.interpolation = self._forward_group(
                [lat.inputs.core, lat.inputs.penu, lat.inputs.interpolation])
        return dto


class Cae3D(nn.Module):
    def __init__(self, enc: Enc3D, dec: Dec3D):
        super().__init__()
        self.enc = enc
        self.dec = dec

    def forward(self, dto: CaeDto):
        dto = self.enc(dto)
        dto = self.dec(dto)
        return dto

    def freeze(self, freeze: bool):
        self.enc.freeze(freeze)
        self.dec.freeze(freeze)


class Cae3DCtp(Cae3D):
    def __init__(self, enc: Enc3DCtp, dec: Dec3D):
        Cae3D.__init__(self, enc, dec)
