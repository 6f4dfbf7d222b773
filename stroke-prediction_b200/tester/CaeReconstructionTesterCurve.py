"""Time-curve analysis of the CAE shape space (API of the reference's tester/CaeReconstructionTesterCurve.py:5-42;
SURVEY §8f n4).

The reference re-runs the whole model — three encoder and four decoder passes — for every one of the ~32 time points it
evaluates per case (:18-42).  Only the interpolated latent depends on the time point, so ``infer_curve`` encodes the case
ONCE, decodes core / penumbra / lesion ONCE, forms all interpolated latents, decodes them as ONE stacked pass (the tester
runs in ``eval()`` mode: BatchNorm uses running statistics, so stacking along the batch is exact) and evaluates every time
point's metrics with one device-to-host read.  ``infer_batch(batch, step)`` keeps the reference's one-point semantics.
"""
import torch

from .. import functions
from ..common import data, metrics
from ..common.dto import CaeDto as CaeDtoUtil
from ..common.dto import MetricMeasuresDto as MetricMeasuresDtoInit
from ..common.dto.Dto import Dto
from .CaeReconstructionTester import CaeReconstructionTester


class CaeReconstructionTesterCurve(CaeReconstructionTester):
    def __init__(self, dataloader, path_model, path_outputs_base='/tmp/', normalization_hours_penumbra=10,
                 ta_to_tr_fixed_hours=range(11), ta_to_tr_relative_steps=[0, 0.25, 0.5, 0.75, 1, 1.25, 1.5, 1.75, 2]):
        CaeReconstructionTester.__init__(self, dataloader, path_model, path_outputs_base=path_outputs_base,
                                         normalization_hours_penumbra=normalization_hours_penumbra)
        self._steps_fixed = ta_to_tr_fixed_hours
        self._steps_relative = ta_to_tr_relative_steps

    def infer_batch(self, batch: dict, step: float):
        with torch.no_grad():
            dto = self.inference_step(batch, step)
        batch_metrics = self.batch_metrics_step(dto)
        return batch_metrics, dto

    # ------------------------------------------------------------------------------------------ cached sweep
    def infer_curve(self, batch: dict, steps, chunk=64):
        """[(batch_metrics, dto)] for every entry of `steps` (hours tA->tR; ``None`` = the case's own), equal to
        ``[self.infer_batch(batch, s) for s in steps]`` at a fraction of the work."""
        steps = list(steps)
        if self._model.training:
            # training-mode BatchNorm would couple the stacked time points through their batch statistics
            return [self.infer_batch(batch, s) for s in steps]
        with torch.no_grad():
            base = self.inference_step(batch, steps[0] if steps else None)      # encoder x3, decoder x4, once
            lat, rec, given = base.latents.gtruth, base.reconstructions.gtruth, base.given_variables
            times = []
            for s in steps:
                t = self.get_time_to_treatment(batch, given.globals, s)
                times.append(self._to_device(t))
            recs = [rec.interpolation]
            lats = [lat.interpolation]
            for lo in range(1, len(steps), chunk):
                zs = [functions.latent_interp(lat.core, lat.penu, t) for t in times[lo:lo + chunk]]
                lats += zs
                recs += self._model.dec._forward_group(zs)                     # ONE stacked decoder pass per chunk
            pairs = [(r, given.gtruth.lesion) for r in recs]
            pairs += [(rec.core, given.gtruth.core), (rec.penu, given.gtruth.penu)]
            measures = metrics.binary_measures_many(pairs)                      # one device-to-host read
        core_m, penu_m = measures[-2], measures[-1]
        out = []
        for i, s in enumerate(steps):
            dto = CaeDtoUtil.init_dto(given.globals, times[i], given.scalar_types.core, given.scalar_types.penu, None, None,
                                      given.gtruth.core, given.gtruth.penu, given.gtruth.lesion)
            dto.mode = getattr(base, 'mode', None)
            dto.latents.gtruth = Dto(core=lat.core, penu=lat.penu, lesion=lat.lesion, interpolation=lats[i])
            dto.reconstructions.gtruth = Dto(core=rec.core, penu=rec.penu, lesion=rec.lesion, interpolation=recs[i])
            m = MetricMeasuresDtoInit.init_dto()
            m.lesion, m.core, m.penu = measures[i], core_m, penu_m
            out.append((m, dto))
        return out

    def curve_steps(self, batch: dict):
        """The reference's evaluation schedule (:22-42) as one list of (hours, note)."""
        ta_to_tr = float(batch[data.KEY_GLOBAL][:, 1, :, :, :])
        to_to_ta = float(batch[data.KEY_GLOBAL][:, 0, :, :, :])
        tr_to_penu = self._normalization_hours_penumbra - to_to_ta
        sched = [(None, '')]
        sched += [(step, 'ta_to_tr fixed=' + str(step)) for step in self._steps_fixed]
        sched += [(step * ta_to_tr, 'ta_to_tr ratio=' + str(step) + '\t(' + str(step * ta_to_tr) + ')')
                  for step in self._steps_relative]
        sched += [(step * tr_to_penu, 'tr_to_penumbra=' + str(step) + '\t(' + str(step * tr_to_penu) + ')')
                  for step in [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0]]
        return sched

    def run_inference(self):
        for batch in self._dataloader:
            sched = self.curve_steps(batch)
            results = self.infer_curve(batch, [s for s, _ in sched])
            for k, ((batch_metrics, dto), (_, note)) in enumerate(zip(results, sched)):
                if k == 0:      # 1) ground-truth tA-->tR: printed without a note and saved (:21-24)
                    self.print_inference(batch, batch_metrics, dto)
                    self.save_inference(dto, batch)
                else:
                    self.print_inference(batch, batch_metrics, dto, note)
