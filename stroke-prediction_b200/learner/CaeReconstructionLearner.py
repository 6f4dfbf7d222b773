"""CAE shape-space training (API of the reference's learner/CaeReconstructionLearner.py).

``loss_step`` (reference :52-70) = [2 monotonicity hinges + 3 whole-batch Dice terms + ramped latent L1] / (5 + f);
every term is a fused reduction kernel with a closed-form gradient kernel (functions.py).
"""
import numpy

from .. import functions
from ..common import metrics
from ..common.dto import MetricMeasuresDto as MetricMeasuresDtoInit
from ..common.dto.CaeDto import CaeDto
from ..common.inference.CaeInference import CaeInference
from .Learner import Learner


class CaeReconstructionLearner(Learner, CaeInference):
    FN_VIS_BASE = '_cae1_'
    FNB_MARKS = '_cae1'
    N_EPOCHS_ADAPT_BETA1 = 4

    def __init__(self, dataloader_training, dataloader_validation, cae_model, optimizer, scheduler, n_epochs,
                 path_previous_base, path_outputs_base, criterion, normalization_hours_penumbra=10):
        Learner.__init__(self, dataloader_training, dataloader_validation, cae_model, optimizer, scheduler, n_epochs,
                         path_previous_base, path_outputs_base)
        CaeInference.__init__(self, cae_model, normalization_hours_penumbra)
        self._criterion = criterion

    def adapt_betas(self, epoch):
        """beta1 = 0.5, 0.6, 0.7, 0.8 for epochs 0..3, then the optimizer default (reference :28-40)."""
        betas = self._optimizer.defaults['betas']
        if epoch < self.N_EPOCHS_ADAPT_BETA1:
            betas = (betas[0] - 0.1 * (self.N_EPOCHS_ADAPT_BETA1 - epoch),) + tuple(betas[1:])
        elif epoch != self.N_EPOCHS_ADAPT_BETA1:
            return
        for param_group in self._optimizer.param_groups:
            param_group['betas'] = betas
        print('Momentum betas have been set to:', betas, end=' ')

    def get_start_epoch(self):
        return len(self._metric_dtos['training'])

    def get_start_min_loss(self):
        if self._metric_dtos['validate']:
            return min(dto.loss for dto in self._metric_dtos['validate'])
        return numpy.inf

    def loss_step(self, dto: CaeDto, epoch):
        factor = min(0.04 * max(0, epoch - 25), 1)
        rec, given, lat = dto.reconstructions.gtruth, dto.given_variables.gtruth, dto.latents.gtruth
        loss = functions.hinge_mean(rec.penu, rec.interpolation)
        loss = loss + functions.hinge_mean(rec.penu, rec.core)
        loss = loss + self._criterion(rec.core, given.core)
        loss = loss + self._criterion(rec.penu, given.penu)
        loss = loss + self._criterion(rec.lesion, given.lesion)
        loss = loss + factor * functions.l1_mean(lat.interpolation, lat.lesion)
        return loss / (5 + factor)

    def batch_metrics_step(self, dto: CaeDto, epoch):
        """CaeReconstructionLearner.py:72-80: thresholded overlap of the three reconstructions, one D2H for all."""
        batch_metrics = MetricMeasuresDtoInit.init_dto()
        batch_metrics.lesion, batch_metrics.core, batch_metrics.penu = metrics.binary_measures_many([
            (dto.reconstructions.gtruth.interpolation, dto.given_variables.gtruth.lesion),
            (dto.reconstructions.gtruth.core, dto.given_variables.gtruth.core),
            (dto.reconstructions.gtruth.penu, dto.given_variables.gtruth.penu)])
        return batch_metrics

    def print_epoch(self, epoch, phase, epoch_metrics):
        print('\nEpoch {}/{} {} loss: {:.3}'.format(epoch + 1, self._n_epochs, phase, epoch_metrics.loss), end=' ')
