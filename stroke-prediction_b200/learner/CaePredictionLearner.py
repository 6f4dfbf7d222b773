"""Shape prediction from U-Net segmentations (API of the reference's learner/CaePredictionLearner.py)."""
import torch

from .. import functions
from ..common import metrics
from ..common.dto import MetricMeasuresDto as MetricMeasuresDtoInit
from ..common.dto.CaeDto import CaeDto
from ..common.inference.CaeEncInference import CaeEncInference
from .Learner import Learner


class CaePredictionLearner(Learner, CaeEncInference):
    FN_VIS_BASE = '_cae2_'
    FNB_MARKS = '_cae2'
    N_EPOCHS_ADAPT_BETA1 = 4

    def __init__(self, dataloader_training, dataloader_validation, cae_model, enc_model, optimizer, scheduler, n_epochs,
                 path_previous_base, path_outputs_base, criterion, normalization_hours_penumbra=10):
        Learner.__init__(self, dataloader_training, dataloader_validation, cae_model, optimizer, scheduler, n_epochs,
                         path_previous_base, path_outputs_base)
        CaeEncInference.__init__(self, cae_model, enc_model, normalization_hours_penumbra)
        self._model.freeze(True)
        self._criterion = criterion

    def load_model(self, cuda=True):
        Learner.load_model(self, self.is_cuda)
        enc = torch.load(self.path('load', self.FNB_MODEL, '_enc'), weights_only=False)
        self._new_enc = self.adopt_checkpoint(getattr(self, '_new_enc', None), enc, cuda)

    def save_model(self, suffix=''):
        Learner.save_model(self, suffix)
        torch.save(self.cpu_snapshot(self._new_enc), self.path('save', self.FNB_MODEL, '_enc' + suffix))

    def adapt_betas(self, epoch):
        pass

    def loss_step(self, dto: CaeDto, epoch):
        rec_in, lat_in, lat_gt = dto.reconstructions.inputs, dto.latents.inputs, dto.latents.gtruth
        loss = functions.hinge_mean(rec_in.penu, rec_in.interpolation)
        loss = loss + functions.hinge_mean(rec_in.penu, rec_in.core)
        loss = loss + self._criterion(rec_in.interpolation, dto.given_variables.gtruth.lesion)
        loss = loss + functions.l1_mean(lat_gt.interpolation, lat_in.interpolation)
        loss = loss + functions.l1_mean(lat_gt.core, lat_in.core)
        loss = loss + functions.l1_mean(lat_gt.penu, lat_in.penu)
        return loss / 6

    def batch_metrics_step(self, dto: CaeDto, epoch):
        """CaePredictionLearner.py:59-67 (metrics on the frozen CAE's ground-truth branch), one D2H for all."""
        batch_metrics = MetricMeasuresDtoInit.init_dto()
        batch_metrics.lesion, batch_metrics.core, batch_metrics.penu = metrics.binary_measures_many([
            (dto.reconstructions.gtruth.interpolation, dto.given_variables.gtruth.lesion),
            (dto.reconstructions.gtruth.core, dto.given_variables.gtruth.core),
            (dto.reconstructions.gtruth.penu, dto.given_variables.gtruth.penu)])
        return batch_metrics

    def print_epoch(self, epoch, phase, epoch_metrics):
        print('\nEpoch {}/{} {} loss: {:.3}'.format(epoch + 1, self._n_epochs, phase, epoch_metrics.loss), end=' ')
