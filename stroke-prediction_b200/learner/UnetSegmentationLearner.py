"""U-Net segmentation training (API of the reference's learner/UnetSegmentationLearner.py; its constructor's missing
``self`` in the base-class call — defect D2, SURVEY App. B — is not reproduced)."""
import numpy

from ..common import metrics
from ..common.dto import MetricMeasuresDto as MetricMeasuresDtoInit
from ..common.dto.UnetDto import UnetDto
from ..common.inference.UnetInference import UnetInference
from .Learner import Learner


class UnetSegmentationLearner(Learner, UnetInference):
    FNB_MARKS = '_unet'

    def __init__(self, dataloader_training, dataloader_validation, unet_model, optimizer, scheduler, n_epochs,
                 criterion, path_previous_base=None, path_outputs_base='/tmp/unet-segmentation'):
        Learner.__init__(self, dataloader_training, dataloader_validation, unet_model, optimizer, scheduler, n_epochs,
                         path_previous_base, path_outputs_base)
        self._criterion = criterion

    def loss_step(self, dto: UnetDto, epoch):
        loss = self._criterion(dto.outputs.core, dto.given_variables.core)
        loss = loss + self._criterion(dto.outputs.penu, dto.given_variables.penu)
        return loss / 2

    def batch_metrics_step(self, dto: UnetDto, epoch):
        """UnetSegmentationLearner.py:30-36: thresholded overlap of both outputs, one D2H."""
        batch_metrics = MetricMeasuresDtoInit.init_dto()
        batch_metrics.core, batch_metrics.penu = metrics.binary_measures_many([
            (dto.outputs.core, dto.given_variables.core), (dto.outputs.penu, dto.given_variables.penu)])
        return batch_metrics

    def get_start_epoch(self):
        return len(self._metric_dtos['training'])

    def get_start_min_loss(self):
        if self._metric_dtos['validate']:
            return min(dto.loss for dto in self._metric_dtos['validate'])
        return numpy.inf

    def print_epoch(self, epoch, phase, epoch_metrics):
        print('\nEpoch {}/{} {} loss: {:.3}'.format(epoch + 1, self._n_epochs, phase, epoch_metrics.loss), end=' ')
