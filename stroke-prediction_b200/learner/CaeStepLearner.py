"""Interpolation-step training on a frozen CAE (API of the reference's learner/CaeStepLearner.py)."""
import torch

from .. import functions
from ..common.dto.CaeDto import CaeDto
from .CaeReconstructionLearner import CaeReconstructionLearner


class CaeStepLearner(CaeReconstructionLearner):
    FN_VIS_BASE = '_cae1step_'
    FNB_MARKS = '_cae1step'
    N_EPOCHS_ADAPT_BETA1 = 4

    def loss_step(self, dto: CaeDto, epoch):
        rec = dto.reconstructions.gtruth
        loss = functions.hinge_mean(rec.penu, rec.interpolation)
        loss = loss + self._criterion(rec.interpolation, dto.given_variables.gtruth.lesion)
        return loss / 2

    def get_time_to_treatment(self, batch, global_variables, step):
        """None -> Enc3DStep predicts the step from the clinical globals (reference :23-29)."""
        if step is None:
            return None
        normalization = self._get_normalization(batch)
        t = (step * torch.ones(global_variables.size()[0], 1)) / normalization
        return t.unsqueeze(2).unsqueeze(3).unsqueeze(4)
