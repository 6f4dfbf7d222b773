"""Batch dict -> UnetDto -> model(dto) (API of the reference's common/inference/UnetInference.py:10-26)."""
from ... import ops
from .. import data
from ..dto import UnetDto as UnetDtoUtil
from ..model.Unet3D import Unet3D
from .Inference import Inference


class UnetInference(Inference):
    def __init__(self, model: Unet3D):
        Inference.__init__(self, model)

    def inference_step(self, batch):
        input_modalities = self._to_device(batch[data.KEY_IMAGES])
        labels = ops.as_vol(self._to_device(batch[data.KEY_LABELS]))
        core_gt = ops.extract_channel(labels, 0)
        penu_gt = ops.extract_channel(labels, 1)
        dto = UnetDtoUtil.init_dto(input_modalities, core_gt, penu_gt)
        return self._model(dto)
