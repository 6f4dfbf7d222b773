"""Shape-prediction inference: new encoder on U-Net segmentations + frozen CAE on ground truth
(API of the reference's common/inference/CaeEncInference.py:9-42).

The reference writes ``dto.mode`` where the models read ``dto.flag`` (defect D1), so its second pass
trips the overwrite assert at Cae3D.py:112.  This drop-in implements the *intended* two-pass semantics by setting
``dto.flag`` (and still mirrors the value into ``dto.mode``): pass 1 (flag = inputs) encodes the two segmentations
with the new encoder and decodes them with the frozen decoder; pass 2 (flag = gtruth) runs the frozen CAE on the
ground-truth masks.  ``compat_reference_flag_bug=True`` restores the reference's literal behaviour.
"""
from ... import ops
from .. import data
from ..dto import CaeDto as CaeDtoUtil
from ..dto.CaeDto import CaeDto
from ..model.Cae3D import Cae3D, Enc3D
from .CaeInference import CaeInference


class CaeEncInference(CaeInference):
    def __init__(self, model: Cae3D, new_enc: Enc3D, normalization_hours_penumbra=10,
                 compat_reference_flag_bug=False):
        CaeInference.__init__(self, model, normalization_hours_penumbra)
        self._new_enc = new_enc
        self._compat_flag_bug = compat_reference_flag_bug

    def infer(self, dto: CaeDto):
        pass

    def init_unet_segm_variables(self, batch: dict, dto: CaeDto):
        images = ops.as_vol(self._to_device(batch[data.KEY_IMAGES]))
        dto.given_variables.inputs.core = ops.extract_channel(images, 0)
        dto.given_variables.inputs.penu = ops.extract_channel(images, 1)
        return dto

    def _set_pass(self, dto, flag):
        dto.mode = flag
        if not self._compat_flag_bug:
            dto.flag = flag

    def inference_step(self, batch: dict, step=None):
        dto = self.init_clinical_variables(batch, step)

        self._set_pass(dto, CaeDtoUtil.FLAG_INPUTS)
        dto = self.init_unet_segm_variables(batch, dto)
        dto = self._new_enc(dto)
        dto = self._model.dec(dto)

        self._set_pass(dto, CaeDtoUtil.FLAG_GTRUTH)
        dto = self.init_gtruth_segm_variables(batch, dto)
        dto = self._model(dto)

        return dto
