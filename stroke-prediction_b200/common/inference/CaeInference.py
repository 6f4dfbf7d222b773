"""Batch dict -> CaeDto -> model(dto) (API of the reference's common/inference/CaeInference.py:10-69).

Host-side differences from the reference, none of which change values: the whole label tensor crosses PCIe in one
copy and is split into the three dense single-channel masks on the device (the reference slices on the host and
issues three copies, CaeInference.py:49-56); the B x 1 reshapes use ``reshape`` so batch size 1 works on current
torch.  The flag defect D1 (``dto.mode`` written, ``dto.flag`` read) is kept for this class: the flag
stays 'default' and the empty ``inputs`` branch propagates ``None`` exactly as in the reference.
"""
import torch

from ... import ops
from .. import data
from ..dto import CaeDto as CaeDtoUtil
from ..dto.CaeDto import CaeDto
from ..model.Cae3D import Cae3D
from .Inference import Inference


class CaeInference(Inference):
    def __init__(self, model: Cae3D, normalization_hours_penumbra=10):
        Inference.__init__(self, model)
        self._normalization_hours_penumbra = normalization_hours_penumbra

    def _get_normalization(self, batch):
        """(normalisation hours) - (onset -> admission hours), fp32, shape B x 1 (CaeInference.py:18-22)."""
        to_to_ta = batch[data.KEY_GLOBAL][:, 0, :, :, :].float().reshape(-1, 1)
        return torch.ones(to_to_ta.size()[0], 1) * self._normalization_hours_penumbra - to_to_ta.cpu()

    def get_time_to_treatment(self, batch, global_variables, step):
        normalization = self._get_normalization(batch)
        if step is None:
            ta_to_tr = batch[data.KEY_GLOBAL][:, 1, :, :, :].float().reshape(-1, 1).cpu()
            time_to_treatment = ta_to_tr / normalization
        else:
            time_to_treatment = (step * torch.ones(global_variables.size()[0], 1)) / normalization
        return time_to_treatment.unsqueeze(2).unsqueeze(3).unsqueeze(4)

    def init_clinical_variables(self, batch: dict, step):
        globals_incl_time = batch[data.KEY_GLOBAL].float()
        n = globals_incl_time.size()[0]
        time_to_treatment = self.get_time_to_treatment(batch, globals_incl_time, step)

        if self.is_cuda:
            if time_to_treatment is not None:
                time_to_treatment = self._to_device(time_to_treatment)
            globals_incl_time = self._to_device(globals_incl_time)
            type_core = torch.zeros(n, 1, 1, 1, 1, device=self.device)        # constants: created where they are used
            type_penumbra = torch.ones(n, 1, 1, 1, 1, device=self.device)
        else:
            type_core = torch.zeros(n, 1, 1, 1, 1)
            type_penumbra = torch.ones(n, 1, 1, 1, 1)

        return CaeDtoUtil.init_dto(globals_incl_time, time_to_treatment, type_core, type_penumbra,
                                   None, None, None, None, None)

    def init_gtruth_segm_variables(self, batch: dict, dto: CaeDto):
        labels = batch[data.KEY_LABELS]
        if not self.is_cuda:
            raise RuntimeError("model is not on a CUDA device — stroke_prediction_b200 has no CPU path")
        if labels.dtype != torch.float32:
            labels = labels.float()
        B, C, D, H, W = labels.shape
        # The three masks land in ONE [3B, 1, D, H, W] buffer, channel-major: the encoder recognises the adjacent slices and
        # runs core / penumbra / lesion as one stacked pass.  The host batch crosses PCIe as ONE contiguous (pinned: asynchronous)
        # transfer; the (sample, channel) -> (channel, sample) reorder is a device-side strided copy.  (Copying the three channel
        # slices separately makes torch gather each of them into a pageable temporary on the host first.)
        if not labels.is_contiguous():
            labels = labels.contiguous()
        dev = labels.to(self.device, non_blocking=labels.is_pinned()) if not labels.is_cuda else labels
        buf = torch.empty((3 * B, 1, D, H, W), device=self.device, dtype=torch.float32)
        buf.view(3, B, D * H * W).copy_(dev.view(B, C, D * H * W)[:, :3].transpose(0, 1))
        dto.given_variables.gtruth.core = buf[0:B]
        dto.given_variables.gtruth.penu = buf[B:2 * B]
        dto.given_variables.gtruth.lesion = buf[2 * B:3 * B]
        return dto

    def infer(self, dto: CaeDto):
        return self._model(dto)

    def inference_step(self, batch: dict, step=None):
        dto = self.init_clinical_variables(batch, step)
        dto.mode = CaeDtoUtil.FLAG_GTRUTH   # sic (reference CaeInference.py:67): models read dto.flag
        dto = self.init_gtruth_segm_variables(batch, dto)
        return self.infer(dto)
