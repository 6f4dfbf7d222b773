"""Base class of everything that runs a model forward (API of the reference's common/inference/Inference.py)."""
from abc import abstractmethod


class Inference():
    IMSHOW_VMAX_CBV = 12
    IMSHOW_VMAX_TTD = 40
    FN_VIS_BASE = '_visual_'
    INFERENCE_INITALIZED = False

    @abstractmethod
    def __init__(self, model):
        if not self.INFERENCE_INITALIZED:
            self._model = model
            self.INFERENCE_INITALIZED = True

    @abstractmethod
    def inference_step(self, batch: dict):
        pass

    @property
    def is_cuda(self) -> bool:
        return next(self._model.parameters()).is_cuda

    @property
    def device(self):
        return next(self._model.parameters()).device

    def _to_device(self, t):
        """Host -> device copy of one batch tensor (pinned + asynchronous when it comes from pinned memory)."""
        if t.is_cuda:
            return t
        if not self.is_cuda:
            raise RuntimeError("model is not on a CUDA device — stroke_prediction_b200 has no CPU path")
        return t.to(self.device, non_blocking=t.is_pinned())
