"""Loss of the hot path: whole-batch soft Dice (API of the reference's common/metrics.py:8-28).

The medpy-based evaluation metrics of metrics.py:31-62 are a "next" row and not part of this module.
"""
from torch.nn.modules.loss import _Loss as LossModule

from .. import functions


class BatchDiceLoss(LossModule):
    def __init__(self, label_weights, epsilon=0.0000001, dim=1):
        super(BatchDiceLoss, self).__init__()
        self._epsilon = epsilon
        self._dim = dim
        self._label_weights = label_weights

    def forward(self, outputs, targets):
        assert targets.shape[self._dim] == len(self._label_weights), \
            'Ground truth number of labels does not match with label weight vector'
        if not outputs.is_cuda:
            raise RuntimeError("BatchDiceLoss: CUDA tensors only — there is no CPU path")
        loss = None
        for label, weight in enumerate(self._label_weights):
            o = outputs.narrow(self._dim, label, 1)
            t = targets.narrow(self._dim, label, 1)
            assert o.numel() == t.numel()
            # dice_term returns 1 - w * num/den; the reference sums w * num/den over labels and subtracts from 1 once
            term = functions.dice_term(o, t, weight, self._epsilon)
            loss = term if loss is None else loss + (term - 1.0)
        return loss
