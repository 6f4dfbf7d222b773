"""Loss and per-batch evaluation metrics of the hot path (API of the reference's common/metrics.py).

* ``BatchDiceLoss`` — whole-batch soft Dice (metrics.py:8-28).
* ``binary_measures_torch`` / ``binary_measures_many`` — the thresholded metrics every train / validation batch reports
  (metrics.py:31-62).  The reference copies both volumes to the host and calls the un-vendored
  ``MedPy==0.3.0`` (``requirements.txt:2``); here everything is computed on the device and ALL pairs of a batch come back
  in ONE device-to-host read:
  - dc / precision / sensitivity / specificity from the confusion counts of one reduction kernel per pair (medpy's
    published definitions: dc = 2 TP / (|A| + |B|), precision = TP / (TP + FP), sensitivity = TP / (TP + FN),
    specificity = TN / (TN + FP); each 0.0 when its denominator is 0);
  - hd / assd (``mpm.hd`` / ``mpm.assd``, metrics.py:43-45) from exact Euclidean distance transforms on the device
    (``sp_surface_distances``): like medpy they treat the WHOLE array handed in — the reference hands in the
    B x 1 x D x H x W batch — as one lattice with unit spacing, and stay ``numpy.inf`` unless both masks have voxels.
  ``SURFACE_DISTANCES = False`` skips the distance transforms (hd / assd stay ``inf``).
"""
import numpy
import torch
from torch.nn.modules.loss import _Loss as LossModule

from .. import functions, ops
from .dto.MetricMeasuresDto import BinaryMeasuresDto

SURFACE_DISTANCES = True

# While a `deferred()` context is active, binary_measures_many launches its kernels, returns BinaryMeasuresDto placeholders
# and leaves the device-to-host read to `finish()`, which fills the placeholders in place.  Learner.train_batch uses it to
# run the per-batch metrics (they only read the forward outputs) on a side stream while the backward pass and the
# optimizer step run on the main stream (Learner.py:116-130 computes them after the step; same values).
_PENDING = None


class deferred(object):
    def __enter__(self):
        global _PENDING
        self._prev, self.pending = _PENDING, []
        _PENDING = self.pending
        return self

    def __exit__(self, *exc):
        global _PENDING
        _PENDING = self._prev
        return False

    def finish(self):
        for out, dtos, with_sd in self.pending:
            for dto, row in zip(dtos, out.cpu().tolist()):
                new = measures_from_counts(*row[0:4], hd=row[4], assd=row[5]) if with_sd else measures_from_counts(*row[0:4])
                dto.__dict__.update(new.__dict__)
        self.pending = []


def measures_from_counts(tp, fp, fn, tn, hd=numpy.inf, assd=numpy.inf):
    """medpy.metric.binary dc / precision / sensitivity / specificity from the confusion counts."""
    tp, fp, fn, tn = float(tp), float(fp), float(fn), float(tn)
    size_r, size_t = tp + fp, tp + fn
    dc = 2.0 * tp / (size_r + size_t) if (size_r + size_t) > 0 else 0.0
    precision = tp / (tp + fp) if (tp + fp) > 0 else 0.0
    sensitivity = tp / (tp + fn) if (tp + fn) > 0 else 0.0
    specificity = tn / (tn + fp) if (tn + fp) > 0 else 0.0
    return BinaryMeasuresDto(dc, hd, assd, precision, sensitivity, specificity)


def binary_measures_many(pairs, binary_threshold=0.5, surface_distances=None):
    """[(result, target), ...] (CUDA tensors) -> [BinaryMeasuresDto, ...] with ONE device-to-host read for all pairs."""
    pairs = list(pairs)
    if not pairs:
        return []
    if surface_distances is None:
        surface_distances = SURFACE_DISTANCES
    dev = pairs[0][0].device
    out = torch.empty((len(pairs), 12), device=dev, dtype=torch.float64)
    for i, (r, t) in enumerate(pairs):
        r, t = r.detach(), t.detach()
        ops.binary_counts(r, t, binary_threshold, out=out[i, 0:4])
        if surface_distances:
            ops.surface_distances(r, t, binary_threshold, out=out[i, 4:12])
    if _PENDING is not None:
        dtos = [BinaryMeasuresDto(0.0, numpy.inf, numpy.inf, 0.0, 0.0, 0.0) for _ in pairs]
        _PENDING.append((out, dtos, bool(surface_distances)))
        return dtos
    host = out.cpu().tolist()
    if surface_distances:
        return [measures_from_counts(*row[0:4], hd=row[4], assd=row[5]) for row in host]
    return [measures_from_counts(*row[0:4]) for row in host]


def binary_measures_torch(result, target, cuda=True, binary_threshold=0.5):
    """Signature of metrics.py:49; ``cuda`` is kept for source compatibility — there is no CPU path."""
    if not result.is_cuda:
        raise RuntimeError("binary_measures_torch: CUDA tensors only — there is no CPU path")
    return binary_measures_many([(result, target)], binary_threshold)[0]


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
