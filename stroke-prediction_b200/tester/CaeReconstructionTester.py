"""CAE shape-reconstruction tester (API of the reference's tester/CaeReconstructionTester.py:12-67).

``batch_metrics_step`` evaluates the three (reconstruction, ground truth) pairs on the device with one device-to-host
read (the reference makes three host round trips through medpy).  ``save_inference`` writes the x2 in-plane zoomed NIfTI
volumes of the reference (:28-44); that is file I/O outside the hot-path scope and needs ``nibabel`` plus the reference
affines the reference reads from its private data share — it is skipped with a one-time note when either is missing.
"""
from ..common import data, metrics
from ..common.dto import MetricMeasuresDto as MetricMeasuresDtoInit
from ..common.dto.CaeDto import CaeDto
from ..common.dto.MetricMeasuresDto import MetricMeasuresDto
from ..common.inference.CaeInference import CaeInference
from .Tester import Tester


class CaeReconstructionTester(Tester, CaeInference):
    _noted_no_nifti = False

    def __init__(self, dataloader, path_model, path_outputs_base='/tmp/', normalization_hours_penumbra=10):
        Tester.__init__(self, dataloader, path_model, path_outputs_base=path_outputs_base)
        CaeInference.__init__(self, self._model, normalization_hours_penumbra)

    def batch_metrics_step(self, dto: CaeDto):
        batch_metrics = MetricMeasuresDtoInit.init_dto()
        batch_metrics.lesion, batch_metrics.core, batch_metrics.penu = metrics.binary_measures_many([
            (dto.reconstructions.gtruth.interpolation, dto.given_variables.gtruth.lesion),
            (dto.reconstructions.gtruth.core, dto.given_variables.gtruth.core),
            (dto.reconstructions.gtruth.penu, dto.given_variables.gtruth.penu)])
        return batch_metrics

    def save_inference(self, dto: CaeDto, batch: dict, suffix=''):
        try:
            import nibabel  # noqa: F401
        except ImportError:
            if not CaeReconstructionTester._noted_no_nifti:
                print('(save_inference skipped: nibabel is not installed)')
                CaeReconstructionTester._noted_no_nifti = True
            return
        raise NotImplementedError('NIfTI export needs the reference affines of the private data share '
                                  '(CaeReconstructionTester.py:31-44); override save_inference for your data layout')

    def print_inference(self, batch: dict, batch_metrics: MetricMeasuresDto, dto: CaeDto, note=''):
        output = 'Case Id={}\ttA-tO={:.3f}\ttR-tA={:.3f}\tnormalized_time_to_treatment={:.3f}\t-->\
                  \tDC={:.3f}\tHD={:.3f}\tASSD={:.3f}\tDC Core={:.3f}\tDC Penumbra={:.3f}\t\
                  Precision={:.3}\tRecall/Sensitivity={:.3}\tSpecificity={:.3}\tDistToCornerPRC={:.3}\t{}'
        print(output.format(int(batch[data.KEY_CASE_ID]),
                            float(batch[data.KEY_GLOBAL][:, 0, :, :, :]),
                            float(batch[data.KEY_GLOBAL][:, 1, :, :, :]),
                            float(dto.given_variables.time_to_treatment),
                            batch_metrics.lesion.dc,
                            batch_metrics.lesion.hd,
                            batch_metrics.lesion.assd,
                            batch_metrics.core.dc,
                            batch_metrics.penu.dc,
                            batch_metrics.lesion.precision,
                            batch_metrics.lesion.sensitivity,
                            batch_metrics.lesion.specificity,
                            batch_metrics.lesion.prc_euclidean_distance,
                            note))
