"""U-Net segmentation tester (API of the reference's tester/UnetSegmentationTester.py:12-52); metrics on the device."""
from ..common import data, metrics
from ..common.dto import MetricMeasuresDto as MetricMeasuresDtoInit
from ..common.dto.MetricMeasuresDto import MetricMeasuresDto
from ..common.dto.UnetDto import UnetDto
from ..common.inference.UnetInference import UnetInference
from .Tester import Tester


class UnetSegmentationTester(Tester, UnetInference):
    def __init__(self, dataloader, path_model, path_outputs_base='/tmp/', padding=None):
        Tester.__init__(self, dataloader, path_model, path_outputs_base=path_outputs_base)
        self._pad = padding

    def batch_metrics_step(self, dto: UnetDto):
        batch_metrics = MetricMeasuresDtoInit.init_dto()
        batch_metrics.core, batch_metrics.penu = metrics.binary_measures_many([
            (dto.outputs.core, dto.given_variables.core), (dto.outputs.penu, dto.given_variables.penu)])
        return batch_metrics

    def save_inference(self, dto: UnetDto, batch: dict, suffix=''):
        pass   # NIfTI export (reference :35-44): file I/O on the private data share, outside the hot-path scope

    def print_inference(self, batch: dict, batch_metrics: MetricMeasuresDto, dto: UnetDto = None):
        output = 'Case Id {}:\t DC Core:{:.3},\tDC Penumbra:{:.3}'
        print(output.format(int(batch[data.KEY_CASE_ID]), batch_metrics.core.dc, batch_metrics.penu.dc))
