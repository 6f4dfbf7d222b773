"""Metric accumulators (API of the reference's common/dto/MetricMeasuresDto.py)."""
import numpy

from .Dto import Dto


class MeasuresDto(Dto):
    def add(self, other):
        if not isinstance(other, type(self)):
            raise Exception('A' + str(type(self)) + 'must be added')
        for name, value in other:
            mine = self.__dict__[name]
            if mine is None:
                self.__dict__[name] = value
            elif isinstance(value, MeasuresDto):
                mine.add(value)
            else:
                self.__dict__[name] = mine + value

    def div(self, divisor):
        for name, value in self:
            if value is None:
                continue
            if isinstance(value, MeasuresDto):
                value.div(divisor)
            elif value != numpy.inf:
                self.__dict__[name] = value / divisor


class BinaryMeasuresDto(MeasuresDto):
    def __init__(self, dc, hd, assd, precision, sensitivity, specificity):
        super().__init__()
        self.dc = dc
        self.hd = hd
        self.assd = assd
        self.precision = precision
        self.sensitivity = sensitivity
        self.specificity = specificity

    @property
    def prc_euclidean_distance(self):
        return numpy.sqrt((1 - self.precision) ** 2 + (1 - self.sensitivity) ** 2)


class MetricMeasuresDto(MeasuresDto):
    def __init__(self, loss, core: BinaryMeasuresDto, penu: BinaryMeasuresDto, lesion: BinaryMeasuresDto):
        super().__init__()
        self.loss = loss
        self.core = core
        self.penu = penu
        self.lesion = lesion


def init_dto(loss=None, core_dc=None, core_hd=None, core_assd=None, penu_dc=None, penu_hd=None, penu_assd=None,
             lesion_dc=None, lesion_hd=None, lesion_assd=None, lesion_precision=None, lesion_sensitivity=None,
             lesion_specificity=None):
    return MetricMeasuresDto(loss,
                             BinaryMeasuresDto(core_dc, core_hd, core_assd, None, None, None),
                             BinaryMeasuresDto(penu_dc, penu_hd, penu_assd, None, None, None),
                             BinaryMeasuresDto(lesion_dc, lesion_hd, lesion_assd, lesion_precision,
                                               lesion_sensitivity, lesion_specificity))
