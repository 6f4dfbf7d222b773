"""U-Net data transfer object (API of the reference's common/dto/UnetDto.py)."""
from .Dto import Dto


class UnetDto(Dto):
    def __init__(self, given_variables: Dto, outputs: Dto):
        super().__init__()
        self.given_variables = given_variables
        self.outputs = outputs


def init_dto(input_modalities, gtruth_core=None, gtruth_penumbra=None, gtruth_lesion=None):
    given = Dto(input_modalities=input_modalities, core=gtruth_core, penu=gtruth_penumbra, lesion=gtruth_lesion)
    return UnetDto(given_variables=given, outputs=Dto(core=None, penu=None, lesion=None))
