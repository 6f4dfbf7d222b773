"""CAE data transfer object (API of the reference's common/dto/CaeDto.py)."""
from .Dto import Dto

FLAG_DEFAULT = 'default'
FLAG_GTRUTH = 'gtruth'
FLAG_INPUTS = 'inputs'


class CaeDto(Dto):
    def __init__(self, given_variables: Dto, latents: Dto, reconstructions: Dto):
        super().__init__()
        self.given_variables = given_variables
        self.latents = latents
        self.reconstructions = reconstructions
        self.flag = FLAG_DEFAULT


def _branches():
    return Dto(inputs=Dto(core=None, penu=None, interpolation=None),
               gtruth=Dto(core=None, penu=None, interpolation=None, lesion=None))


def init_dto(global_variables, time_to_treatment, type_core, type_penumbra, inputs_core, inputs_penu,
             gtruth_core, gtruth_penumbra, gtruth_lesion):
    """Same positional signature as the reference (CaeDto.py:18-46)."""
    given = Dto(globals=global_variables,
                time_to_treatment=time_to_treatment,
                scalar_types=Dto(core=type_core, penu=type_penumbra),
                inputs=Dto(core=inputs_core, penu=inputs_penu),
                gtruth=Dto(core=gtruth_core, penu=gtruth_penumbra, lesion=gtruth_lesion))
    return CaeDto(given_variables=given, latents=_branches(), reconstructions=_branches())
