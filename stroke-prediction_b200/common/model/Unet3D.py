"""Drop-in 3-scale U-Net: constructor, attribute names, ``forward(dto) -> dto`` and state_dict layout of the
reference's common/model/Unet3D.py:14-84; the whole graph (5 double-conv blocks, 2 max-pools, 2 trilinear
upsamples, 2 crop+concat skips, 1x1 classifier) runs as one autograd node on libstroke_b200.so kernels.

``LargeUnet3D`` (Unet3D.py:87-148) cannot be constructed in the reference (wrong ``super()``), is never referenced,
and is intentionally not provided.
"""
import torch.nn as nn

from ... import engine
from ..dto.UnetDto import UnetDto


def crop(tensor_in, crop_as, dims=[]):
    """Centre-crop `tensor_in` to the extent of `crop_as` along `dims` (view; Unet3D.py:6-11)."""
    assert len(dims) > 0, "Specify dimensions to be cropped"
    result = tensor_in
    for dim in dims:
        start = (tensor_in.size()[dim] - crop_as.size()[dim]) // 2
        result = result.narrow(dim, start, crop_as.size()[dim])
    return result


class Block3x3x3(nn.Module):
    def __init__(self, n_input, n_channels):
        super(Block3x3x3, self).__init__()
        layers = []
        for cin in (n_input, n_channels):
            layers += [nn.BatchNorm3d(cin), nn.Conv3d(cin, n_channels, 3, stride=1, padding=0), nn.LeakyReLU(0.01, True)]
        self.bn_conv_relu_2x = nn.Sequential(*layers)

    def forward(self, input_maps):
        plan = self.__dict__.get('_sp_plan')
        if plan is None or plan[0] is not self.bn_conv_relu_2x:
            plan = (self.bn_conv_relu_2x, engine.SeqPlan(self.bn_conv_relu_2x))
            self.__dict__['_sp_plan'] = plan
        if not input_maps.is_cuda:
            raise RuntimeError("Block3x3x3: CUDA tensors only — there is no CPU path")
        return engine.run_sequential(plan[1], input_maps)

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop('_sp_plan', None)
        return state


class Unet3D(nn.Module):
    def __init__(self, channels=[2, 32, 64, 128, 64, 32, 32, 2], channel_dim=1, channels_crop=[2, 3, 4]):
        super(Unet3D, self).__init__()
        n_ch_in, ch_b1, ch_b2, ch_b3, ch_b4, ch_b5, ch_bC, n_classes = channels

        self.channel_dim = channel_dim
        self.channels_crop = channels_crop
        # None: each nn.Upsample decides (installed-torch default False; legacy torch-0.3.1 pickles -> True, see
        # engine._align); True / False: explicit switch for both upsampling steps
        self.align_corners = None

        self.block1 = Block3x3x3(n_ch_in, ch_b1)
        self.pool12 = nn.MaxPool3d(2, 2)
        self.block2 = Block3x3x3(ch_b1, ch_b2)
        self.pool23 = nn.MaxPool3d(2, 2)
        self.block3 = Block3x3x3(ch_b2, ch_b3)

        self.upsa34 = nn.Upsample(scale_factor=2, mode='trilinear')
        self.block4 = Block3x3x3(ch_b3 + ch_b2, ch_b4)
        self.upsa45 = nn.Upsample(scale_factor=2, mode='trilinear')
        self.block5 = Block3x3x3(ch_b4 + ch_b1, ch_b5)

        self.classify = nn.Sequential(
            nn.Conv3d(ch_b5, ch_bC, 1, stride=1, padding=0),
            nn.LeakyReLU(0.01, True),
            nn.Conv3d(ch_bC, n_classes, 1, stride=1, padding=0),
            nn.Sigmoid()
        )

    def _plan(self):
        plan = self.__dict__.get('_sp_plan')
        if plan is None:
            plan = engine.UnetPlan(self)
            self.__dict__['_sp_plan'] = plan
        return plan

    def forward(self, dto: UnetDto):
        x = dto.given_variables.input_modalities
        if not x.is_cuda:
            raise RuntimeError("Unet3D: CUDA tensors only — there is no CPU path")
        if self.channel_dim != 1 or list(self.channels_crop) != [2, 3, 4]:
            raise RuntimeError("Unet3D: only channel_dim=1 / channels_crop=[2,3,4] (the reference's use) is supported")
        outs = engine.run_unet(self._plan(), x)
        dto.outputs.core = outs[0]
        dto.outputs.penu = outs[1]
        return dto

    def freeze(self, freeze=False):
        requires_grad = not freeze
        for param in self.parameters():
            param.requires_grad = requires_grad

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop('_sp_plan', None)
        return state
