"""Device-side data transforms: the reference's per-sample numpy / scipy augmentations
(common/data.py:215-380) re-stated for BATCHES that already live on the GPU in torch layout B x C x D x H x W
(``ToTensor``, data.py:299-310: D = z, H = y, W = x = the reference's numpy axis 0).

Class names, constructor arguments and the random decisions (python ``random`` for flips / patch origins, exactly the
calls the reference makes) follow the reference; ``__call__`` takes and returns the batch dict (``images`` / ``labels`` /
``clinical`` / ``case_id``).  ``ElasticDeform`` draws its uniform noise on the device (``torch.rand``, fp64) unless the
caller hands in the fields (``noise=``, reference draw order) — e.g. the numpy ``RandomState`` stream for a bit-for-bit
replay of a reference run.  Nothing here computes on the CPU.
"""
import random

import torch

from .. import ops
from . import data

KEY_IMAGES, KEY_LABELS, KEY_GLOBAL, KEY_CASE_ID = data.KEY_IMAGES, data.KEY_LABELS, data.KEY_GLOBAL, data.KEY_CASE_ID


def _has(batch, key):
    t = batch.get(key)
    return torch.is_tensor(t) and t.dim() == 5 and t.numel() > 0 and t.shape[-1] > 1


def _copy(batch):
    return dict(batch)


class ToTensor(object):
    """numpy sample [x][y][z][c] (or a list of them) -> device batch B x C x Z x Y x X (data.py:299-310); the only host
    work is the H2D copy."""

    def __init__(self, device='cuda'):
        self._device = device

    def __call__(self, sample):
        out = _copy(sample)
        for k in (KEY_IMAGES, KEY_LABELS, KEY_GLOBAL):
            v = sample.get(k)
            if v is None or (isinstance(v, list) and not v):
                continue
            t = torch.as_tensor(v)
            if t.dim() == 4:
                t = t.unsqueeze(0)
            out[k] = t.to(self._device, non_blocking=True).permute(0, 4, 3, 2, 1).contiguous()
        return out


class HemisphericFlip(object):
    """Flip along the X axis with probability 1/2 (data.py:232-245).  One decision per batch element."""

    def __call__(self, batch):
        n = next(batch[k].shape[0] for k in (KEY_LABELS, KEY_IMAGES) if _has(batch, k))
        flips = [random.random() > 0.5 for _ in range(n)]
        return _flip_selected(batch, flips)


class HemisphericFlipFixedToCaseId(object):
    def __init__(self, split_id):
        self.split_id = split_id

    def __call__(self, batch):
        ids = batch[KEY_CASE_ID]
        ids = ids.tolist() if torch.is_tensor(ids) else list(ids)
        return _flip_selected(batch, [int(c) > self.split_id for c in ids])


def _flip_selected(batch, flips):
    out = _copy(batch)
    if not any(flips):
        return out
    for k in (KEY_IMAGES, KEY_LABELS):
        if not _has(batch, k):
            continue
        x = batch[k]
        flipped = ops.flip_w(x)
        if all(flips):
            out[k] = flipped
        else:
            sel = torch.tensor(flips, device=x.device).view(-1, 1, 1, 1, 1)
            out[k] = torch.where(sel, flipped, x.float())
    return out


class RandomPatch(object):
    """Random patch of w x h x d voxels (x, y, z) from the padded images and the matching unpadded label window
    (data.py:248-277); ONE origin per call, like the reference's per-sample transform applied to a batch of one."""

    def __init__(self, w, h, d, pad_x, pad_y, pad_z):
        self._padx, self._pady, self._padz = pad_x, pad_y, pad_z
        self._w, self._h, self._d = w, h, d

    def __call__(self, batch):
        img = batch[KEY_IMAGES]
        sz, sy, sx = img.shape[2:]
        rand_x = random.randint(0, sx - self._w)
        rand_y = random.randint(0, sy - self._h)
        rand_z = random.randint(0, sz - self._d)
        out = _copy(batch)
        out[KEY_IMAGES] = ops.crop_volume(img, (rand_z, rand_y, rand_x), (self._d, self._h, self._w))
        if _has(batch, KEY_LABELS):
            out[KEY_LABELS] = ops.crop_volume(batch[KEY_LABELS], (rand_z, rand_y, rand_x),
                                              (self._d - 2 * self._padz, self._h - 2 * self._pady, self._w - 2 * self._padx))
        return out


class PadImages(object):
    """Constant border around the images (data.py:280-296)."""

    def __init__(self, pad_x, pad_y, pad_z, pad_value=0):
        self._padx, self._pady, self._padz = pad_x, pad_y, pad_z
        self._pad_value = float(pad_value)

    def __call__(self, batch):
        out = _copy(batch)
        if _has(batch, KEY_IMAGES):
            out[KEY_IMAGES] = ops.pad_volume(batch[KEY_IMAGES], self._padz, self._pady, self._padx, self._pad_value)
        return out


class ElasticDeform(object):
    """Elastic deformation (data.py:313-351): every label channel (and, with ``apply_to_images``, image channel) is warped
    by its OWN triple of smoothed noise fields — the reference draws three new fields per channel from one running
    RandomState — of amplitude alpha (z: alpha * 0.22), Gaussian sigma in voxels."""

    def __init__(self, alpha=100, sigma=4, apply_to_images=False, generator=None):
        self._alpha, self._sigma, self._apply_to_images = alpha, sigma, apply_to_images
        self._generator = generator

    def _warp(self, x, noise):
        B, C, D, H, W = x.shape
        if H != W:
            raise RuntimeError("ElasticDeform: the reference's coordinate grid (np.meshgrid 'xy', data.py:340) needs X == Y")
        if noise is None:
            noise = torch.rand((3, B, C, D, H, W), device=x.device, dtype=torch.float64, generator=self._generator)
        else:
            noise = noise.to(x.device, torch.float64)
            if tuple(noise.shape) != (3, B, C, D, H, W):
                raise RuntimeError("ElasticDeform: noise must be [3, B, C, D, H, W] (draw order dx, dy, dz per channel)")
        fields = ops.gauss3d(noise * 2 - 1, self._sigma)
        return ops.elastic_warp(x, fields[0], fields[1], fields[2], float(self._alpha), 0.22)

    def __call__(self, batch, noise=None, image_noise=None):
        out = _copy(batch)
        out[KEY_LABELS] = self._warp(batch[KEY_LABELS], noise)
        if self._apply_to_images and _has(batch, KEY_IMAGES):
            out[KEY_IMAGES] = self._warp(batch[KEY_IMAGES], image_noise)
        return out


class ResamplePlaneXY(object):
    """In-plane down- / up-sampling of every slice (data.py:354-380, scipy.ndimage.zoom order 0 or 1)."""

    def __init__(self, scale_factor=1, mode='nearest'):
        self._scale_factor = scale_factor
        self._order = 1 if mode == 'bilinear' else 0

    def __call__(self, batch):
        out = _copy(batch)
        for k in (KEY_IMAGES, KEY_LABELS):
            if _has(batch, k):
                out[k] = ops.zoom_plane_xy(batch[k], self._scale_factor, self._order)
        return out


class Compose(object):
    def __init__(self, transforms):
        self.transforms = list(transforms)

    def __call__(self, batch):
        for t in self.transforms:
            batch = t(batch)
        return batch
