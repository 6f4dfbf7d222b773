"""Signed-distance-map interpolation baseline on the device (reference: test_sdm_resampling.py:15-52).

``sdm_interpolate`` mirrors ``sdm_interpolate_numpy(..., resample=False)``: signed distance maps of the thresholded core
and penumbra by exact Euclidean distance transforms (``sp_signed_distance``), linear interpolation between them at the
normalised time, segmentation by the sign.  The reference's optional x12 down / up sampling of the maps (``resample=True``,
cubic ``scipy.ndimage.zoom``) is a host-side spline resampler that is not part of this path and is not provided.
"""
import torch

from .. import ops


def _artificial_core(penu_bin, dilate):
    """test_sdm_resampling.py:24-29: no voxel above threshold -> one voxel at the penumbra's centre of mass, dilated
    `dilate` times with the connectivity-1 cross (device-side; only for the degenerate case)."""
    idx = penu_bin.nonzero().float()
    cog = [int(v) for v in idx.mean(0).tolist()]
    core = torch.zeros_like(penu_bin, dtype=torch.float32)
    core[cog[0], cog[1], cog[2]] = 1.0
    for _ in range(dilate):
        grown = core.clone()
        for ax in range(3):
            n = core.shape[ax]
            grown.narrow(ax, 1, n - 1).copy_(torch.maximum(grown.narrow(ax, 1, n - 1), core.narrow(ax, 0, n - 1)))
            grown.narrow(ax, 0, n - 1).copy_(torch.maximum(grown.narrow(ax, 0, n - 1), core.narrow(ax, 1, n - 1)))
        core = grown
    return core


def sdm_interpolate(core, penu, interpolation, threshold=0.5, dilate=3):
    """core, penu: 1 x 1 x D x H x W CUDA volumes; interpolation: normalised time in [0, 1] (float or 0-dim tensor).

    Returns (recon_core, recon_intp, recon_penu) signed distance maps with the reference's sign conventions: the core map
    is NEGATIVE inside the core (segment with ``< 0``), the penumbra and interpolated maps POSITIVE inside (``> 0``)."""
    if not (core.is_cuda and penu.is_cuda):
        raise RuntimeError("sdm_interpolate: CUDA tensors only — there is no CPU path")
    c = core[0, 0].contiguous().float()
    p = penu[0, 0].contiguous().float()
    penu_dist = ops.signed_distance(p, threshold, outside_is_lt=True, sign=1.0)
    if not bool((c > threshold).any()):
        # the reference subtracts edt(core > threshold) of the ORIGINAL (empty) core: the map is 0, not negative, inside the
        # artificial core (test_sdm_resampling.py:30-31)
        c = _artificial_core(p > threshold, dilate)
        core_dist = ops.signed_distance(c, 0.5, outside_is_lt=False, sign=-1.0).clamp_(min=0.0)
    else:
        core_dist = ops.signed_distance(c, threshold, outside_is_lt=False, sign=-1.0)
    t = float(interpolation)
    intp = penu_dist * t - core_dist * (1 - t)
    return core_dist, intp, penu_dist
