"""Batch-dictionary keys of the reference data pipeline (common/data.py:18-27) and synthetic batch generators.

The reference's NIfTI dataset and scipy augmentations are out of the hot-path scope; benchmarks and
tests use synthetic volumes of the shapes/distributions fixed in SURVEY §8(d).
"""
import numpy as np
import torch

KEY_CASE_ID = 'case_id'
KEY_CLINICAL_IDX = 'clinical_idx'
KEY_IMAGES = 'images'
KEY_LABELS = 'labels'
KEY_GLOBAL = 'clinical'

DIM_HORIZONTAL_NUMPY_3D = 0
DIM_DEPTH_NUMPY_3D = 2
DIM_CHANNEL_NUMPY_3D = 3
DIM_CHANNEL_TORCH3D_5 = 1


def _ellipsoid_masks(rng, B, size, radii_fracs):
    """Nested random ellipsoids (same centre, growing radii) -> float32 {0,1} masks [B, len(radii), D, H, W]."""
    D, H, W = size
    zz, yy, xx = np.meshgrid(np.arange(D), np.arange(H), np.arange(W), indexing='ij')
    out = np.zeros((B, len(radii_fracs), D, H, W), dtype=np.float32)
    for b in range(B):
        c = np.array([rng.uniform(0.35, 0.65) * D, rng.uniform(0.35, 0.65) * H, rng.uniform(0.35, 0.65) * W])
        aniso = rng.uniform(0.8, 1.25, size=3)
        for j, f in enumerate(radii_fracs):
            r = np.maximum(np.array([D, H, W]) * f * aniso, 1.0)
            out[b, j] = ((((zz - c[0]) / r[0]) ** 2 + ((yy - c[1]) / r[1]) ** 2 + ((xx - c[2]) / r[2]) ** 2) <= 1.0)
    return out


def synthetic_cae_batch(batch, size=(28, 128, 128), seed=4):
    """Config-2 style batch: labels B x 3 x D x H x W (core, penumbra, lesion; core < lesion < penumbra) and
    clinical B x 5 x 1 x 1 x 1 float64 (ch0 onset->admission h, ch1 admission->recanalisation h)."""
    rng = np.random.RandomState(seed)
    m = _ellipsoid_masks(rng, batch, size, (0.13, 0.27, 0.19))   # core, penumbra, lesion
    clinical = np.zeros((batch, 5, 1, 1, 1), dtype=np.float64)
    clinical[:, 0, 0, 0, 0] = rng.uniform(0.5, 4.0, size=batch)
    clinical[:, 1, 0, 0, 0] = rng.uniform(0.0, 5.0, size=batch)
    clinical[:, 2, 0, 0, 0] = rng.randint(0, 25, size=batch)
    clinical[:, 3, 0, 0, 0] = rng.randint(0, 2, size=batch)
    clinical[:, 4, 0, 0, 0] = rng.uniform(40, 90, size=batch)
    return {KEY_IMAGES: torch.zeros(batch, 2, 1, 1, 1), KEY_LABELS: torch.from_numpy(m),
            KEY_GLOBAL: torch.from_numpy(clinical), KEY_CASE_ID: torch.arange(batch)}


def synthetic_unet_batch(batch, out_size=(28, 128, 128), pad=(20, 20, 20), seed=4):
    """Config-1 style batch: images B x 2 x (D+2p) x (H+2p) x (W+2p) with an exactly-zero border of `pad` voxels
    (PadImages, data.py:280-296), interior CBV ~ U[0,12) / TTD ~ U[0,40); labels B x 2 x D x H x W (core, penumbra)."""
    rng = np.random.RandomState(seed)
    D, H, W = out_size
    img = np.zeros((batch, 2, D + 2 * pad[0], H + 2 * pad[1], W + 2 * pad[2]), dtype=np.float32)
    img[:, 0, pad[0]:pad[0] + D, pad[1]:pad[1] + H, pad[2]:pad[2] + W] = rng.uniform(0, 12, size=(batch, D, H, W))
    img[:, 1, pad[0]:pad[0] + D, pad[1]:pad[1] + H, pad[2]:pad[2] + W] = rng.uniform(0, 40, size=(batch, D, H, W))
    m = _ellipsoid_masks(rng, batch, out_size, (0.13, 0.27))
    return {KEY_IMAGES: torch.from_numpy(img), KEY_LABELS: torch.from_numpy(m),
            KEY_GLOBAL: torch.zeros(batch, 5, 1, 1, 1, dtype=torch.float64), KEY_CASE_ID: torch.arange(batch)}
