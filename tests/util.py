"""Shared helpers for the parity tests (fixtures, tolerances)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# fp32 parity bar of BASELINE.json north_star: activations and gradients within 1e-4 relative error,
# thresholded segmentations / Dice within 1e-3.
TOL_ACT = 1e-4
TOL_GRAD = 1e-4
TOL_DICE = 1e-3


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def state_from(fx, prefix):
    out = {}
    for k, v in fx.items():
        if k.startswith(prefix):
            out[k[len(prefix):]] = torch.from_numpy(np.array(v))
    return out


def unpack_masks(fx):
    shape = tuple(int(x) for x in fx["labels_shape"])
    n = int(np.prod(shape))
    bits = np.unpackbits(fx["labels_bits"])[:n]
    return torch.from_numpy(bits.reshape(shape).astype(np.float32))


def rel_l2(a, b):
    a = torch.as_tensor(a).detach().double().cpu().reshape(-1)
    b = torch.as_tensor(b).detach().double().cpu().reshape(-1)
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def rel_max(a, b):
    a = torch.as_tensor(a).detach().double().cpu().reshape(-1)
    b = torch.as_tensor(b).detach().double().cpu().reshape(-1)
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)
