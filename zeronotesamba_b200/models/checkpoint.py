"""Synthetic checkpoints with the reference's key layout.

The shipped ``models/saved/shift_pret_cnn_16.pth`` is a missing large blob
(/root/reference/.MISSING_LARGE_BLOBS) and PyTorch's default init gives time-constant embeddings
(cosines == 1, gradients ~1e-10), so parity and throughput runs use a
deterministic He-normal state_dict.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from ..engine import CONV_SPECS

BRANCHES = ("anchor", "postve")


def state_dict_layout(prefix: str = "") -> Dict[str, tuple]:
    out: Dict[str, tuple] = {}
    for br in BRANCHES:
        for name, co, ci, kh, kw, _ in CONV_SPECS:
            out[f"{prefix}{br}.pretrained.{name}.weight"] = (co, ci, kh, kw)
            out[f"{prefix}{br}.pretrained.{name}.bias"] = (co,)
        out[f"{prefix}{br}.fc1.weight"] = (1, 128, 1)
        out[f"{prefix}{br}.fc1.bias"] = (1,)
    return out


def he_normal_state_dict(seed: int = 0, prefix: str = "") -> Dict[str, torch.Tensor]:
    """He-normal conv weights (std sqrt(2/fan_in)), zero biases, fc1 std sqrt(1/128); numpy PCG64."""
    rng = np.random.default_rng(seed)
    out: Dict[str, torch.Tensor] = {}
    for key, shape in state_dict_layout(prefix).items():
        if key.endswith("bias"):
            arr = np.zeros(shape, np.float32)
        elif ".fc1." in key:
            arr = (rng.standard_normal(shape) * math.sqrt(1.0 / 128.0)).astype(np.float32)
        else:
            fan_in = shape[1] * shape[2] * shape[3]
            arr = (rng.standard_normal(shape) * math.sqrt(2.0 / fan_in)).astype(np.float32)
        out[key] = torch.from_numpy(arr)
    return out
