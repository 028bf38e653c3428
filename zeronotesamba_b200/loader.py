"""Model / loss / optimizer factory of the downstream (beat tracking) experiments.

Interface of /root/reference/zeroNoteSamba/loader.py::load_models (lines 8-69): ``load_models(_status, _pre, _lr)``
returns ``(criterion, optimizer, model)`` with the reference's status names and learning-rate rules.  Two additions:
the checkpoint location is a parameter (the reference hard-codes ``models/saved/*_pret_cnn_16.pth``, blobs it does
not ship) and a ``state_dict`` can be handed over directly; the optimizer is this package's ``FusedAdam`` (Adam
defaults, one fused kernel per tensor) and the loss ``FusedBCELoss`` (BCELoss forward + backward in one launch) unless
``fused=False`` asks for ``torch.optim.Adam`` / ``torch.nn.BCELoss``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Optional, Tuple

import torch

from .models.models import DS_CNN, Down_CNN


@dataclass(frozen=True)
class _Recipe:
    """How one ``_status`` is assembled."""
    build: Callable[[], torch.nn.Module]
    checkpoint: Optional[str]                                  # default file; None = random initialisation
    load_into: Callable[[torch.nn.Module], torch.nn.Module]    # sub-module that receives the checkpoint
    encoders: Callable[[torch.nn.Module], Tuple[torch.nn.Module, ...]]   # what "frozen" freezes
    finetune_lr: Callable[[float], float]                      # learning rate when the encoders are trained too


_RECIPES: Dict[str, _Recipe] = {
    # ZeroNS: two pretrained encoders under the max/mean head (loader.py:22-43; the 0.5 * lr * 10e-2 rule is :43)
    "pretrained": _Recipe(Down_CNN, "models/saved/shift_pret_cnn_16.pth", lambda m: m.pretext,
                          lambda m: (m.pretext.anchor.pretrained, m.pretext.postve.pretrained),
                          lambda lr: 0.5 * lr * 10e-2),
    # CLMR baseline: one shared encoder (loader.py:45-62)
    "clmr": _Recipe(DS_CNN, "models/saved/clmr_pret_cnn_16.pth", lambda m: m, lambda m: (m.pretrained,),
                    lambda lr: 0.5 * lr),
}
_VANILLA = _Recipe(DS_CNN, None, lambda m: m, lambda m: (), lambda lr: lr)          # any other status (loader.py:64-67)


def load_models(_status: str, _pre: str, _lr: float, checkpoint: Optional[str] = None,
                state_dict: Optional[Dict[str, torch.Tensor]] = None, fused: bool = True):
    """Loss, optimizer and model for a downstream run.
    -- _status: "pretrained" (ZeroNS), "clmr", anything else = vanilla (random initialisation)
    -- _pre: "frozen" keeps the pretrained encoders fixed and trains the rest at ``_lr``
    -- _lr: base learning rate
    """
    recipe = _RECIPES.get(_status, _VANILLA)
    model = recipe.build().cuda()
    if fused:
        from .models.loss_functions import FusedBCELoss
        criterion = FusedBCELoss().cuda()
    else:
        criterion = torch.nn.BCELoss().cuda()
    if recipe.checkpoint is not None:
        weights = state_dict if state_dict is not None else torch.load(checkpoint or recipe.checkpoint,
                                                                        map_location=torch.device("cuda"))
        recipe.load_into(model).load_state_dict(weights)
    frozen = _pre == "frozen" and recipe.checkpoint is not None
    if frozen:
        for enc in recipe.encoders(model):
            enc.requires_grad_(False)
    trainable = [p for p in model.parameters() if p.requires_grad]
    lr = _lr if (frozen or recipe.checkpoint is None) else recipe.finetune_lr(_lr)
    if fused:
        from .pretext import FusedAdam
        optimizer = FusedAdam(trainable, lr=lr, betas=(0.9, 0.999))
    else:
        optimizer = torch.optim.Adam(trainable, lr=lr, betas=(0.9, 0.999))
    return criterion, optimizer, model
