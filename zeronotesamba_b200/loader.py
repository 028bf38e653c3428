"""Drop-in for /root/reference/zeroNoteSamba/loader.py::load_models (lines 8-69): loss, optimizer
and model for the downstream (beat tracking) experiments.  Same `_status` / `_pre` / `_lr` semantics;
the checkpoint path is a parameter (the reference hard-codes "models/saved/shift_pret_cnn_16.pth",
a blob that is not shipped) and a synthetic state_dict can be passed directly."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .models.models import DS_CNN, Down_CNN


def load_models(_status: str, _pre: str, _lr: float, checkpoint: Optional[str] = None,
                state_dict: Optional[Dict[str, torch.Tensor]] = None
                ) -> Tuple[torch.nn.BCELoss, torch.optim.Adam, torch.nn.Module]:
    """
    Function for loading loss, optimizer, and model.
    -- _status: pretrained, vanilla, clmr?
    -- _pre: frozen weights
    -- _lr: learning rate
    """
    criterion = torch.nn.BCELoss().cuda()
    model: torch.nn.Module
    if _status == "pretrained":
        model = Down_CNN().cuda()
        if state_dict is None:
            state_dict = torch.load(checkpoint or "models/saved/shift_pret_cnn_16.pth", map_location=torch.device("cuda"))
        model.pretext.load_state_dict(state_dict)
        if _pre == "frozen":
            for param in model.pretext.anchor.pretrained.parameters():
                param.requires_grad = False
            for param in model.pretext.postve.pretrained.parameters():
                param.requires_grad = False
            optimizer = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=_lr, betas=(0.9, 0.999))
        else:
            optimizer = torch.optim.Adam(model.parameters(), lr=0.5 * _lr * 10e-2, betas=(0.9, 0.999))   # loader.py:43
    elif _status == "clmr":
        model = DS_CNN().cuda()
        if state_dict is None:
            state_dict = torch.load(checkpoint or "models/saved/clmr_pret_cnn_16.pth", map_location=torch.device("cuda"))
        model.load_state_dict(state_dict)
        if _pre == "frozen":
            for param in model.pretrained.parameters():
                param.requires_grad = False
            optimizer = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()), lr=_lr, betas=(0.9, 0.999))
        else:
            optimizer = torch.optim.Adam(model.parameters(), lr=0.5 * _lr, betas=(0.9, 0.999))
    else:
        model = DS_CNN().cuda()
        optimizer = torch.optim.Adam(model.parameters(), lr=_lr, betas=(0.9, 0.999))
    return criterion, optimizer, model
