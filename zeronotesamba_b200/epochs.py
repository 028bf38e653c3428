"""Downstream fine-tuning epochs -- drop-in for the compute of
/root/reference/zeroNoteSamba/epochs.py::train_epoch / val_epoch (lines 8-187): one file per step
(batch size 1, variable T), ``model(vqt1, vqt2)`` -> BCELoss against the pulse mask -> backward ->
Adam.  Batch-1 inputs run time-folded over the eight clip slots (models.fold_plan).

The reference scores every output with madmom's DBN + mir_eval inside the loop
(processing/evaluate.py, CPU, out of scope here): pass ``evaluator(cpu_output, times, threshold=...,
librosa=...) -> 6 floats`` to get the same metrics; without it the six metrics are returned as 0.0.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Tuple

import torch


def _zeros(*_a, **_k):
    return (0.0,) * 6


def _forward(model, _status, vqt, train: bool):
    if _status == "pretrained":
        vqt1 = torch.reshape(vqt[0, :, :], (1, 1, vqt.shape[1], vqt.shape[2])).cuda()
        vqt2 = torch.reshape(vqt[1, :, :], (1, 1, vqt.shape[1], vqt.shape[2])).cuda()
        return model(vqt1, vqt2)
    vqt = torch.reshape(vqt[:, :], (1, 1, vqt.shape[0], vqt.shape[1])).cuda()
    return model(vqt)


def train_epoch(model: torch.nn.Module, criterion: torch.nn.BCELoss, optimizer: torch.optim.Adam, _status: str,
                indices: List[str], real_times: Dict[str, Any], inputs: Dict[str, Any], masks: Dict[str, Any],
                threshold: bool, librosa: bool, evaluator: Optional[Callable] = None
                ) -> Tuple[torch.nn.Module, torch.optim.Adam, float, float, float, float, float, float, float]:
    """Training epoch (reference signature plus the optional evaluator)."""
    ev = evaluator or _zeros
    full_loss, sums = 0.0, [0.0] * 6
    model.train()
    n = 0
    for wav in indices:
        msk = masks[wav]
        msk = torch.reshape(msk, (1, msk.shape[0])).cuda()
        optimizer.zero_grad()
        output = _forward(model, _status, inputs[wav], True)
        loss = criterion(output, msk)
        loss.backward()
        optimizer.step()
        full_loss += loss.item()
        res = ev(output.squeeze(0).cpu().detach().numpy(), real_times[wav], threshold=threshold, librosa=librosa)
        sums = [a + b for a, b in zip(sums, res)]
        n += 1
    n = max(n, 1)
    return (model, optimizer, full_loss / n) + tuple(v / n for v in sums)


def val_epoch(model: torch.nn.Module, criterion: torch.nn.BCELoss, _status: str, indices: List[str],
              real_times: Dict[str, Any], inputs: Dict[str, Any], masks: Dict[str, Any], threshold: bool, librosa: bool,
              evaluator: Optional[Callable] = None) -> Tuple[float, float, float, float, float, float, float]:
    """Validation epoch (reference signature plus the optional evaluator)."""
    ev = evaluator or _zeros
    full_loss, sums = 0.0, [0.0] * 6
    model.eval()
    n = 0
    for wav in indices:
        with torch.no_grad():
            msk = masks[wav]
            msk = torch.reshape(msk, (1, msk.shape[0])).cuda()
            output = _forward(model, _status, inputs[wav], False)
            full_loss += criterion(output, msk).item()
        res = ev(output.squeeze(0).cpu().numpy(), real_times[wav], threshold=threshold, librosa=librosa)
        sums = [a + b for a, b in zip(sums, res)]
        n += 1
    n = max(n, 1)
    return (full_loss / n,) + tuple(v / n for v in sums)
