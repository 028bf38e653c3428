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


def length_buckets(indices: List[str], inputs: Dict[str, Any], max_batch: int = 16) -> List[List[str]]:
    """Group files of IDENTICAL frame count into batches of at most ``max_batch`` (first-seen order of the lengths).

    Batching files of one length is exact: every clip keeps its own zero "same" padding at its true edges, which a
    padded batch of mixed lengths would not (the eight convolutions see 68 frames either side).  Files whose
    length occurs once stay single and run through the batch-1 time-folded path."""
    by_len: Dict[int, List[str]] = {}
    for wav in indices:
        by_len.setdefault(int(inputs[wav].shape[-1]), []).append(wav)
    out: List[List[str]] = []
    for group in by_len.values():
        out += [group[i:i + max_batch] for i in range(0, len(group), max_batch)]
    return out


def batched_inference(model: torch.nn.Module, _status: str, indices: List[str], inputs: Dict[str, Any],
                      max_batch: int = 16) -> Dict[str, torch.Tensor]:
    """Activation curves ``{file: (T,) tensor on the GPU}`` of every file, equal-length files sharing a forward pass
    (what epochs.py:127-160 computes one file at a time)."""
    model.eval()
    result: Dict[str, torch.Tensor] = {}
    with torch.no_grad():
        for group in length_buckets(indices, inputs, max_batch):
            if len(group) == 1:
                result[group[0]] = _forward(model, _status, inputs[group[0]], False).squeeze(0)
                continue
            if _status == "pretrained":
                x = torch.stack([inputs[w] for w in group]).cuda()              # (n, 2, 96, T)
                out = model(x[:, 0:1], x[:, 1:2])
            else:
                out = model(torch.stack([inputs[w] for w in group]).unsqueeze(1).cuda())
            for i, w in enumerate(group):
                result[w] = out[i]
    return result


def val_epoch(model: torch.nn.Module, criterion: torch.nn.BCELoss, _status: str, indices: List[str],
              real_times: Dict[str, Any], inputs: Dict[str, Any], masks: Dict[str, Any], threshold: bool, librosa: bool,
              evaluator: Optional[Callable] = None, batch_files: int = 1
              ) -> Tuple[float, float, float, float, float, float, float]:
    """Validation epoch (reference signature plus the optional evaluator).  ``batch_files`` > 1 runs files of equal
    length together (length_buckets); losses and metrics are still taken per file, in the order of ``indices``."""
    ev = evaluator or _zeros
    full_loss, sums = 0.0, [0.0] * 6
    model.eval()
    n = 0
    if batch_files > 1:
        outputs = batched_inference(model, _status, indices, inputs, batch_files)
        for wav in indices:
            msk = masks[wav]
            msk = torch.reshape(msk, (1, msk.shape[0])).cuda()
            output = outputs[wav].unsqueeze(0)
            full_loss += criterion(output, msk).item()
            res = ev(output.squeeze(0).cpu().numpy(), real_times[wav], threshold=threshold, librosa=librosa)
            sums = [a + b for a, b in zip(sums, res)]
            n += 1
        n = max(n, 1)
        return (full_loss / n,) + tuple(v / n for v in sums)
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
