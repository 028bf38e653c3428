"""Drop-in for /root/reference/zeroNoteSamba/models/loss_functions.py (NTXent, lines 7-55).

Same constructor and ``forward(anchors, poss) -> (loss, mean cos(anchor,pos), mean cos(anchor,neg))``
contract; the reference's per-row Python loop (~20 launches and 3 host syncs per row) is one fused
forward+backward launch (zns_ntxent_fwd_bwd).  The loss is a 0-d tensor on the embeddings' device
with autograd attached (the reference's lives on the CPU); the two similarities are Python floats
as in the reference (one host sync) unless ``sync_stats=False``, in which case they are 0-d device
tensors and nothing synchronises.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn

from .. import _lib as L


class _NTXentFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchors, poss, batch_len, temperature):
        a = anchors.contiguous().float()
        p = poss.contiguous().float()
        n, d = a.shape
        res = torch.empty(3, device=a.device, dtype=torch.float32)
        da = torch.empty_like(a)
        dp = torch.empty_like(p)
        L.check(L.lib().zns_ntxent_fwd_bwd(L.ptr(a), L.ptr(p), n, d, int(batch_len), float(temperature), L.ptr(res),
                                           L.ptr(da), L.ptr(dp), L.current_stream()))
        ctx.save_for_backward(da, dp)
        ctx.mark_non_differentiable(res)
        return res[0].clone(), res

    @staticmethod
    def backward(ctx, g_loss, _g_res):
        da, dp = ctx.saved_tensors
        return da * g_loss, dp * g_loss, None, None


class NTXent(nn.Module):
    """
    Compute NT-Xent loss for contrastive learning.
    """

    def __init__(self, batch_len: int, temperature: float = 0.25, sync_stats: bool = True):
        """
        Arguments:
        -- batch_len: batch size
        -- temperature: parameter
        """
        super(NTXent, self).__init__()
        self.batch_len = batch_len
        self.temperature = temperature
        self.sync_stats = sync_stats

    def forward(self, anchors: torch.Tensor, poss: torch.Tensor) -> Tuple[torch.Tensor, float, float]:
        """
        Arguments:
        -- anchors: tensor of shape (batch_len, embedding_size)
        -- poss: tensor of shape (batch_len, embedding_size)
        """
        if not anchors.is_cuda:
            raise RuntimeError("zeronotesamba_b200.NTXent runs on the GPU only (no CPU fallback)")
        loss, res = _NTXentFunction.apply(anchors, poss, self.batch_len, self.temperature)
        if self.sync_stats:
            host = res.tolist()
            return loss, host[1], host[2]
        return loss, res[1], res[2]



class _BCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target):
        o = output.contiguous().float()
        t = target.contiguous().float()
        loss = torch.empty(1, device=o.device, dtype=torch.float32)
        d_out = torch.empty_like(o)
        L.check(L.lib().zns_bce_fwd_bwd(L.ptr(o), L.ptr(t), o.numel(), L.ptr(loss), L.ptr(d_out), L.current_stream()))
        ctx.save_for_backward(d_out)
        ctx.shape = output.shape
        return loss[0].clone()

    @staticmethod
    def backward(ctx, g_loss):
        (d_out,) = ctx.saved_tensors
        return (d_out * g_loss).view(ctx.shape), None


class FusedBCELoss(nn.Module):
    """``torch.nn.BCELoss()`` (mean reduction) as the downstream loops use it (loader.py:20, epochs.py:52-54):
    ``criterion(output, mask) -> 0-d loss`` with autograd; loss and d loss / d output come from one launch
    (zns_bce_fwd_bwd) instead of the ~10 elementwise launches of the composed op."""

    def forward(self, output: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if not output.is_cuda:
            raise RuntimeError("zeronotesamba_b200.FusedBCELoss runs on the GPU only (no CPU fallback)")
        if output.shape != target.shape:
            raise ValueError(f"Using a target size ({tuple(target.shape)}) that is different to the input size ({tuple(output.shape)})")
        return _BCEFunction.apply(output, target)
