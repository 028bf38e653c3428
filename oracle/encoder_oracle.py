"""CPU oracle for the encoders, the NT-Xent objective and the Adam step -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path never does.

A plain fp32 restatement (torch CPU functional ops + numpy) of
  * ``_CNN`` / ``DS_CNN`` / ``Pretext_CNN`` / ``Down_CNN``
    (/root/reference/zeroNoteSamba/models/models.py:7-150),
  * ``NTXent.forward`` (/root/reference/zeroNoteSamba/models/loss_functions.py:24-55),
  * one ``train_epoch`` / ``val_epoch`` batch (/root/reference/zeroNoteSamba/pretext.py:475-490,
    :546-564) with ``torch.optim.Adam(lr=1e-6)`` defaults (pretext.py:202).
PINNED: ``tools/make_golden.py`` imports the reference's own modules from /root/reference in the
build container, runs them on seeded inputs and writes ``tests/golden/encoder_golden.npz``;
``tests/test_oracle_encoder.py`` checks this restatement against those vectors.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# (name, Cout, Cin, kh, kw, pad_h, pad_w, pool_h)   models.py:16-28
CONV_SPECS = (
    ("cv1", 64, 1, 3, 11, 1, 5, 1),
    ("cv2", 64, 64, 7, 13, 3, 6, 3),
    ("cv3", 128, 64, 5, 15, 2, 7, 1),
    ("cv4", 128, 128, 9, 17, 4, 8, 4),
    ("cv5", 256, 128, 3, 19, 1, 9, 1),
    ("cv6", 256, 256, 5, 21, 2, 10, 8),
    ("cv7", 128, 256, 1, 23, 0, 11, 1),
    ("cv8", 128, 128, 1, 25, 0, 12, 1),
)
BRANCHES = ("anchor", "postve")


def state_dict_layout(prefix: str = "") -> Dict[str, Tuple[int, ...]]:
    """Key -> shape of ``Pretext_CNN.state_dict()`` (36 tensors), in module registration order."""
    out: Dict[str, Tuple[int, ...]] = {}
    for br in BRANCHES:
        for name, co, ci, kh, kw, _, _, _ in CONV_SPECS:
            out[f"{prefix}{br}.pretrained.{name}.weight"] = (co, ci, kh, kw)
            out[f"{prefix}{br}.pretrained.{name}.bias"] = (co,)
        out[f"{prefix}{br}.fc1.weight"] = (1, 128, 1)
        out[f"{prefix}{br}.fc1.bias"] = (1,)
    return out


def cnn_forward(sd: Dict[str, torch.Tensor], branch: str, x: torch.Tensor, dropout_p: float = 0.0,
                train: bool = False) -> torch.Tensor:
    """``_CNN.forward`` (models.py:32-74): conv -> [maxpool over frequency] -> ReLU -> Dropout, x8."""
    out = x
    for name, _, _, _, _, ph, pw, pool in CONV_SPECS:
        w = sd[f"{branch}.pretrained.{name}.weight"]
        b = sd[f"{branch}.pretrained.{name}.bias"]
        out = F.conv2d(out, w, b, padding=(ph, pw))
        if pool > 1:
            out = F.max_pool2d(out, (pool, 1))
        out = F.relu(out)
        out = F.dropout(out, p=dropout_p, training=train)
    return torch.squeeze(out, dim=2)


def ds_cnn_forward(sd, branch, x, dropout_p=0.0, train=False):
    """``DS_CNN.forward`` (models.py:93-103): _CNN -> Conv1d(128,1,1) -> Sigmoid -> (B, T)."""
    h = cnn_forward(sd, branch, x, dropout_p, train)
    z = F.conv1d(h, sd[f"{branch}.fc1.weight"], sd[f"{branch}.fc1.bias"])
    e = torch.sigmoid(z)
    return e.reshape(e.size(0), e.size(1) * e.size(2))


def pretext_forward(sd, anc, pos, dropout_p=0.0, train=False):
    """``Pretext_CNN.forward`` (models.py:117-124)."""
    return ds_cnn_forward(sd, "anchor", anc, dropout_p, train), ds_cnn_forward(sd, "postve", pos, dropout_p, train)


def down_forward(sd, anc, pos, reduction: str = "max"):
    """``Down_CNN.forward`` (models.py:139-150); ``sd`` keys without the ``pretext.`` prefix."""
    a, p = pretext_forward(sd, anc, pos)
    if reduction == "mean":
        return torch.div(a + p, 2)
    return torch.maximum(a, p)


def ntxent(anchors: torch.Tensor, poss: torch.Tensor, batch_len: int, temperature: float = 0.25):
    """``NTXent.forward`` (loss_functions.py:24-55), differentiable closed form.

    loss = mean over ``batch_len`` slots (rows beyond anchors.shape[0] contribute zeros,
    loss_functions.py:30) of -log(exp(s_ii/tau) / sum_j exp(s_ij/tau)); cosine with eps = 1e-8
    clamped on each norm (torch >= 2 CosineSimilarity: x/max(|x|,eps) . y/max(|y|,eps); torch 1.13,
    the reference's pin, clamps the product instead -- identical unless a norm is < 1e-8).
    Returns (loss tensor, mean cos(anchor, positive), mean over rows of mean cos(anchor, negatives)).
    """
    n = anchors.shape[0]
    an = anchors.norm(dim=1, keepdim=True)
    pn = poss.norm(dim=1, keepdim=True)
    s = (anchors / torch.clamp(an, min=1e-8)) @ (poss / torch.clamp(pn, min=1e-8)).t()
    z = s / temperature
    lse = torch.logsumexp(z, dim=1)
    per_row = lse - torch.diagonal(z)
    loss = per_row.sum() / batch_len
    sd = s.detach()
    cos_pos = float(torch.diagonal(sd).sum() / batch_len)
    cos_neg = float(((sd.sum(dim=1) - torch.diagonal(sd)) / (batch_len - 1)).sum() / batch_len)
    return loss, cos_pos, cos_neg


def ntxent_loop_numpy(anchors: np.ndarray, poss: np.ndarray, batch_len: int, temperature: float):
    """Row-by-row float32 restatement of the reference loop (loss_functions.py:35-49)."""
    a = anchors.astype(np.float32)
    p = poss.astype(np.float32)
    full = np.zeros(batch_len, dtype=np.float32)
    cos_pos = 0.0
    cos_neg = 0.0

    def cs(x, y):  # CosineSimilarity(dim=1, eps=1e-8)
        xn = x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), np.float32(1e-8))
        yn = y / np.maximum(np.linalg.norm(y, axis=1, keepdims=True), np.float32(1e-8))
        return (xn * yn).sum(axis=1, dtype=np.float32).astype(np.float32)

    for xx in range(a.shape[0]):
        sim_pos = cs(a[xx:xx + 1], p[xx:xx + 1])
        cos_pos += float(sim_pos[0])
        sim_num = np.exp(sim_pos / np.float32(temperature))
        sim_all = cs(np.repeat(a[xx:xx + 1], p.shape[0], axis=0), p)
        cos_neg += (float(sim_all.sum(dtype=np.float32)) - float(sim_all[xx])) / (batch_len - 1)
        sim_den = np.exp(sim_all / np.float32(temperature)).sum(dtype=np.float32)
        full[xx] = -np.log(sim_num[0] / sim_den)
    return float(full.mean(dtype=np.float32)), cos_pos / batch_len, cos_neg / batch_len


def adam_step(p: np.ndarray, g: np.ndarray, m: np.ndarray, v: np.ndarray, step: int, lr: float = 1e-6,
              beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
    """torch.optim.Adam single-tensor update (no weight decay, no amsgrad); ``step`` counts from 1."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = np.sqrt(v) / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p.astype(np.float32), m.astype(np.float32), v.astype(np.float32)


def pretext_step(sd: Dict[str, torch.Tensor], batch: torch.Tensor, batch_len: int, temperature: float,
                 lr: float = 1e-6, step: int = 1, opt_state=None):
    """One zerons training batch (pretext.py:475-490) with dropout disabled.

    ``batch`` is (B, 2, 96, T): channel 0 -> anchor branch, channel 1 -> positive branch.
    Returns dict(loss, cos_pos, cos_neg, anc_emb, pos_emb, grads{key}, new_sd{key}).
    """
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    anc = batch[:, 0:1]
    pos = batch[:, 1:2]
    a, p = pretext_forward(params, anc, pos)
    loss, cp, cn = ntxent(a, p, batch_len, temperature)
    loss.backward()
    grads = {k: v.grad.detach().clone() for k, v in params.items()}
    new_sd = {}
    for k in params:
        m0 = np.zeros(params[k].shape, np.float32) if opt_state is None else opt_state[k][0]
        v0 = np.zeros(params[k].shape, np.float32) if opt_state is None else opt_state[k][1]
        pn, _, _ = adam_step(params[k].detach().numpy().astype(np.float64), grads[k].numpy().astype(np.float64),
                             m0.astype(np.float64), v0.astype(np.float64), step, lr)
        new_sd[k] = torch.from_numpy(pn)
    return dict(loss=float(loss.detach()), cos_pos=cp, cos_neg=cn, anc_emb=a.detach(), pos_emb=p.detach(),
                grads=grads, new_sd=new_sd)


def he_normal_state_dict(seed: int = 0, prefix: str = "") -> Dict[str, torch.Tensor]:
    """Synthetic checkpoint with the reference key layout (the shipped .pth is a missing blob).

    He-normal conv weights (std = sqrt(2 / fan_in)), zero biases; fc1 weight std sqrt(1/128).
    Deterministic in ``seed`` (numpy PCG64).  Identical generator lives in
    ``zeronotesamba_b200.models.checkpoint`` for the product side; the two are compared in tests.
    """
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
