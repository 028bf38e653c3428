"""Drop-in for /root/reference/zeroNoteSamba/models/models.py (lines 7-150).

Same classes, constructor arguments, forward signatures, parameter names and shapes, so
``Pretext_CNN().state_dict()`` has the reference's 36 keys and ``model.pretext.load_state_dict(
torch.load("shift_pret_cnn_16.pth"))`` works unchanged (sample_script.py:41-42, loader.py:25-27).
The layers are containers for the fp32 master weights only: the arithmetic is libzns_sm100
(cv1 SIMT, cv2..cv8 tcgen05 implicit GEMM, fused pool/ReLU/dropout, fc1+sigmoid head) driven by
``EncoderEngine``; bf16 operands, fp32 accumulation.  CUDA tensors only -- there is no CPU path.
"""
from __future__ import annotations

from typing import Any, Dict, List, Tuple

import torch
import torch.nn as nn

from .. import _lib as L
from ..engine import CONV_SPECS, EncoderEngine, branch_param_names

DROPOUT_P = 0.1   # models.py:30


def _require_cuda(x: torch.Tensor) -> None:
    if not x.is_cuda:
        raise RuntimeError("zeronotesamba_b200 models run on a CUDA device only (no CPU fallback); "
                           "move the model and its inputs to cuda first")


class _EngineCache:
    """EncoderEngines per (batch, n_br, device), kept on the owning module and shared by every clip length: an engine's
    workspaces are sized for the longest clip seen so far (plus 25 % head room when they have to grow) and re-viewed for
    shorter ones, so a dataset with variable T (one file per step, epochs.py:45-63) does not reallocate per length.

    Activations live in the engine's workspaces, not in autograd-saved tensors, so an engine whose forward still awaits its
    backward (``_pending``) must not run another forward.  A key therefore holds a small POOL: ``anc = model(a); pos =
    model(p); loss.backward()`` on one module (the reference's CLMR loop, pretext.py:503-507) takes two engines.  At most
    ``MAX_PENDING`` forwards of a geometry may be outstanding; one more recycles the oldest, whose backward then raises.
    Engines of a pool draw different dropout masks (seed = pool index)."""

    MAX_PENDING = 3

    def __init__(self):
        self._engines: Dict[Tuple[int, int, int], List[EncoderEngine]] = {}

    def get(self, batch: int, T: int, n_br: int, device) -> EncoderEngine:
        key = (batch, n_br, device.index if device.index is not None else torch.cuda.current_device())
        pool = self._engines.get(key)
        if pool is None:
            if len(self._engines) >= 4:                   # bounded number of (batch, branches) geometries
                self._engines.pop(next(iter(self._engines)))
            pool = self._engines[key] = []
        slot = next((i for i, e in enumerate(pool) if not getattr(e, "_pending", False)), None)
        if slot is None:
            if len(pool) < self.MAX_PENDING:
                slot = len(pool)
                pool.append(None)
            else:                                         # recycle the engine of the oldest outstanding forward
                slot = min(range(len(pool)), key=lambda i: getattr(pool[i], "_version", 0))
        eng = pool[slot]
        if eng is None or T > eng.T_cap:
            cap = T if eng is None else max(T, int(1.25 * eng.T_cap))
            version = getattr(eng, "_version", 0)
            pool[slot] = None
            del eng
            eng = pool[slot] = EncoderEngine(batch, cap, n_br, device, seed=0x9E3779B1 * slot & 0x7FFFFFFF)
            eng._version = version
        eng._pending = False
        eng.set_T(T)
        return eng


# Cumulative time-axis receptive-field halo of the eight convolutions: sum of (kw - 1) / 2 over
# kw = 11, 13, 15, 17, 19, 21, 23, 25 (models.py:16-23).
TIME_HALO = 68
FOLD_SLOTS = 8


def fold_plan(T: int, slots: int = FOLD_SLOTS, halo: int = TIME_HALO):
    """Time folding for batch-1 inputs (the reference fine-tunes and infers one file at a time,
    epochs.py:45-63, sample_script.py:46-48).  The act layout always carries eight clip slots; with a
    single clip seven would be padding.  Instead the T frames are cut into eight overlapping segments
    of W_s frames at hop delta (7 * delta + W_s == T, W_s - delta >= 2 * halo): every output frame is
    taken from a segment in which it lies >= `halo` frames away from an artificial edge, so its whole
    receptive field -- forward and backward -- is real data and the result equals the unfolded
    computation; true clip edges coincide with segment edges, where zero "same" padding is right.
    Returns (W_s, delta, cuts) with cuts[s]..cuts[s+1] the output frames segment s owns, or None when
    the clip is too short to fold."""
    w_min = -(-(T + (slots - 1) * 2 * halo) // slots)
    for w_s in range(w_min, w_min + slots):
        if w_s >= T:
            break
        if (T - w_s) % (slots - 1) == 0:
            delta = (T - w_s) // (slots - 1)
            if delta > 0 and w_s - delta >= 2 * halo:
                cuts = [0] + [(s + 1) * delta + halo for s in range(slots - 1)] + [T]
                return w_s, delta, cuts
    return None


def _unfold(emb: torch.Tensor, plan, T: int) -> torch.Tensor:
    _, delta, cuts = plan
    out = torch.empty(1, T, device=emb.device, dtype=emb.dtype)
    for s in range(FOLD_SLOTS):
        out[0, cuts[s]:cuts[s + 1]] = emb[s, cuts[s] - s * delta:cuts[s + 1] - s * delta]
    return out


def _fold_grad(d_full: torch.Tensor, plan) -> torch.Tensor:
    w_s, delta, cuts = plan
    out = torch.zeros(FOLD_SLOTS, w_s, device=d_full.device, dtype=torch.float32)
    for s in range(FOLD_SLOTS):
        out[s, cuts[s] - s * delta:cuts[s + 1] - s * delta] = d_full[0, cuts[s]:cuts[s + 1]]
    return out


def _check_input(x: torch.Tensor) -> Tuple[int, int]:
    _require_cuda(x)
    if x.dim() != 4 or x.shape[1] != 1 or x.shape[2] != 96:
        raise ValueError(f"expected input of shape (B, 1, 96, T), got {tuple(x.shape)}")
    return x.shape[0], x.shape[3]


class _EncoderFunction(torch.autograd.Function):
    """n_br DS_CNN branches: (x_0 [, x_1], *params) -> emb_0 [, emb_1].  No gradient w.r.t. inputs."""
    _clock = 0

    @staticmethod
    def forward(ctx, cache: _EngineCache, n_br: int, train: bool, dropout_p: float, grad_on: bool, *tensors):
        xs = [t.contiguous().float() for t in tensors[:n_br]]
        flat = tensors[n_br:]
        names = branch_param_names()
        params = [dict(zip(names, flat[br * len(names):(br + 1) * len(names)])) for br in range(n_br)]
        B, T = xs[0].shape[0], xs[0].shape[3]
        plan = fold_plan(T) if B == 1 else None
        if plan is not None:
            eng = cache.get(FOLD_SLOTS, plan[0], n_br, xs[0].device)
        else:
            eng = cache.get(B, T, n_br, xs[0].device)
        # a backward may follow (train or eval mode).  grad_on = the caller's torch.is_grad_enabled(): inside forward() grad mode is
        # always off, and under no_grad the parameters still "need" gradients although no backward will ever release the engine
        need_grad = bool(grad_on) and any(ctx.needs_input_grad[5 + n_br:])
        eng.pack_weights(params, need_dgrad=need_grad)
        if train and dropout_p > 0:
            # the keep masks are hash(element, seed ^ step counter, layer): a new counter value per training forward
            # gives a fresh mask (reference: nn.Dropout draws from the global generator, models.py:30)
            L.check(L.lib().zns_counter_add(L.ptr(eng.step_ctr), 1, L.current_stream()))
        if plan is not None:
            embs = eng.forward(xs, plan[1], params, train=train, dropout_p=dropout_p, x_row_stride=T, need_grad=need_grad)
            out = tuple(_unfold(e, plan, T) for e in embs)
        else:
            embs = eng.forward(xs, 96 * T, params, train=train, dropout_p=dropout_p, need_grad=need_grad)
            out = tuple(e.clone() for e in embs)
        ctx.eng, ctx.params, ctx.n_br, ctx.n_names, ctx.plan = eng, params, n_br, len(names), plan
        _EncoderFunction._clock += 1                           # process-wide order of forwards (oldest-first recycling)
        ctx.version = eng._version = _EncoderFunction._clock
        eng._pending = bool(need_grad)                         # the workspaces are reserved until this forward's backward
        return out if n_br > 1 else out[0]

    @staticmethod
    def backward(ctx, *d_embs):
        eng: EncoderEngine = ctx.eng
        if eng._version != ctx.version:
            raise RuntimeError("encoder workspaces were overwritten by a later forward of the same geometry (more than "
                               f"{_EngineCache.MAX_PENDING} forwards outstanding); call backward before further forwards")
        names = branch_param_names()
        grads = [{n: torch.zeros_like(ctx.params[br][n]) for n in names} for br in range(ctx.n_br)]
        if ctx.plan is not None:
            d = [_fold_grad(g.float(), ctx.plan) if g is not None else torch.zeros_like(eng.emb[i]) for i, g in enumerate(d_embs)]
        else:
            d = [(g if g is not None else torch.zeros_like(eng.emb[i])).contiguous().float() for i, g in enumerate(d_embs)]
        eng.backward(d, ctx.params, grads)
        eng._pending = False
        flat = [grads[br][n] for br in range(ctx.n_br) for n in names]
        return (None, None, None, None, None) + (None,) * ctx.n_br + tuple(flat)


class _CNN(nn.Module):
    """
    Convolutional layers (parameter container; models.py:7-74).
    """

    def __init__(self) -> None:
        super(_CNN, self).__init__()
        for name, co, ci, kh, kw, _ in CONV_SPECS:
            setattr(self, name, nn.Conv2d(in_channels=ci, out_channels=co, kernel_size=(kh, kw), padding=(kh // 2, kw // 2)))
        self.relu = nn.ReLU(inplace=True)
        self.maxpl1 = nn.MaxPool2d((3, 1), padding=(0, 0))
        self.maxpl2 = nn.MaxPool2d((4, 1), padding=(0, 0))
        self.maxpl3 = nn.MaxPool2d((8, 1), padding=(0, 0))
        self.dp = nn.Dropout(p=DROPOUT_P)
        self._cache = _EngineCache()

    def forward(self, x: torch.Tensor) -> Any:
        """
        Pass input through convolutional layers: (B, 1, 96, T) -> (B, 128, T).  Forward only
        (inference / feature extraction); training goes through DS_CNN / Pretext_CNN / Down_CNN.
        -- x: input (vqt)
        """
        B, T = _check_input(x)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("_CNN.forward on its own is forward-only; wrap it in torch.no_grad() "
                               "or train through DS_CNN / Pretext_CNN / Down_CNN")
        xs = [x.contiguous().float()]
        names = branch_param_names()[:-2]
        params = [{n: dict(self.named_parameters())[n.replace("pretrained.", "")] for n in names}]
        zero_fc = torch.zeros(1, 128, 1, device=x.device), torch.zeros(1, device=x.device)
        params[0]["fc1.weight"], params[0]["fc1.bias"] = zero_fc
        eng = self._cache.get(B, T, 1, x.device)
        eng.pack_weights(params, need_dgrad=False)
        eng.forward(xs, 96 * T, params, train=self.training, dropout_p=self.dp.p, need_grad=False)
        out = torch.empty(B, 128, 1, T, device=x.device)
        L.check(L.lib().zns_act_to_nchw(L.ptr(eng.x8[0]), L.ptr(out), B, 128, 1, T, 1, L.current_stream()))
        return torch.squeeze(out, dim=2)


class DS_CNN(nn.Module):
    """
    Fully-convolutional architecture for beat tracking (models.py:77-103).
    """

    def __init__(self) -> None:
        super(DS_CNN, self).__init__()
        self.pretrained = _CNN()
        # Output
        self.fc1 = nn.Conv1d(in_channels=128, out_channels=1, kernel_size=1, padding=0)
        self.sig = nn.Sigmoid()
        self._cache = _EngineCache()
        self._pair_cache = _EngineCache()

    def _flat_params(self) -> List[torch.Tensor]:
        named = dict(self.named_parameters())
        return [named[n] for n in branch_param_names()]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """
        Pass input through the encoder: (B, 1, 96, T) -> (B, T) sigmoid activations.
        -- x: input (vqt)
        """
        _check_input(x)
        return _EncoderFunction.apply(self._cache, 1, self.training, self.pretrained.dp.p, torch.is_grad_enabled(), x,
                                      *self._flat_params())

    def forward_pair(self, a: torch.Tensor, p: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """``(self(a), self(p))`` in one pass with SHARED weights -- the CLMR baseline applies one DS_CNN
        to both views (pretext.py:503-504); both views go through every tensor-core launch together and
        autograd sums the two gradient contributions of each parameter."""
        ba, ta = _check_input(a)
        bp, tp = _check_input(p)
        if (ba, ta) != (bp, tp):
            raise ValueError("forward_pair needs two inputs of the same shape")
        params = self._flat_params()
        return _EncoderFunction.apply(self._pair_cache, 2, self.training, self.pretrained.dp.p, torch.is_grad_enabled(), a, p,
                                      *params, *params)


class Pretext_CNN(nn.Module):
    """
    DS_CNN tailored for percussive and non-percussive stems (models.py:106-124).
    """

    def __init__(self) -> None:
        super(Pretext_CNN, self).__init__()
        self.anchor = DS_CNN()
        self.postve = DS_CNN()
        self._cache = _EngineCache()

    def forward(self, anc: torch.Tensor, pos: torch.Tensor) -> Tuple[Any, Any]:
        """
        Pass vqts through each model (both branches share every tensor-core launch).
        """
        ba, ta = _check_input(anc)
        bp, tp = _check_input(pos)
        if (ba, ta) != (bp, tp):
            return self.anchor(anc), self.postve(pos)
        train = self.anchor.training
        anc_emb, pos_emb = _EncoderFunction.apply(self._cache, 2, train, self.anchor.pretrained.dp.p, torch.is_grad_enabled(), anc, pos,
                                                  *self.anchor._flat_params(), *self.postve._flat_params())
        return anc_emb, pos_emb


class Down_CNN(nn.Module):
    """
    Use of Pretext_CNN for downstream tasks (models.py:127-150).
    """

    def __init__(self, reduction: str = "max") -> None:
        super(Down_CNN, self).__init__()
        self.pretext = Pretext_CNN()
        self.reduction = reduction

    def forward(self, anc: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
        """
        Pass each input through each model. Merge (maximum, or mean) and output.
        """
        anc_emb, pos_emb = self.pretext(anc, pos)
        if anc_emb.requires_grad or pos_emb.requires_grad:
            # autograd needs the merge on the tape
            return torch.div(anc_emb + pos_emb, 2) if self.reduction == "mean" else torch.maximum(anc_emb, pos_emb)
        emb = torch.empty_like(anc_emb)
        L.check(L.lib().zns_merge(L.ptr(anc_emb), L.ptr(pos_emb), L.ptr(emb), emb.numel(),
                                  1 if self.reduction == "mean" else 0, L.current_stream()))
        return emb
