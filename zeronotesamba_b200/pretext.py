"""Pretext (ZeroNS) training hot path -- drop-in for the step semantics of
/root/reference/zeroNoteSamba/pretext.py: ``train_epoch`` (:453-524), ``val_epoch`` (:527-592), the
crop sampler (:308-321) and ``Adam(lr=1e-6)`` (:202), plus what the north star adds: VQT inside
the loop on the GPU and data-parallel training over the GPUs of one node.

Two ways in:
  * ``train_epoch(model, loader, criterion, optimizer)`` / ``val_epoch(...)`` -- the reference's
    signatures and return values.  With a stock ``torch.optim`` optimizer the loop is the
    reference's (autograd through the fused kernels, ``optimizer.step()``); with this module's
    ``FusedAdam`` it runs ``PretextTrainer.step`` (whole step in one CUDA graph, no host syncs).
  * ``PretextTrainer`` -- flat fp32 parameter / gradient / moment buffers (the nn.Parameters become
    views, so ``state_dict()`` keeps the reference layout), one gradient all-reduce per step over
    NCCL when ``torch.distributed`` is initialised, fused Adam.

Bank building, pickling, plotting and Spleeter (pretext.py:30-172,418-448) are out of scope.
"""
from __future__ import annotations

import os
import random
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from . import dist_utils
from .engine import EncoderEngine, branch_param_names
from .models.loss_functions import NTXent
from .models.models import Pretext_CNN

CROP_FRAMES = 313          # pretext.py:285,312
CLIP_FRAMES = 626          # pretext.py:255-256


def sample_crop_starts(batch_len: int, rng: Optional[random.Random] = None, n_frames: int = CLIP_FRAMES,
                       crop: int = CROP_FRAMES) -> List[int]:
    """``random.sample(range(0, 313), batch_len)`` -- distinct crop starts of one clip (pretext.py:312)."""
    r = rng if rng is not None else random
    return r.sample(range(0, n_frames - crop), batch_len)


def crop_batch(vqt_pair: torch.Tensor, starts: torch.Tensor, out: Optional[torch.Tensor] = None,
               crop: int = CROP_FRAMES) -> torch.Tensor:
    """vqt_pair (2, 96, F) CUDA fp32, starts int32 CUDA (n,) -> (n, 2, 96, crop): one batch == the n
    shifts of ONE clip (pretext.py:314-321, DataLoader(shuffle=False))."""
    c, bins, frames = vqt_pair.shape
    n = starts.numel()
    if out is None:
        out = torch.empty(n, c, bins, crop, device=vqt_pair.device, dtype=torch.float32)
    L.check(L.lib().zns_crop_gather(L.ptr(vqt_pair), c, bins, frames, L.ptr(starts), n, crop, L.ptr(out), L.current_stream()))
    return out


class FusedAdam(torch.optim.Optimizer):
    """``torch.optim.Adam`` defaults (betas (0.9, 0.999), eps 1e-8, no weight decay / amsgrad) as one
    flat fused kernel.  Passing it to ``train_epoch`` selects the graph-captured trainer."""

    def __init__(self, params, lr: float = 1e-6, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._trainer: Optional["PretextTrainer"] = None

    def step(self, closure=None):  # generic path: per-tensor fused update
        assert closure is None
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                stt = self.state[p]
                if not stt:
                    stt["step"] = 0
                    stt["exp_avg"] = torch.zeros_like(p)
                    stt["exp_avg_sq"] = torch.zeros_like(p)
                stt["step"] += 1
                g = p.grad.contiguous()
                L.check(L.lib().zns_adam_flat(L.ptr(p), L.ptr(g), L.ptr(stt["exp_avg"]), L.ptr(stt["exp_avg_sq"]), p.numel(),
                                              group["lr"], b1, b2, group["eps"], stt["step"], None, 1.0, L.current_stream()))


class PretextTrainer:
    """Fused ZeroNS pretext step on one GPU (one process per GPU under torchrun)."""

    def __init__(self, model: Pretext_CNN, batch_len: int = 16, temperature: float = 0.25, lr: float = 1e-6,
                 betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8, crop_frames: int = CROP_FRAMES,
                 dropout_p: Optional[float] = None, use_graph: bool = True, seed: int = 0, distributed: bool = True,
                 p2p_adam: bool = True):
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("PretextTrainer needs the model on a CUDA device (no CPU fallback)")
        self.model, self.device = model, dev
        self.B, self.T = int(batch_len), int(crop_frames)
        self.temperature, self.lr, self.betas, self.eps = float(temperature), float(lr), betas, float(eps)
        self.dropout_p = model.anchor.pretrained.dp.p if dropout_p is None else float(dropout_p)
        self.use_graph = use_graph
        self.distributed = bool(distributed) and dist_utils.is_distributed()
        self.world = torch.distributed.get_world_size() if self.distributed else 1
        names = branch_param_names()
        named = [dict(model.anchor.named_parameters()), dict(model.postve.named_parameters())]
        plist = [named[br][n] for br in range(2) for n in names]
        # flat buffers (every segment 16-byte aligned); parameters become views
        offs, total = [], 0
        for p in plist:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4
        # Multi-GPU: parameters and gradients live in symmetric memory so that every rank can address
        # every other rank's buffers over NVLink; the optimizer is then ONE kernel that reduce-scatters
        # the gradients, applies Adam to the owned shard and all-gathers the new parameters by peer
        # stores (zns_adam_p2p).  ZNS_P2P_ADAM=0 (or a failed rendezvous) falls back to NCCL all-reduce
        # + local Adam.
        self._symm = None
        if self.distributed and p2p_adam and os.environ.get("ZNS_P2P_ADAM", "1") != "0":
            try:
                self._symm = dist_utils.SymmetricFlat(total, dev)
            except Exception as exc:   # pragma: no cover - depends on the fabric
                print(f"[zns] symmetric-memory rendezvous failed ({exc!r}); using NCCL all-reduce + local Adam")
                self._symm = None
        if self._symm is not None:
            self.flat_p, self.flat_g = self._symm.p, self._symm.g
        else:
            self.flat_p = torch.zeros(total, device=dev)
            self.flat_g = torch.zeros(total, device=dev)
        self.flat_m = torch.zeros(total, device=dev)
        self.flat_v = torch.zeros(total, device=dev)
        self.params: List[Dict[str, torch.Tensor]] = [{}, {}]
        self.grads: List[Dict[str, torch.Tensor]] = [{}, {}]
        # data-parallel buckets (element ranges of the flat buffers): per branch, "late" = cv5.weight .. fc1.bias (final after
        # the first phase of the backward pass, 80 % of the parameters), "early" = cv1.weight .. cv4.bias
        n_per = len(names)
        i5 = names.index("pretrained.cv5.weight")
        self._bucket_late, self._bucket_early = [], []
        for br in range(2):
            lo, mid = offs[br * n_per], offs[br * n_per + i5]
            hi = offs[(br + 1) * n_per] if br == 0 else total
            self._bucket_early.append((lo, mid))
            self._bucket_late.append((mid, hi))
        for i, p in enumerate(plist):
            seg = self.flat_p[offs[i]:offs[i] + p.numel()].view_as(p)
            seg.copy_(p.data)
            p.data = seg
            gseg = self.flat_g[offs[i]:offs[i] + p.numel()].view_as(p)
            p.grad = gseg
            br, n = divmod(i, len(names))
            self.params[br][names[n]] = seg
            self.grads[br][names[n]] = gseg
        if self.distributed:
            dist_utils.broadcast_parameters(self.flat_p, 0)   # replicas start from rank 0's weights
        # every data-parallel rank draws its own dropout masks
        rank = torch.distributed.get_rank() if self.distributed else 0
        self.engine = EncoderEngine(self.B, self.T, 2, dev, seed=(int(seed) + 0x9E3779B1 * rank) & 0xFFFFFFFF)
        self.engine._ensure_grad_ws()
        self.batch_buf = torch.zeros(self.B, 2, 96, self.T, device=dev)
        self.result = torch.zeros(3, device=dev)
        self.d_emb = [torch.zeros(self.B, self.T, device=dev) for _ in range(2)]
        # overlapped exchange (N > 1, peer-memory optimizer): ZNS_DP_OVERLAP=0 selects the exchange after the whole backward
        self.dp_overlap = self._symm is not None and os.environ.get("ZNS_DP_OVERLAP", "1") != "0"
        self._graph_late: Optional[torch.cuda.CUDAGraph] = None
        self._graph_early: Optional[torch.cuda.CUDAGraph] = None
        self._comm = torch.cuda.Stream(device=dev)
        self._ev_late = torch.cuda.Event()
        self._ev_comm = torch.cuda.Event()
        self._graph_fb: Optional[torch.cuda.CUDAGraph] = None
        self._graph_opt: Optional[torch.cuda.CUDAGraph] = None
        self._graph_eval: Optional[torch.cuda.CUDAGraph] = None
        self.n_launches_step = 0
        # front-end (VQT in the loop)
        self._vqt_plan = None
        self._audio_buf: Optional[torch.Tensor] = None
        self._vqt_buf: Optional[torch.Tensor] = None
        self._starts_buf = torch.zeros(self.B, dtype=torch.int32, device=dev)
        self._graph_front: Optional[torch.cuda.CUDAGraph] = None
        # the front-end writes a staging batch so that the NEXT clip's VQT + crops can run on a side
        # stream while the current step's encoders are busy (prefetch_audio / step_prefetched)
        self._stage_buf = torch.zeros_like(self.batch_buf)
        self._side = torch.cuda.Stream(device=dev)
        self._ev_ready = torch.cuda.Event()
        self._ev_consumed = torch.cuda.Event()
        self._ev_consumed.record(torch.cuda.current_stream())

    # ---- pieces ------------------------------------------------------------------------------
    def _forward_backward(self):
        lib, st = L.lib(), L.current_stream()
        eng = self.engine
        L.check(lib.zns_counter_add(L.ptr(eng.step_ctr), 1, st))
        eng.pack_weights_async(self.params, need_dgrad=True)      # side stream, joined before cv2
        L.check(lib.zns_zero(L.ptr(self.flat_g), self.flat_g.numel() * 4, st))     # memset node, not a fill kernel
        eng.forward([self.batch_buf[:, 0], self.batch_buf[:, 1]], 2 * 96 * self.T, self.params, train=True,
                    dropout_p=self.dropout_p)
        L.check(lib.zns_ntxent_fwd_bwd(L.ptr(eng.emb[0]), L.ptr(eng.emb[1]), self.B, self.T, self.B, self.temperature,
                                       L.ptr(self.result), L.ptr(self.d_emb[0]), L.ptr(self.d_emb[1]), st))
        eng.backward(self.d_emb, self.params, self.grads)

    def _optimizer_p2p(self, ranges=None):
        """Fused reduce-scatter + Adam + all-gather over peer memory for the whole flat buffer or for element ranges of it."""
        import ctypes
        sm = self._symm
        for lo, hi in (ranges if ranges is not None else [(0, self.flat_p.numel())]):
            if hi <= lo:
                continue
            g_ptrs = (ctypes.c_void_p * sm.world)(*[int(a) + 4 * lo for a in sm.g_ptrs])
            p_ptrs = (ctypes.c_void_p * sm.world)(*[int(a) + 4 * lo for a in sm.p_ptrs])
            L.check(L.lib().zns_adam_p2p(sm.world, sm.rank, g_ptrs, p_ptrs, self.flat_m.data_ptr() + 4 * lo,
                                         self.flat_v.data_ptr() + 4 * lo, hi - lo, self.lr, self.betas[0], self.betas[1],
                                         self.eps, 0, L.ptr(self.engine.step_ctr), L.current_stream()))

    # ---- overlapped data-parallel step: exchange of the late bucket under the early half of the backward pass -----------
    def _fb_late(self):
        lib, st = L.lib(), L.current_stream()
        eng = self.engine
        L.check(lib.zns_counter_add(L.ptr(eng.step_ctr), 1, st))
        eng.pack_weights_async(self.params, need_dgrad=True)
        L.check(lib.zns_zero(L.ptr(self.flat_g), self.flat_g.numel() * 4, st))
        eng.forward([self.batch_buf[:, 0], self.batch_buf[:, 1]], 2 * 96 * self.T, self.params, train=True,
                    dropout_p=self.dropout_p)
        L.check(lib.zns_ntxent_fwd_bwd(L.ptr(eng.emb[0]), L.ptr(eng.emb[1]), self.B, self.T, self.B, self.temperature,
                                       L.ptr(self.result), L.ptr(self.d_emb[0]), L.ptr(self.d_emb[1]), st))
        eng.backward(self.d_emb, self.params, self.grads, phase="late")

    def _fb_early(self):
        self.engine.backward(self.d_emb, self.params, self.grads, phase="early")

    def _step_overlapped(self):
        """late phase -> [comm stream: barrier, exchange + Adam of the late bucket, barrier]  ||  early phase -> exchange of
        the early bucket.  The backward pass reads the packed 16-bit weight copies, never the fp32 masters, so the late
        bucket's parameters may be updated (by every peer) while the early phase still runs."""
        main = torch.cuda.current_stream()
        if self.use_graph:
            if self._graph_late is None:
                snap = (self.flat_p.clone(), self.flat_m.clone(), self.flat_v.clone(), self.engine.step_ctr.clone())
                self._graph_late = self._capture(self._fb_late)
                self._graph_early = self._capture(self._fb_early)
                self.flat_p.copy_(snap[0]); self.flat_m.copy_(snap[1]); self.flat_v.copy_(snap[2])
                self.engine.step_ctr.copy_(snap[3])
            self._graph_late.replay()
        else:
            self._fb_late()
        self._ev_late.record(main)
        self._comm.wait_event(self._ev_late)
        with torch.cuda.stream(self._comm):
            self._symm.barrier(2)              # every rank's late-bucket gradients are complete
            self._optimizer_p2p(self._bucket_late)
            self._symm.barrier(3)              # every rank's stores into this rank's late-bucket parameters have landed
            self._ev_comm.record(self._comm)
        if self.use_graph:
            self._graph_early.replay()
        else:
            self._fb_early()
        self._symm.barrier(0)
        self._optimizer_p2p(self._bucket_early)
        self._symm.barrier(1)
        main.wait_event(self._ev_comm)

    def _reduce_and_update(self):
        """Gradient exchange + optimizer of one step (eager calls between the captured graphs)."""
        if self._symm is not None:
            self._symm.barrier(0)          # every rank's gradients are complete
            self._optimizer_p2p()
            self._symm.barrier(1)          # every rank's parameter stores have landed
        else:
            if self.distributed:
                dist_utils.allreduce_gradients(self.flat_g)     # NCCL over NVLink; Adam applies 1/world
            if self.use_graph and self._graph_opt is not None:
                self._graph_opt.replay()
            else:
                self._optimizer()

    def set_lr(self, lr: float) -> None:
        """Change the learning rate of the following steps (the captured optimizer graph bakes it in: dropped, the optimizer then
        runs as an eager launch behind the forward/backward graph)."""
        if float(lr) != self.lr:
            self.lr = float(lr)
            self._graph_opt = None

    def _optimizer(self):
        L.check(L.lib().zns_adam_flat(L.ptr(self.flat_p), L.ptr(self.flat_g), L.ptr(self.flat_m), L.ptr(self.flat_v),
                                      self.flat_p.numel(), self.lr, self.betas[0], self.betas[1], self.eps, 0,
                                      L.ptr(self.engine.step_ctr), 1.0 / self.world, L.current_stream()))

    def _eval_forward(self):
        lib, st = L.lib(), L.current_stream()
        eng = self.engine
        eng.pack_weights(self.params, need_dgrad=False)
        eng.forward([self.batch_buf[:, 0], self.batch_buf[:, 1]], 2 * 96 * self.T, self.params, train=False)
        L.check(lib.zns_ntxent_fwd_bwd(L.ptr(eng.emb[0]), L.ptr(eng.emb[1]), self.B, self.T, self.B, self.temperature,
                                       L.ptr(self.result), None, None, st))

    def _capture(self, fn) -> torch.cuda.CUDAGraph:
        # warm-up on a side stream (first-call attribute setting, lazy module loading), then capture
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g

    # ---- public ------------------------------------------------------------------------------
    def load_batch(self, batch: torch.Tensor) -> None:
        """batch (B, 2, 96, T): channel 0 -> anchor branch, channel 1 -> positive branch (pretext.py:476-477)."""
        if batch.data_ptr() != self.batch_buf.data_ptr():
            self.batch_buf.copy_(batch, non_blocking=True)

    def step(self, batch: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One training step (pretext.py:479-488).  Returns the device tensor
        [loss, mean cos(anchor,pos), mean cos(anchor,neg)] of this step (no host sync)."""
        if batch is not None:
            self.load_batch(batch)
        if self.dp_overlap:
            self._step_overlapped()
            return self.result
        if self.use_graph:
            if self._graph_fb is None:
                snap = (self.flat_p.clone(), self.flat_m.clone(), self.flat_v.clone(), self.engine.step_ctr.clone())
                self._graph_fb = self._capture(self._forward_backward)
                if self._symm is None:
                    self._graph_opt = self._capture(self._optimizer)
                # capture warm-ups really ran: restore the state they touched
                self.flat_p.copy_(snap[0]); self.flat_m.copy_(snap[1]); self.flat_v.copy_(snap[2])
                self.engine.step_ctr.copy_(snap[3])
            self._graph_fb.replay()
            self._reduce_and_update()
        else:
            self._forward_backward()
            self._reduce_and_update()
        return self.result

    def eval_step(self, batch: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Forward + loss only (val_epoch, pretext.py:546-564)."""
        if batch is not None:
            self.load_batch(batch)
        if self.use_graph:
            if self._graph_eval is None:
                self._graph_eval = self._capture(self._eval_forward)
            self._graph_eval.replay()
        else:
            self._eval_forward()
        return self.result

    # ---- VQT in the loop ---------------------------------------------------------------------------
    def _front(self):
        self._vqt_plan.forward(self._audio_buf, out=self._vqt_buf)
        crop_batch(self._vqt_buf, self._starts_buf, out=self._stage_buf, crop=self.T)

    def _ensure_front(self, n: int, sample_rate: int, mode: str):
        from .processing.input_rep import VQTPlan
        if self._vqt_plan is None or self._audio_buf is None or self._audio_buf.shape[1] != n:
            self._vqt_plan = VQTPlan(sample_rate, mode, 2, n)
            self._audio_buf = torch.zeros(2, n, device=self.device)
            self._vqt_buf = torch.zeros(2, 96, self._vqt_plan.frames(n), device=self.device)
            self._graph_front = None

    def _run_front(self, anchor_audio, positive_audio, starts):
        self._audio_buf[0].copy_(anchor_audio, non_blocking=True)
        self._audio_buf[1].copy_(positive_audio, non_blocking=True)
        self._starts_buf.copy_(starts, non_blocking=True)
        if self.use_graph:
            if self._graph_front is None:
                self._graph_front = self._capture(self._front)
            self._graph_front.replay()
        else:
            self._front()

    def prefetch_audio(self, anchor_audio: torch.Tensor, positive_audio: torch.Tensor, starts: torch.Tensor,
                       sample_rate: int = 16000, mode: str = "vqt") -> None:
        """Start the front-end (H2D/D2D copies, VQT, crops) of the NEXT source clip on a side stream; it
        overlaps whatever the current stream is doing (typically the previous clip's training step).
        Consume it with ``step_prefetched()``."""
        self._ensure_front(anchor_audio.numel(), sample_rate, mode)
        if self.use_graph and self._graph_front is None:
            self._run_front(anchor_audio, positive_audio, starts)      # first call captures on the current stream
            self._ev_ready.record(torch.cuda.current_stream())
            return
        self._side.wait_event(self._ev_consumed)       # the staging batch of the previous prefetch has been taken
        cuda_inputs = [t for t in (anchor_audio, positive_audio, starts) if t.is_cuda]
        if cuda_inputs:
            # device inputs may still be in flight on the caller's stream, and the caller may free them right after this call
            self._side.wait_stream(torch.cuda.current_stream())
            for t in cuda_inputs:
                t.record_stream(self._side)
        with torch.cuda.stream(self._side):
            self._run_front(anchor_audio, positive_audio, starts)
            self._ev_ready.record(self._side)

    def step_prefetched(self) -> torch.Tensor:
        """Training step on the clip handed to the last ``prefetch_audio``."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._ev_ready)
        self.batch_buf.copy_(self._stage_buf, non_blocking=True)
        self._ev_consumed.record(cur)
        return self.step()

    def step_from_audio(self, anchor_audio: torch.Tensor, positive_audio: torch.Tensor, starts: torch.Tensor,
                        sample_rate: int = 16000, mode: str = "vqt", run_step: bool = True) -> torch.Tensor:
        """Whole in-loop path on the device: two 16 kHz stems of one source clip (anchor = other stems,
        positive = drums; pretext.py:83-84,144) -> VQT (2, 96, F) -> B crops at ``starts`` -> training step."""
        self._ensure_front(anchor_audio.numel(), sample_rate, mode)
        self._run_front(anchor_audio, positive_audio, starts)
        self.batch_buf.copy_(self._stage_buf, non_blocking=True)
        return self.step() if run_step else self.result


# ---------------------------------------------------------------------------------------------------
# reference-signature epoch functions
# ---------------------------------------------------------------------------------------------------
def _trainer_for(model: Pretext_CNN, criterion: NTXent, optimizer: FusedAdam, T: int) -> PretextTrainer:
    """The graph-captured trainer behind ``train_epoch(..., FusedAdam)``.  It is rebuilt when the batch geometry changes; the
    Adam moments and the step count then move over to the new trainer (same model), and a learning rate changed in
    ``optimizer.param_groups`` (a scheduler) is picked up before every epoch."""
    tr = optimizer._trainer
    g = optimizer.param_groups[0]
    if tr is None or tr.model is not model or tr.T != T or tr.B != criterion.batch_len:
        old = tr
        tr = PretextTrainer(model, batch_len=criterion.batch_len, temperature=criterion.temperature, lr=g["lr"],
                            betas=g["betas"], eps=g["eps"], crop_frames=T)
        if old is not None and old.model is model and old.flat_m.numel() == tr.flat_m.numel():
            tr.flat_m.copy_(old.flat_m)
            tr.flat_v.copy_(old.flat_v)
            tr.engine.step_ctr.copy_(old.engine.step_ctr)
        optimizer._trainer = tr
    tr.set_lr(g["lr"])
    return tr


def train_epoch(model: torch.nn.Module, train_loader: Iterable, criterion: NTXent, optimizer: torch.optim.Optimizer,
                pt_task: str = "zerons") -> Tuple[torch.nn.Module, float, float, float]:
    """
    Function for CL model training (reference signature and return values, pretext.py:453-524).
    -- model: model to train
    -- train_loader: loader with batches that contain 1 anchor, 1 positive, and negatives
    -- criterion: loss function
    -- optimizer: optimizer defined
    -- pt_task: pretext task to run
    """
    if pt_task not in ("zerons", "clmr"):
        raise ValueError("Which pretext task are we running?")
    device = next(model.parameters()).device
    model.train()
    n_batches = 0
    if pt_task == "clmr":
        # CLMR baseline (pretext.py:494-511): ONE DS_CNN applied to both views, shared weights
        full_train_loss = full_train_anpos = full_train_anneg = 0.0
        for [batch] in train_loader:
            anchors = batch[:, 0:1, :, :].to(device)
            postves = batch[:, 1:2, :, :].to(device)
            optimizer.zero_grad()
            anc_emb, pos_emb = model.forward_pair(anchors, postves)
            loss, sim_an_pos, sim_an_neg = criterion(anc_emb, pos_emb)
            loss.backward()
            optimizer.step()
            full_train_loss += loss.item()
            full_train_anpos += sim_an_pos
            full_train_anneg += sim_an_neg
            n_batches += 1
        full_train_loss /= n_batches
        full_train_anpos /= n_batches
        full_train_anneg /= n_batches
    elif isinstance(optimizer, FusedAdam) and isinstance(model, Pretext_CNN):
        acc = torch.zeros(3, device=device)
        for [batch] in train_loader:
            if batch.shape[0] != criterion.batch_len:
                raise ValueError("fused trainer needs full batches (the reference's loader always yields batch_len crops)")
            tr = _trainer_for(model, criterion, optimizer, batch.shape[3])
            acc += tr.step(batch.to(device, non_blocking=True))
            n_batches += 1
        full_train_loss, full_train_anpos, full_train_anneg = (acc / max(n_batches, 1)).tolist()
    else:
        full_train_loss = full_train_anpos = full_train_anneg = 0.0
        for [batch] in train_loader:
            anchors = batch[:, 0:1, :, :].to(device)
            postves = batch[:, 1:2, :, :].to(device)
            optimizer.zero_grad()
            anc_emb, pos_emb = model(anchors, postves)
            loss, sim_an_pos, sim_an_neg = criterion(anc_emb, pos_emb)
            loss.backward()
            optimizer.step()
            full_train_loss += loss.item()
            full_train_anpos += sim_an_pos
            full_train_anneg += sim_an_neg
            n_batches += 1
        full_train_loss /= n_batches
        full_train_anpos /= n_batches
        full_train_anneg /= n_batches
    print("*** Mean training batch loss is {:.3f}.".format(full_train_loss))
    print("*** Mean training anchor / positive similiarity is {:.3f}.".format(full_train_anpos))
    print("*** Mean training anchor / negative similiarity is {:.3f}.".format(full_train_anneg))
    return model, full_train_loss, full_train_anpos, full_train_anneg


def val_epoch(model: torch.nn.Module, val_loader: Iterable, criterion: NTXent, optimizer: torch.optim.Optimizer,
              pt_task: str = "zerons") -> Tuple[float, float, float]:
    """
    Validation pass (reference signature and return values, pretext.py:527-592).
    """
    if pt_task not in ("zerons", "clmr"):
        raise ValueError("Which pretext task are we running?")
    device = next(model.parameters()).device
    model.eval()
    full_val_loss = full_val_anpos = full_val_anneg = 0.0
    n_batches = 0
    for [batch] in val_loader:
        with torch.no_grad():
            anchors = batch[:, 0:1, :, :].to(device)
            postves = batch[:, 1:2, :, :].to(device)
            anc_emb, pos_emb = model(anchors, postves) if pt_task == "zerons" else model.forward_pair(anchors, postves)
            loss, sim_an_pos, sim_an_neg = criterion(anc_emb, pos_emb)
            full_val_loss += loss.item()
            full_val_anpos += sim_an_pos
            full_val_anneg += sim_an_neg
            n_batches += 1
    return full_val_loss / n_batches, full_val_anpos / n_batches, full_val_anneg / n_batches


def build_from_config(ymldict: Dict, device: Optional[torch.device] = None):
    """Model / criterion / optimizer as ``train_model`` builds them (pretext.py:185-202), reading the
    same YAML keys (``batch_size``, ``temp``, ``pt_task``) plus ``lr`` (which the reference ignores)."""
    batch_len = int(float(ymldict.get("batch_size", -1)))
    tmp = float(ymldict.get("temp", -1.0))
    pt_task = ymldict.get("pt_task")
    lr = float(ymldict.get("lr", 0.000001))
    device = device or torch.device("cuda", torch.cuda.current_device())
    criterion = NTXent(batch_len=batch_len, temperature=tmp)
    if pt_task == "zerons":
        model = Pretext_CNN().to(device)
        optimizer = FusedAdam(model.parameters(), lr=lr)
    elif pt_task == "clmr":
        from .models.models import DS_CNN
        model = DS_CNN().to(device)
        optimizer = FusedAdam(model.parameters(), lr=0.00001)     # pretext.py:208
    else:
        raise ValueError("Which pretext task are we running?")
    return model, criterion, optimizer


def vqt_bank(anchor_audio: torch.Tensor, positive_audio: torch.Tensor, sample_rate: int = 16000,
             mode: str = "vqt") -> torch.Tensor:
    """On-GPU replacement of the pickled banks of pretext.py:89-172: anchor / positive stems
    (CUDA fp32 [n, N]) -> bank (n, 2, 96, 1 + N//256), channel 0 = anchor, 1 = positive (pretext.py:144-145)."""
    from .processing.input_rep import xqt_batch
    a = xqt_batch(anchor_audio.contiguous().float(), sample_rate, mode)
    p = xqt_batch(positive_audio.contiguous().float(), sample_rate, mode)
    return torch.stack([a, p], dim=1)


def crop_batches(bank: torch.Tensor, batch_len: int, rng: Optional[random.Random] = None, crop: int = CROP_FRAMES):
    """Index-only crop sampler over a bank (pretext.py:308-321): yields ``[batch]`` like the reference's
    DataLoader -- one batch per source clip, ``batch_len`` distinct random starts -- without ever
    materialising the (1440*16, 2, 96, 313) staging array."""
    n, _, _, frames = bank.shape
    for i in range(n):
        starts = sample_crop_starts(batch_len, rng, n_frames=frames, crop=crop)
        st = torch.tensor(starts, dtype=torch.int32, device=bank.device)
        yield [crop_batch(bank[i], st, crop=crop)]


# ---------------------------------------------------------------------------------------------------
# train_model: the epoch driver of pretext.py:175-415 on device-resident banks
# ---------------------------------------------------------------------------------------------------
def _rank_world() -> Tuple[int, int]:
    if dist_utils.is_distributed():
        return torch.distributed.get_rank(), torch.distributed.get_world_size()
    return 0, 1


def _mean_over_ranks(values: Sequence[float], device: torch.device) -> List[float]:
    """Average per-rank epoch statistics (every rank saw an equal share of the clips)."""
    if not dist_utils.is_distributed():
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    torch.distributed.all_reduce(t)
    return (t / torch.distributed.get_world_size()).tolist()


def train_model(ymldict: Dict, train_bank: torch.Tensor, val_bank: torch.Tensor, model_dir: str = "models",
                chunks_per_epoch: int = 20, val_chunks: int = 10, seed: int = 0, model: Optional[torch.nn.Module] = None,
                verbose: bool = True):
    """Pretext training driver with the reference's epoch structure (pretext.py:175-415) and YAML keys (``batch_size``,
    ``num_epochs``, ``temp``, ``pt_task``; ``lr`` is honoured), on banks that already live on the GPU.

    ``train_bank`` / ``val_bank``: CUDA fp32 (n, 2, 96, F) log-VQT banks, channel 0 = anchor stems, 1 = drums
    (``vqt_bank`` builds them from audio; the reference unpickles 28800 / 6400 clips of 626 frames, pretext.py:255-263).

    Per epoch, as the reference does: shuffle the training clips, walk them in ``chunks_per_epoch`` chunks, draw
    ``batch_size`` distinct crop starts per clip so that one batch is the shifts of one clip (pretext.py:308-321),
    ``train_epoch`` per chunk, average; validation crops are drawn once at epoch 0 (pretext.py:271-279) and reused;
    the ``state_dict`` is saved to ``{model_dir}/shift_pret_cnn_{batch_size}.pth`` (``clmr_pret_cnn_...`` for CLMR)
    whenever the validation loss improves (pretext.py:409-412).

    Data parallel (one process per GPU, torch.distributed initialised): every rank holds the same banks and takes the
    clips ``rank::world`` of the epoch's permutation (drawn from a seed shared by all ranks), statistics are averaged
    over ranks, and rank 0 alone writes the checkpoint.  Clips a chunk cannot deal evenly to the ranks are dropped for
    that epoch (the fused step is collective).
    Returns (model, history) with history = dict of per-epoch lists."""
    epochs = int(float(ymldict.get("num_epochs", -1)))
    batch_len = int(float(ymldict.get("batch_size", -1)))
    pt_task = ymldict.get("pt_task")
    if pt_task not in ("zerons", "clmr"):
        raise ValueError("Which pretext task are we running?")
    if train_bank.device.type != "cuda" or val_bank.device.type != "cuda":
        raise RuntimeError("train_model needs the banks on a CUDA device (no CPU fallback)")
    device = train_bank.device
    built, criterion, optimizer = build_from_config(ymldict, device)
    if model is None:
        model = built
    else:                                                   # caller-provided weights (resume): same optimizer family
        optimizer = FusedAdam(model.parameters(), lr=optimizer.param_groups[0]["lr"])
    model_name = ("shift_pret_cnn_{}.pth" if pt_task == "zerons" else "clmr_pret_cnn_{}.pth").format(batch_len)
    rank, world = _rank_world()
    crop = CROP_FRAMES if train_bank.shape[3] > CROP_FRAMES else train_bank.shape[3]
    shared = random.Random(seed)                            # same stream on every rank: permutations
    local = random.Random(seed * 7919 + 1 + rank)           # per rank: crop starts
    history = dict(train_loss=[], train_an_pos=[], train_an_neg=[], val_loss=[], val_an_pos=[], val_an_neg=[], saved=[])
    best_val = float("inf")
    say = print if (verbose and rank == 0) else (lambda *a, **k: None)

    def my_share(indices: List[int]) -> List[int]:
        return [indices[k] for k in dist_utils.shard_clips(len(indices), rank, world)]

    # validation crops: drawn once, kept as start indices (the reference materialises 6400 x batch_len crops)
    val_ids = my_share(list(range(val_bank.shape[0])))
    if crop == val_bank.shape[3]:
        val_starts = {i: [0] * batch_len for i in val_ids}
    else:
        val_starts = {i: sample_crop_starts(batch_len, local, n_frames=val_bank.shape[3], crop=crop) for i in val_ids}

    def batches(bank, ids, starts_of):
        for i in ids:
            st = torch.tensor(starts_of(i), dtype=torch.int32, device=device)
            yield [crop_batch(bank[i], st, crop=crop)]

    def draw_train(_i):
        if crop == train_bank.shape[3]:
            return [0] * batch_len
        return sample_crop_starts(batch_len, local, n_frames=train_bank.shape[3], crop=crop)

    for epoch in range(epochs):
        say("\n--- Epoch {} ---\n".format(epoch))
        order = list(range(train_bank.shape[0]))
        shared.shuffle(order)
        n_chunks = max(1, min(chunks_per_epoch, len(order) // max(world, 1)))
        per_chunk = len(order) // n_chunks
        sums = [0.0, 0.0, 0.0]
        for jj in range(n_chunks):
            ids = my_share(order[jj * per_chunk:(jj + 1) * per_chunk])
            say("{} : Training...".format(jj))
            model, t_loss, t_pos, t_neg = train_epoch(model, batches(train_bank, ids, draw_train), criterion, optimizer,
                                                      pt_task=pt_task)
            sums = [a + b for a, b in zip(sums, (t_loss, t_pos, t_neg))]
        tr_stats = _mean_over_ranks([s / n_chunks for s in sums], device)
        say("\n!!! Mean training batch loss is {:.3f}.".format(tr_stats[0]))
        say("!!! Mean training anchor / positive similiarity is {:.3f}.".format(tr_stats[1]))
        say("!!! Mean training anchor / negative similiarity is {:.3f}.".format(tr_stats[2]))
        say("\n{} : Validating...".format(epoch))
        n_vc = max(1, min(val_chunks, len(val_ids)))
        per_vc = len(val_ids) // n_vc
        vsums = [0.0, 0.0, 0.0]
        for zz in range(n_vc):
            ids = val_ids[zz * per_vc:(zz + 1) * per_vc]
            v = val_epoch(model, batches(val_bank, ids, lambda i: val_starts[i]), criterion, optimizer, pt_task=pt_task)
            vsums = [a + b for a, b in zip(vsums, v)]
        va_stats = _mean_over_ranks([s / n_vc for s in vsums], device)
        say("\n!!! Mean validation batch loss is {:.3f}.".format(va_stats[0]))
        say("!!! Mean validation anchor / positive similiarity is {:.3f}.".format(va_stats[1]))
        say("!!! Mean validation anchor / negative similiarity is {:.3f}.".format(va_stats[2]))
        improved = va_stats[0] < best_val
        if improved:
            best_val = va_stats[0]
            if rank == 0:                                   # replicas are identical: one writer
                os.makedirs(model_dir, exist_ok=True)
                path = os.path.join(model_dir, model_name)
                torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, path)
                say("...Saved model to " + path)
        for k, v in zip(("train_loss", "train_an_pos", "train_an_neg", "val_loss", "val_an_pos", "val_an_neg", "saved"),
                        (*tr_stats, *va_stats, improved)):
            history[k].append(v)
    return model, history
