"""Encoder engine: owns the act-layout workspaces of up to two DS_CNN branches and sequences the
libzns_sm100 kernels for forward and backward.

Mirrors ``_CNN.forward`` + ``DS_CNN.forward`` (/root/reference/zeroNoteSamba/models/models.py:32-74,
93-103) layer by layer; both branches of ``Pretext_CNN`` (models.py:114-124) go through every
tensor-core launch together (``n_br = 2``).  Host code only issues launches on the current CUDA
stream: no synchronisation, no allocation after construction, so a whole step can be captured in a
CUDA graph.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib as L

# (name, c_out, c_in, kh, kw, pool)  -- models.py:16-28
CONV_SPECS = (
    ("cv1", 64, 1, 3, 11, 1),
    ("cv2", 64, 64, 7, 13, 3),
    ("cv3", 128, 64, 5, 15, 1),
    ("cv4", 128, 128, 9, 17, 4),
    ("cv5", 256, 128, 3, 19, 1),
    ("cv6", 256, 256, 5, 21, 8),
    ("cv7", 128, 256, 1, 23, 1),
    ("cv8", 128, 128, 1, 25, 1),
)
N_BINS = 96


def branch_param_names() -> List[str]:
    """Parameter names of one DS_CNN in registration (= state_dict) order."""
    names = []
    for name, *_ in CONV_SPECS:
        names += [f"pretrained.{name}.weight", f"pretrained.{name}.bias"]
    return names + ["fc1.weight", "fc1.bias"]


def _fwd_tag(c_out: int, H: int) -> str:
    """Which tensor-core kernel zns_conv_fwd dispatches to (csrc/conv_umma.cu)."""
    import os
    if os.environ.get("ZNS_CONV_TRANSPOSED") and (c_out == 128 or (c_out == 64 and H % 2 == 0)):
        return f"conv_fwdT_umma(c_out={c_out})"
    if c_out == 64 and H % 2 == 0 and not os.environ.get("ZNS_CONV_NO_STACK"):
        return "conv_fwd_stack_umma(c_out=64, 2 rows on N)"
    return f"conv_fwd_umma<{c_out}>"


class EncoderEngine:
    """Workspaces + launch sequences for ``n_br`` encoders of ``batch`` clips with up to ``T`` frames.

    The act layout [G][H][T][8][C] is compact for any frame count, so the buffers are flat storages sized for the
    capacity ``T_cap`` and ``set_T`` re-views their prefix for a shorter clip: a dataset with variable lengths (the
    downstream loops feed one file per step, epochs.py:45-63) keeps ONE engine instead of reallocating GBs per length."""

    # (attribute, rows H, channels) of the forward activations; *_s = bf16 shadow wanted (x operand of a weight gradient)
    _ACTS = (("x1", 96, 64, True),      # dropout(relu(cv1))
             ("p2", 32, 64, True),      # dropout(relu(pool3(cv2))): the pool is fused into cv2's epilogue, the pre-pool tensor is never written
             ("x3", 32, 128, True),
             ("p4", 8, 128, True),      # dropout(relu(pool4(cv4))), fused likewise
             ("x5", 8, 256, True),
             ("y6", 8, 256, False),     # cv6 pre-pool (pool 8 > the two 256-column accumulators of a tile: separate pool kernel)
             ("p6", 1, 256, True),
             ("x7", 1, 128, True),
             ("x8", 1, 128, False))
    # row of the first maximum of every pool window (one byte per pooled element): the backward pass routes through it
    _ARGS = (("arg2", 32, 64), ("arg4", 8, 128))

    def __init__(self, batch: int, T: int, n_br: int, device: torch.device, seed: int = 0):
        assert n_br in (1, 2)
        L.check(L.lib().zns_device_check())
        self.B, self.T_cap, self.n_br, self.device = batch, T, n_br, device
        self.G = (batch + 7) // 8
        self.seed = seed
        bf = torch.bfloat16
        # forward activations and the forward weight pack are fp16 (11 significant bits; same tensor-core rate as bf16;
        # log-magnitude inputs in [-21, 2] and O(1..100) activations are far inside its range), gradients stay bf16
        fa = torch.float16
        G = self.G
        self._store: Dict[str, List[torch.Tensor]] = {}
        # bf16 shadows of the activations that are the x operand of a weight gradient (written by the producing kernel next
        # to the fp16 tensor): both operands of one tcgen05.mma must share a type, and gradients are bf16
        self._shadow: Dict[int, List[torch.Tensor]] = {}
        for name, H, Cc, shadowed in self._ACTS:
            self._store[name] = [torch.zeros(G * H * T * 8 * Cc, dtype=fa, device=device) for _ in range(n_br)]
            setattr(self, name, [None] * n_br)
            if shadowed:
                self._store[name + "_s"] = [torch.zeros(G * H * T * 8 * Cc, dtype=bf, device=device) for _ in range(n_br)]
                self._shadow[id(getattr(self, name))] = [None] * n_br
        for name, H, Cc in self._ARGS:
            self._store[name] = [torch.zeros(G * H * T * 8 * Cc, dtype=torch.uint8, device=device) for _ in range(n_br)]
            setattr(self, name, [None] * n_br)
        self._store["emb"] = [torch.zeros(batch * T, device=device) for _ in range(n_br)]
        self.emb: List[torch.Tensor] = [None] * n_br
        self.set_T(T)
        # packed weights / gradients for cv2..cv8
        self.wf: Dict[str, List[torch.Tensor]] = {}
        self.wd: Dict[str, List[torch.Tensor]] = {}
        self.gp: Dict[str, List[torch.Tensor]] = {}
        for name, co, ci, kh, kw, _ in CONV_SPECS[1:]:
            self.wf[name] = [torch.zeros(kh * kw, co, ci, dtype=fa, device=device) for _ in range(n_br)]
            self.wd[name] = [torch.zeros(kh * kw, ci, co, dtype=bf, device=device) for _ in range(n_br)]
        self._grad_ws_ready = False
        self.step_ctr = torch.zeros(1, dtype=torch.int32, device=device)  # dropout seed word / Adam step
        self._x_in: List[Optional[torch.Tensor]] = [None] * n_br
        self._x_stride = 0
        self._train = False
        self._need_shadow = False
        self._p = 0.0
        # optional per-launch timing (bench.py): list of (tag, flops, start_event, end_event)
        self.timers: Optional[list] = None
        # backward: the weight gradient of layer L runs on a side stream next to the data gradient of the
        # same layer (both only read dY_L), so the partial last wave of one kernel is filled by CTAs of
        # the other (each kernel occupies an SM exclusively: ~200 KB of shared memory per CTA)
        self._side = torch.cuda.Stream(device=device)
        self.overlap_wgrad = True

    def set_T(self, T: int) -> None:
        """Work on clips of ``T <= T_cap`` frames: re-view the prefix of every workspace (no allocation, no copy)."""
        if T > self.T_cap or T < 1:
            raise ValueError(f"T = {T} outside this engine's capacity {self.T_cap}")
        self.T = T
        G = self.G
        for name, H, Cc, shadowed in self._ACTS:
            views = getattr(self, name)
            for br in range(self.n_br):
                views[br] = self._store[name][br][:G * H * T * 8 * Cc].view(G, H, T, 8, Cc)
                if shadowed:
                    self._shadow[id(views)][br] = self._store[name + "_s"][br][:G * H * T * 8 * Cc].view(G, H, T, 8, Cc)
        for name, H, Cc in self._ARGS:
            views = getattr(self, name)
            for br in range(self.n_br):
                views[br] = self._store[name][br][:G * H * T * 8 * Cc].view(G, H, T, 8, Cc)
        for br in range(self.n_br):
            self.emb[br] = self._store["emb"][br][:self.B * T].view(self.B, T)

    def _timed(self, tag: str, flops: float):
        eng = self

        class _T:
            def __enter__(self_inner):
                if eng.timers is not None:
                    self_inner.e0 = torch.cuda.Event(enable_timing=True)
                    self_inner.e1 = torch.cuda.Event(enable_timing=True)
                    self_inner.e0.record()
                return self_inner

            def __exit__(self_inner, *exc):
                if eng.timers is not None:
                    self_inner.e1.record()
                    eng.timers.append((tag, flops, self_inner.e0, self_inner.e1))
                return False

        return _T()

    def _conv_flops(self, name, H) -> float:
        _, co, ci, kh, kw, _ = next(s for s in CONV_SPECS if s[0] == name)
        return 2.0 * self.B * H * self.T * co * ci * kh * kw * self.n_br

    # -- workspaces needed only for backward -------------------------------------------------------
    def _ensure_grad_ws(self):
        if self._grad_ws_ready:
            return
        bf, dev, G, T = torch.bfloat16, self.device, self.G, self.T_cap
        n = G * 96 * T * 8 * 64
        self.ga = [torch.zeros(n, dtype=bf, device=dev) for _ in range(self.n_br)]
        self.gb = [torch.zeros(n, dtype=bf, device=dev) for _ in range(self.n_br)]
        total = sum(co * ci * kh * kw for _, co, ci, kh, kw, _ in CONV_SPECS[1:])
        self.gp_flat = [torch.zeros(total, device=dev) for _ in range(self.n_br)]
        for br in range(self.n_br):
            off = 0
            for name, co, ci, kh, kw, _ in CONV_SPECS[1:]:
                self.gp.setdefault(name, []).append(self.gp_flat[br][off:off + co * ci * kh * kw])
                off += co * ci * kh * kw
        self._grad_ws_ready = True

    # -- weights -----------------------------------------------------------------------------------
    def pack_weights_async(self, params: Sequence[Dict[str, torch.Tensor]], need_dgrad: bool):
        """pack_weights on the side stream; forward() joins it right before cv2, so the packs overlap cv1."""
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        self._side.wait_event(ready)
        with torch.cuda.stream(self._side):
            self.pack_weights(params, need_dgrad)
            self._pack_event = torch.cuda.Event()
            self._pack_event.record(self._side)

    def pack_weights(self, params: Sequence[Dict[str, torch.Tensor]], need_dgrad: bool):
        """fp32 state_dict-layout weights -> fp16 forward packs and bf16 flipped/transposed packs for the data gradient:
        ONE multi-tensor launch for every layer of every branch."""
        ws, wf, wd, geo = [], [], [], []
        for br in range(self.n_br):
            for name, co, ci, kh, kw, _ in CONV_SPECS[1:]:
                ws.append(params[br][f"pretrained.{name}.weight"])
                wf.append(self.wf[name][br])
                wd.append(self.wd[name][br] if need_dgrad else None)
                geo.append((co, ci, kh, kw))
        L.check(L.lib().zns_pack_weights_multi(len(ws), L.ptr_array(ws), L.int_array([g[0] for g in geo]),
                                               L.int_array([g[1] for g in geo]), L.int_array([g[2] for g in geo]),
                                               L.int_array([g[3] for g in geo]), L.ptr_array(wf), L.ptr_array(wd), 1,
                                               L.current_stream()))

    # -- forward -----------------------------------------------------------------------------------
    def _conv(self, name, H, ins, outs, params, relu, drop, layer_id):
        _, co, ci, kh, kw, _ = next(s for s in CONV_SPECS if s[0] == name)
        d = L.conv_desc(self.B, H, self.T, ci, co, kh, kw, relu=relu, dropout_p=self._p if drop else 0.0,
                        seed=self.seed, rng_stream=layer_id * 2, seed_dev=self.step_ctr if drop and self._p > 0 else None,
                        fmt=L.FMT_FORWARD_F16)
        bias = [params[br][f"pretrained.{name}.bias"] for br in range(self.n_br)]
        with self._timed(_fwd_tag(co, H), self._conv_flops(name, H)):
            sh = self._shadow.get(id(outs)) if self._need_shadow else None
            L.check(L.lib().zns_conv_fwd(C.byref(d), self.n_br, L.ptr_array(ins), L.ptr_array(self.wf[name]),
                                         L.ptr_array(bias), None, L.ptr_array(outs), L.ptr_array(sh) if sh else None,
                                         L.current_stream()))

    def _conv_pool(self, name, H, pool, ins, outs, args, params, layer_id):
        """conv + bias -> MaxPool((pool, 1)) -> ReLU -> Dropout in ONE launch (zns_conv_pool_fwd); outs are the pooled tensors."""
        _, co, ci, kh, kw, _ = next(s for s in CONV_SPECS if s[0] == name)
        d = L.conv_desc(self.B, H, self.T, ci, co, kh, kw, relu=0, dropout_p=self._p, seed=self.seed, rng_stream=layer_id * 2,
                        seed_dev=self.step_ctr if self._p > 0 else None, fmt=L.FMT_FORWARD_F16)
        bias = [params[br][f"pretrained.{name}.bias"] for br in range(self.n_br)]
        sh = self._shadow[id(outs)] if self._need_shadow else None
        with self._timed(_fwd_tag(co, H), self._conv_flops(name, H)):
            L.check(L.lib().zns_conv_pool_fwd(C.byref(d), pool, self.n_br, L.ptr_array(ins), L.ptr_array(self.wf[name]),
                                              L.ptr_array(bias), L.ptr_array(outs), L.ptr_array(sh) if sh else None,
                                              L.ptr_array(args) if self._need_shadow else None, L.current_stream()))

    def _pool(self, H, Cc, pool, ys, outs, layer_id):
        sh = self._shadow[id(outs)] if self._need_shadow else None
        L.check(L.lib().zns_pool_fwd_nbr(self.n_br, L.ptr_array(ys), L.ptr_array(outs), self.B, H, self.T, Cc, pool, self._p,
                                         self.seed, L.ptr(self.step_ctr) if self._p > 0 else None, layer_id * 2, 1,
                                         L.ptr_array(sh) if sh else None, L.current_stream()))

    def forward(self, xs: Sequence[torch.Tensor], x_clip_stride: int, params: Sequence[Dict[str, torch.Tensor]],
                train: bool, dropout_p: float = 0.1, x_row_stride: Optional[int] = None,
                need_grad: Optional[bool] = None) -> List[torch.Tensor]:
        """xs[br]: fp32 CUDA tensor; clip b / row h / frame w is read at
        ``data_ptr + b * x_clip_stride + h * x_row_stride + w`` (x_row_stride defaults to T).
        Returns [emb (B, T)] per branch."""
        lib, st = L.lib(), L.current_stream()
        self._train, self._p = train, (dropout_p if train else 0.0)
        self._need_shadow = train if need_grad is None else bool(need_grad)   # bf16 copies only when a backward follows
        self._x_in, self._x_stride = list(xs), x_clip_stride
        self._x_row = self.T if x_row_stride is None else int(x_row_stride)
        sh1 = self._shadow[id(self.x1)] if self._need_shadow else None
        L.check(lib.zns_conv1_fwd_nbr(self.n_br, L.ptr_array(xs), x_clip_stride, self._x_row,
                                      L.ptr_array([p["pretrained.cv1.weight"] for p in params[:self.n_br]]),
                                      L.ptr_array([p["pretrained.cv1.bias"] for p in params[:self.n_br]]), L.ptr_array(self.x1),
                                      self.B, N_BINS, self.T, self._p, self.seed,
                                      L.ptr(self.step_ctr) if self._p > 0 else None, 100, 1,
                                      L.ptr_array(sh1) if sh1 else None, st))
        if getattr(self, "_pack_event", None) is not None:      # packs issued by pack_weights_async
            torch.cuda.current_stream().wait_event(self._pack_event)
            self._pack_event = None
        self._conv_pool("cv2", 96, 3, self.x1, self.p2, self.arg2, params, layer_id=2)
        self._conv("cv3", 32, self.p2, self.x3, params, relu=1, drop=True, layer_id=3)
        self._conv_pool("cv4", 32, 4, self.x3, self.p4, self.arg4, params, layer_id=4)
        self._conv("cv5", 8, self.p4, self.x5, params, relu=1, drop=True, layer_id=5)
        self._conv("cv6", 8, self.x5, self.y6, params, relu=0, drop=False, layer_id=6)
        self._pool(8, 256, 8, self.y6, self.p6, 6)
        self._conv("cv7", 1, self.p6, self.x7, params, relu=1, drop=True, layer_id=7)
        self._conv("cv8", 1, self.x7, self.x8, params, relu=1, drop=True, layer_id=8)
        L.check(lib.zns_head_fwd_nbr(self.n_br, L.ptr_array(self.x8), L.ptr_array([p["fc1.weight"] for p in params[:self.n_br]]),
                                     L.ptr_array([p["fc1.bias"] for p in params[:self.n_br]]), L.ptr_array(self.emb),
                                     self.B, self.T, 1, st))
        return self.emb

    # -- backward ----------------------------------------------------------------------------------
    def _wgrad(self, name, H, xs, dys, grads):
        """Weight + bias gradient of one layer; with overlap on, issued on the side stream.  Returns an
        event that fires when dys may be overwritten."""
        if not self.overlap_wgrad or self.timers is not None:
            self._wgrad_now(name, H, xs, dys, grads)
            return None
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        self._side.wait_event(ready)
        with torch.cuda.stream(self._side):
            self._wgrad_now(name, H, xs, dys, grads)
            done = torch.cuda.Event()
            done.record(self._side)
        return done

    def _wgrad_now(self, name, H, xs, dys, grads):
        _, co, ci, kh, kw, _ = next(s for s in CONV_SPECS if s[0] == name)
        lib, st = L.lib(), L.current_stream()
        d = L.conv_desc(self.B, H, self.T, ci, co, kh, kw)
        xb = self._shadow[id(xs)]                      # bf16 copy of the fp16 forward activation (same type as dy)
        tag = "conv_wgrad_umma<128> stacked dy (c_out=64)" if co == 64 else "conv_wgrad_umma<128>"
        with self._timed(tag, self._conv_flops(name, H)):
            L.check(lib.zns_conv_wgrad(C.byref(d), self.n_br, L.ptr_array(xb), L.ptr_array(dys), L.ptr_array(self.gp[name]), st))
        L.check(lib.zns_bias_grad_nbr(self.n_br, L.ptr_array(dys), self.B, H, self.T, co,
                                      L.ptr_array([grads[br][f"pretrained.{name}.bias"] for br in range(self.n_br)]), st))

    def _unpack_all(self, grads, layers=None):
        """Packed [tap][c_out][c_in] weight gradients of every (or the named) layer and branch -> state_dict layout (+=), one
        launch; the packed accumulators are cleared behind the read (next step's atomics start from zero: no fill kernel)."""
        gp, g, geo = [], [], []
        for br in range(self.n_br):
            for name, co, ci, kh, kw, _ in CONV_SPECS[1:]:
                if layers is not None and name not in layers:
                    continue
                gp.append(self.gp[name][br])
                g.append(grads[br][f"pretrained.{name}.weight"])
                geo.append((co, ci, kh, kw))
        L.check(L.lib().zns_unpack_grads_multi(len(gp), L.ptr_array(gp), L.int_array([x[0] for x in geo]),
                                               L.int_array([x[1] for x in geo]), L.int_array([x[2] for x in geo]),
                                               L.int_array([x[3] for x in geo]), 1.0, 1, 1, L.ptr_array(g), L.current_stream()))

    def _dgrad(self, name, H, dys, masks, outs):
        _, co, ci, kh, kw, _ = next(s for s in CONV_SPECS if s[0] == name)
        scale = 1.0 / (1.0 - self._p) if self._p > 0 else 1.0
        d = L.conv_desc(self.B, H, self.T, co, ci, kh, kw, relu=0, out_scale=scale)
        with self._timed(_fwd_tag(ci, H), self._conv_flops(name, H)):
            L.check(L.lib().zns_conv_fwd(C.byref(d), self.n_br, L.ptr_array(dys), L.ptr_array(self.wd[name]), None,
                                         L.ptr_array(masks), L.ptr_array(outs), None, L.current_stream()))

    def _unpool(self, H, Cc, pool, ys, dps, outs):
        L.check(L.lib().zns_pool_bwd_nbr(self.n_br, L.ptr_array(ys), L.ptr_array(dps), L.ptr_array(outs), self.B, H, self.T, Cc,
                                         pool, 1, L.current_stream()))

    def _unpool_arg(self, H, Cc, pool, args, dps, outs):
        L.check(L.lib().zns_pool_bwd_arg_nbr(self.n_br, L.ptr_array(args), L.ptr_array(dps), L.ptr_array(outs), self.B, H, self.T,
                                             Cc, pool, L.current_stream()))

    LATE_LAYERS = ("cv5", "cv6", "cv7", "cv8")     # their gradients are complete after phase "late" of backward()
    EARLY_LAYERS = ("cv2", "cv3", "cv4")

    def backward(self, d_embs: Sequence[torch.Tensor], params: Sequence[Dict[str, torch.Tensor]],
                 grads: Sequence[Dict[str, torch.Tensor]], phase: Optional[str] = None) -> None:
        """Accumulate (+=) parameter gradients into ``grads[br][name]`` (fp32, state_dict layout;
        they must be zeroed by the caller).  ``d_embs[br]``: (B, T) fp32 gradient of the loss.

        ``phase`` splits the pass for data-parallel training: "late" runs the head and cv8..cv5 (80 % of the parameters:
        their gradients, fc1 included, are final when it returns, so their exchange can overlap the rest), "early" runs
        cv4..cv1 on the dp4 that "late" left in the workspace.  None runs both."""
        self._ensure_grad_ws()
        if not self._need_shadow:
            raise RuntimeError("backward() needs a forward(..., train=True) or forward(..., need_grad=True) before it")
        lib, st = L.lib(), L.current_stream()
        scale = 1.0 / (1.0 - self._p) if self._p > 0 else 1.0
        ga, gb = self.ga, self.gb
        nb = self.n_br
        main = torch.cuda.current_stream()

        def join(ev):
            if ev is not None:
                main.wait_event(ev)

        if phase in (None, "late"):
            self._backward_late(d_embs, params, grads, join)
            if phase == "late":
                self._unpack_all(grads, self.LATE_LAYERS)
                return
        self._backward_early(params, grads, join)
        self._unpack_all(grads, None if phase is None else self.EARLY_LAYERS)

    def _backward_late(self, d_embs, params, grads, join):
        lib, st = L.lib(), L.current_stream()
        scale = 1.0 / (1.0 - self._p) if self._p > 0 else 1.0
        ga, gb = self.ga, self.gb
        nb = self.n_br
        L.check(lib.zns_head_bwd_nbr(nb, L.ptr_array(self.x8), L.ptr_array(self.emb), L.ptr_array(list(d_embs)[:nb]),
                                     L.ptr_array([p["fc1.weight"] for p in params[:nb]]),
                                     L.ptr_array([g["fc1.weight"] for g in grads[:nb]]),
                                     L.ptr_array([g["fc1.bias"] for g in grads[:nb]]), L.ptr_array(ga), self.B, self.T, scale, 1, st))
        # dY_L alternates between ga and gb; the wgrad of layer L (side stream) must have finished
        # reading its buffer before the dgrad of layer L-1 writes into it.
        w8 = self._wgrad("cv8", 1, self.x7, ga, grads)
        self._dgrad("cv8", 1, ga, self.x7, gb)          # dy7
        w7 = self._wgrad("cv7", 1, self.p6, gb, grads)
        join(w8)
        self._dgrad("cv7", 1, gb, self.p6, ga)          # dp6 (masked)
        join(w7)
        self._unpool(8, 256, 8, self.y6, ga, gb)        # dy6
        w6 = self._wgrad("cv6", 8, self.x5, gb, grads)
        self._dgrad("cv6", 8, gb, self.x5, ga)          # dy5
        w5 = self._wgrad("cv5", 8, self.p4, ga, grads)
        join(w6)
        self._dgrad("cv5", 8, ga, self.p4, gb)          # dp4
        join(w5)

    def _backward_early(self, params, grads, join):
        lib, st = L.lib(), L.current_stream()
        ga, gb = self.ga, self.gb
        nb = self.n_br
        self._unpool_arg(32, 128, 4, self.arg4, gb, ga)  # dy4
        w4 = self._wgrad("cv4", 32, self.x3, ga, grads)
        self._dgrad("cv4", 32, ga, self.x3, gb)         # dy3
        w3 = self._wgrad("cv3", 32, self.p2, gb, grads)
        join(w4)
        self._dgrad("cv3", 32, gb, self.p2, ga)         # dp2
        join(w3)
        self._unpool_arg(96, 64, 3, self.arg2, ga, gb)   # dy2
        w2 = self._wgrad("cv2", 96, self.x1, gb, grads)
        self._dgrad("cv2", 96, gb, self.x1, ga)         # dy1
        join(w2)
        L.check(lib.zns_conv1_wgrad_nbr(nb, L.ptr_array(ga), L.ptr_array(self._x_in), self._x_stride, self._x_row,
                                        L.ptr_array([g["pretrained.cv1.weight"] for g in grads[:nb]]),
                                        L.ptr_array([g["pretrained.cv1.bias"] for g in grads[:nb]]), self.B, N_BINS, self.T, st))
