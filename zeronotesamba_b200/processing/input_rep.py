"""Drop-in for /root/reference/zeroNoteSamba/processing/input_rep.py::generate_XQT (lines 11-57).

``generate_XQT(signal, sample_rate, mode)`` keeps the reference's contract -- 1-D float32 numpy
array in, float32 ``(96, 1 + len(signal)//256)`` numpy array of ``log(|XQT| + 1e-9)`` out, the same
exception text for an unknown mode -- but the transform runs in libzns_sm100 on the current CUDA
device (batched octave-wise decimation + framed filterbank + fused log-magnitude).  There is no
CPU fallback.  ``xqt_batch`` is the tensor-in / tensor-out batched variant used inside the
training loop.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import numpy as np
import numpy.typing as npt

from .. import _lib as L

HOP_LENGTH = 256          # input_rep.py:18
OCTAVE_RESO = 12          # input_rep.py:20
NUM_OCTAVES = 8           # input_rep.py:21
N_BINS = NUM_OCTAVES * OCTAVE_RESO
FMIN_C0 = 440.0 * 2.0 ** ((12 - 69) / 12.0)   # librosa.note_to_hz("C0"), input_rep.py:24


class VQTPlan:
    """Owns a zns_vqt_plan (filter kernels, decimator taps, scratch) for one (sr, mode, capacity)."""

    def __init__(self, sample_rate: int, mode: str, max_batch: int, max_samples: int):
        if mode not in ("vqt", "cqt"):
            raise Exception("Mode can only be vqt or cqt!")  # input_rep.py:56-57
        self.sample_rate, self.mode = int(sample_rate), mode
        self.max_batch, self.max_samples = int(max_batch), int(max_samples)
        h = C.c_void_p()
        gamma = -1.0 if mode == "vqt" else 0.0
        L.check(L.lib().zns_vqt_plan_create(self.sample_rate, HOP_LENGTH, N_BINS, OCTAVE_RESO, FMIN_C0, gamma,
                                            self.max_batch, self.max_samples, C.byref(h)))
        self._h = h

    def frames(self, n_samples: int) -> int:
        return L.lib().zns_vqt_num_frames(int(n_samples), HOP_LENGTH)

    def forward(self, y, out=None):
        """y: CUDA float32 tensor [B, N] (contiguous) -> CUDA float32 [B, 96, 1 + N//256]."""
        import torch
        assert y.is_cuda and y.dtype == torch.float32 and y.dim() == 2 and y.is_contiguous()
        b, n = y.shape
        if out is None:
            out = torch.empty(b, N_BINS, self.frames(n), device=y.device, dtype=torch.float32)
        L.check(L.lib().zns_vqt_forward(self._h, L.ptr(y), b, n, L.ptr(out), L.current_stream()))
        return out

    def forward_host(self, y: np.ndarray) -> np.ndarray:
        """y: float32 numpy [B, N] -> float32 numpy [B, 96, F]; H2D + transform + D2H + sync."""
        y = np.ascontiguousarray(y, dtype=np.float32)
        b, n = y.shape
        out = np.empty((b, N_BINS, self.frames(n)), dtype=np.float32)
        L.check(L.lib().zns_vqt_forward_host(self._h, y.ctypes.data, b, n, out.ctypes.data, L.current_stream()))
        return out

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                L.lib().zns_vqt_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass


_PLANS: Dict[Tuple[int, str, int, int], VQTPlan] = {}


def get_plan(sample_rate: int, mode: str, batch: int, n_samples: int) -> VQTPlan:
    if mode not in ("vqt", "cqt"):
        raise Exception("Mode can only be vqt or cqt!")
    # capacity rounded up so that clips of similar length share a plan
    cap = 1 << max(12, int(n_samples - 1).bit_length())
    key = (int(sample_rate), mode, int(batch), cap)
    plan = _PLANS.get(key)
    if plan is None:
        plan = _PLANS[key] = VQTPlan(sample_rate, mode, batch, cap)
    return plan


def generate_XQT(signal: npt.NDArray[np.float32], sample_rate: int, mode: str) -> npt.NDArray[np.float32]:
    """
    Generates a high-resolution XQT spectrogram (same call as the reference, input_rep.py:11).
    -- signal: signal to compute XQT on
    -- sample_rate: self-explanatory
    -- mode: can be either vqt or cqt
    """
    if mode not in ("vqt", "cqt"):
        raise Exception("Mode can only be vqt or cqt!")
    sig = np.ascontiguousarray(signal, dtype=np.float32)
    if sig.ndim != 1:
        raise ValueError("generate_XQT expects a 1-D signal")
    plan = get_plan(sample_rate, mode, 1, sig.shape[0])
    return plan.forward_host(sig[None, :])[0]


def xqt_batch(y, sample_rate: int = 16000, mode: str = "vqt"):
    """Batched on-device variant: CUDA float32 [B, N] -> CUDA float32 [B, 96, 1 + N//256]."""
    plan = get_plan(sample_rate, mode, y.shape[0], y.shape[1])
    return plan.forward(y)
