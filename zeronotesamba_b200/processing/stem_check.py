"""Drop-in for the RMS gate of /root/reference/zeroNoteSamba/processing/stem_check.py
(``check_CL_clips``, lines 21-51), which pretext.py:66-81 uses to accept or reject a candidate clip:
frame RMS (2048 / hop 512, librosa.feature.rms defaults) of both stems, fraction of frames with
``ros/2 < stem < 4*ros``, accept iff ``lower_p < fraction <= upper_p``.  Runs on the GPU next to the
VQT (zns_rms_gate); ``rms_fraction_batch`` is the batched tensor variant."""
from __future__ import annotations

import numpy as np
import numpy.typing as npt

from .. import _lib as L


def rms_fraction_batch(stem, ros):
    """stem, ros: CUDA float32 [B, N] -> CUDA float32 [B], fraction of accepted RMS frames per clip."""
    import torch
    assert stem.is_cuda and ros.is_cuda and stem.shape == ros.shape and stem.dim() == 2
    stem, ros = stem.contiguous().float(), ros.contiguous().float()
    b, n = stem.shape
    counts = torch.empty(b, dtype=torch.int32, device=stem.device)
    L.check(L.lib().zns_rms_gate(L.ptr(stem), L.ptr(ros), b, n, L.ptr(counts), L.current_stream()))
    return counts.float() / float(1 + n // 512)


def check_CL_clips(anchor: npt.NDArray[np.float32], positive: npt.NDArray[np.float32], lower_p: float, upper_p: float) -> bool:
    """
    Function for thresholding anchor vs positive. Goal is to make sure drum clip has enough energy.
    -- anchor: selected stem combination
    -- positive: other stem combination
    -- lower_p: lower RMS percentage threshold
    -- upper_p: upper RMS percentage threshold
    """
    import torch
    a = torch.from_numpy(np.ascontiguousarray(anchor, dtype=np.float32).reshape(1, -1)).cuda()
    p = torch.from_numpy(np.ascontiguousarray(positive, dtype=np.float32).reshape(1, -1)).cuda()
    rms_perc = float(rms_fraction_batch(a, p)[0])
    return bool(lower_p < rms_perc <= upper_p)
