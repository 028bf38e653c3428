"""Test-side helpers (torch reference conversions; never imported by the product package)."""
import numpy as np
import torch


def to_act(x: torch.Tensor) -> torch.Tensor:
    """fp32 NCHW [B,C,H,W] -> act bf16 [G,H,W,8,C] (zero-padded clips)."""
    B, C, H, W = x.shape
    G = (B + 7) // 8
    xp = torch.zeros(G * 8, C, H, W, dtype=x.dtype, device=x.device)
    xp[:B] = x
    return xp.view(G, 8, C, H, W).permute(0, 3, 4, 1, 2).contiguous().to(torch.bfloat16)


def from_act(a: torch.Tensor, B: int) -> torch.Tensor:
    """act bf16 [G,H,W,8,C] -> fp32 NCHW [B,C,H,W]."""
    G, H, W, _, C = a.shape
    return a.float().permute(0, 3, 4, 1, 2).reshape(G * 8, C, H, W)[:B].contiguous()


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).float()


def pack_wf(w: torch.Tensor) -> torch.Tensor:
    """[co,ci,kh,kw] -> bf16 [kh*kw, co, ci]."""
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci).contiguous().to(torch.bfloat16)


def pack_wd(w: torch.Tensor) -> torch.Tensor:
    """[co,ci,kh,kw] -> bf16 [kh*kw (flipped), ci, co]."""
    co, ci, kh, kw = w.shape
    return w.flip(2, 3).permute(2, 3, 1, 0).reshape(kh * kw, ci, co).contiguous().to(torch.bfloat16)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def vqt_check(out: np.ndarray, ref: np.ndarray):
    """Tolerance of the VQT parity tests: on bins with |V_ref| >= 1e-2 max|V_ref| the
    magnitudes agree to 1e-4 relative; everywhere the absolute magnitude error is <= 1e-6 max|V_ref|.
    Returns (max_rel_above_floor, max_abs_over_max)."""
    v = np.exp(out.astype(np.float64)) - 1e-9
    r = np.exp(ref.astype(np.float64)) - 1e-9
    mx = r.max()
    big = r >= 1e-2 * mx
    rel = float((np.abs(v - r)[big] / r[big]).max())
    ab = float(np.abs(v - r).max() / mx)
    return rel, ab
