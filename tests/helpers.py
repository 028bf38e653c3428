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


def check_adam_deltas(d_got, d_ref, g_ref, w0, lr, g_floor=1e-6, min_sign=0.99):
    """Discriminating check of one Adam step (first step: delta = -lr * g / (|g| + eps), eps = 1e-8).

    On entries whose reference gradient is well above eps and above the bf16 noise of the gradient (|g_ref| > g_floor)
    the update must (a) be there -- the reference moved the weight by ~lr --, (b) have the reference's sign on at least
    ``min_sign`` of the entries, and (c) where the sign agrees, match the reference's size to 1e-3 relative plus two ulps
    of the fp32 weight (w + delta is rounded to the fp32 grid by both implementations).
    Returns (entries checked, sign agreement)."""
    d_got, d_ref, g_ref, w0 = (np.asarray(a, dtype=np.float64) for a in (d_got, d_ref, g_ref, w0))
    sel = np.abs(g_ref) > g_floor
    n = int(sel.sum())
    if n == 0:
        return 0, 1.0
    assert np.all(np.abs(d_ref[sel]) > 0.5 * lr), "golden deltas are not ~lr where |g| >> eps"
    same = np.sign(d_got[sel]) == np.sign(d_ref[sel])
    ulp = np.spacing(np.abs(w0[sel]).astype(np.float32)).astype(np.float64)
    tol = 1e-3 * np.abs(d_ref[sel]) + 2 * ulp
    bad = (np.abs(d_got[sel] - d_ref[sel]) > tol) & same
    assert not bad.any(), f"{int(bad.sum())} of {n} updates differ in size: worst {np.abs(d_got[sel] - d_ref[sel])[bad].max():.3e}"
    frac = float(same.mean())
    assert frac >= min_sign, f"update sign agreement {frac:.4f} < {min_sign} over {n} entries"
    return n, frac
