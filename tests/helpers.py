"""Test-side helpers (torch reference conversions; never imported by the product package)."""
import numpy as np
import torch


def to_act(x: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    """fp32 NCHW [B,C,H,W] -> act bf16 (or fp16) [G,H,W,8,C] (zero-padded clips)."""
    B, C, H, W = x.shape
    G = (B + 7) // 8
    xp = torch.zeros(G * 8, C, H, W, dtype=x.dtype, device=x.device)
    xp[:B] = x
    return xp.view(G, 8, C, H, W).permute(0, 3, 4, 1, 2).contiguous().to(dtype)


def from_act(a: torch.Tensor, B: int) -> torch.Tensor:
    """act bf16 [G,H,W,8,C] -> fp32 NCHW [B,C,H,W]."""
    G, H, W, _, C = a.shape
    return a.float().permute(0, 3, 4, 1, 2).reshape(G * 8, C, H, W)[:B].contiguous()


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).float()


def round16(x: torch.Tensor, dtype) -> torch.Tensor:
    return x.to(dtype).float()


def pack_wf(w: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
    """[co,ci,kh,kw] -> bf16 (or fp16) [kh*kw, co, ci]."""
    co, ci, kh, kw = w.shape
    return w.permute(2, 3, 0, 1).reshape(kh * kw, co, ci).contiguous().to(dtype)


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


def adam_delta_slack(g_ref, lr, eps=1e-8, rel_g=0.5):
    """Change of a first-step Adam update -lr g / (|g| + eps) when the gradient entry moves by rel_g of itself."""
    r = eps / np.abs(np.asarray(g_ref, dtype=np.float64))
    return lr * r * rel_g / (1.0 + r) ** 2


def check_adam_deltas(d_got, d_ref, g_ref, w0, lr, g_got=None, g_floor=1e-6, min_sign=0.99, eps=1e-8):
    """Discriminating check of one Adam step (first step: delta = -lr * g / (|g| + eps), eps = 1e-8) on sampled entries.

    (A) arithmetic, every entry (needs ``g_got``, this implementation's own gradient): the applied update equals Adam's
        first step on that gradient to 1e-3 relative of lr plus two ulps of the fp32 weight (w + delta is rounded to the
        fp32 grid) -- "one-step weight updates within 1e-3"; fails with no update, a wrong lr / eps / bias correction.
    (B) against the reference's update, on entries whose reference gradient stands clear of eps AND of this tensor's
        reduced-precision gradient noise (|g_ref| > max(g_floor, 10 * rms(g_got - g_ref))): the reference moved the weight
        by ~lr, the sign agrees on >= ``min_sign`` of them, and where it agrees the size matches to 1e-3 relative + two ulps
        + the update's own sensitivity to the gradient entry, lr (eps / |g|) |dg / g| / (1 + eps / |g|)^2 with |dg / g| <= 0.5.
    Returns (entries checked in (B), sign agreement)."""
    d_got, d_ref, g_ref, w0 = (np.asarray(a, dtype=np.float64) for a in (d_got, d_ref, g_ref, w0))
    ulp_all = np.spacing(np.abs(w0).astype(np.float32)).astype(np.float64)
    floor = g_floor
    if g_got is not None:
        g_got = np.asarray(g_got, dtype=np.float64)
        want = -lr * g_got / (np.abs(g_got) + eps)
        bad = np.abs(d_got - want) > 1e-3 * lr + 2 * ulp_all
        assert not bad.any(), f"{int(bad.sum())} of {d_got.size} updates are not Adam's step on the gradient: worst {np.abs(d_got - want)[bad].max():.3e}"
        floor = max(g_floor, 10.0 * float(np.sqrt(np.mean((g_got - g_ref) ** 2))))
    sel = np.abs(g_ref) > floor
    n = int(sel.sum())
    if n == 0:
        return 0, 1.0
    assert np.all(np.abs(d_ref[sel]) > 0.5 * lr), "golden deltas are not ~lr where |g| >> eps"
    same = np.sign(d_got[sel]) == np.sign(d_ref[sel])
    tol = 1e-3 * np.abs(d_ref[sel]) + 2 * ulp_all[sel] + adam_delta_slack(g_ref[sel], lr)
    bad = (np.abs(d_got[sel] - d_ref[sel]) > tol) & same
    assert not bad.any(), f"{int(bad.sum())} of {n} updates differ in size: worst {np.abs(d_got[sel] - d_ref[sel])[bad].max():.3e}"
    frac = float(same.mean())
    assert frac >= min_sign, f"update sign agreement {frac:.4f} < {min_sign} over {n} entries"
    return n, frac
