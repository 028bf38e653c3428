"""CPU replay of the tcgen05 VQT level plans (csrc/vqt_umma.cu) against the oracle.

The level kernels execute a host-built list of MMAs over views of a chunk-major shared-memory image.  The
list, the coefficient image and the epilogue arithmetic are replayed here in numpy (fp16 operands, float64
accumulation) for one 128-row tile per level, and compared with the oracle's decimator
(oracle.vqt_oracle.resample_2to1_f64) and time-domain filter kernels (octave_time_kernels): this pins the
descriptor arithmetic, the coefficient tiles and the column maps without a GPU.  (The descriptor semantics
themselves -- rows at 16-byte pitch, overlapping views -- were measured on a B200: profiles/r02_umma_view_probe.txt.)
"""
import ctypes as C

import numpy as np
import pytest

from oracle import vqt_oracle as vo
from zeronotesamba_b200 import _lib as L

MAX_MMA = 112


class VqtMma(C.Structure):
    _fields_ = [("a_off", C.c_uint32), ("b_off", C.c_uint32), ("n", C.c_uint16), ("d_col", C.c_uint16),
                ("term", C.c_uint16), ("b_rows", C.c_uint16)]


class VqtLevel(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("q", "hb", "ha", "rtot", "a_lbo", "fpr", "hop", "n_fft", "bin0", "dec_w", "wacc",
                                       "dec_a_col", "dec_b_col", "fb_a_col", "fb_b_col", "fb_b_stride", "tmem_cols",
                                       "n_mma", "b_bytes")] + \
               [("dec_scale", C.c_float), ("fb_scale", C.c_float), ("mma", VqtMma * MAX_MMA)]


def level_plan(level, mode="vqt"):
    lv = VqtLevel()
    img = np.zeros(200 * 1024, dtype=np.uint16)
    gamma = -1.0 if mode == "vqt" else 0.0
    L.check(L.lib().zns_dbg_vqt_level_plan(16000, 256, 96, 12, vo.FMIN_C0, gamma, level, C.byref(lv), C.sizeof(lv),
                                           img.ctypes.data, img.size))
    return lv, img[: lv.b_bytes // 2].view(np.float16)


def split(x):
    x = x.astype(np.float32)
    h1 = x.astype(np.float16)
    h2 = ((x - h1.astype(np.float32)) * np.float32(2048.0)).astype(np.float16)
    return h1, h2


def replay(lv, bimg, sig, row0):
    """Emulate one tile: returns the TMEM accumulator [128][512] (float64)."""
    q, R = lv.q, 8 * lv.q
    n_rows = 128 + lv.hb + lv.ha
    planes = [np.zeros(q * lv.rtot * 8 if q > 1 else lv.rtot * 8, dtype=np.float16) for _ in range(2)]
    first = (row0 - lv.hb) * R
    idx = first + np.arange(n_rows * R)
    x = np.where((idx >= 0) & (idx < sig.size), sig[np.clip(idx, 0, sig.size - 1)], 0.0).astype(np.float32)
    h1, h2 = split(x)
    for i in range(n_rows * q):
        r, c = divmod(i, q)
        off = (16 * r if q == 1 else c * lv.a_lbo + 16 * r) // 2
        planes[0][off:off + 8] = h1[8 * i:8 * i + 8]
        planes[1][off:off + 8] = h2[8 * i:8 * i + 8]
    D = np.zeros((128, 512))
    r = np.arange(128)[:, None]
    k = np.arange(16)[None, :]
    for i in range(lv.n_mma):
        m = lv.mma[i]
        a_idx = (m.a_off + (k // 8) * lv.a_lbo + 16 * r) // 2 + (k % 8)
        A = planes[m.term][a_idx].astype(np.float64)
        n = np.arange(m.n)[:, None]
        b_idx = (m.b_off + (k // 8) * 16 * m.b_rows + 16 * n) // 2 + (k % 8)
        B = bimg[b_idx].astype(np.float64)
        D[:, m.d_col:m.d_col + m.n] += A @ B.T
    return D


@pytest.mark.parametrize("level", range(8))
def test_level_plan_replay_matches_oracle(level):
    lv, bimg = level_plan(level)
    R = 8 * lv.q
    rng = np.random.default_rng(level)
    n_sig = R * 128 * 3 + 37
    sig = (0.5 * rng.standard_normal(n_sig)).astype(np.float32)
    row0 = 128                    # an interior tile
    D = replay(lv, bimg, sig, row0)
    # ---- decimator ----
    if lv.dec_w:
        want = vo.resample_2to1_f64(sig)          # sqrt(2) sum h x, zero extended
        got = (D[:, lv.dec_a_col:lv.dec_a_col + lv.dec_w] + D[:, lv.dec_b_col:lv.dec_b_col + lv.dec_w] / 2048.0) * lv.dec_scale
        t = (row0 + np.arange(128))[:, None] * lv.dec_w + np.arange(lv.dec_w)[None, :]
        ref = want[t]
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    else:
        assert level == 7
    # ---- filterbank ----
    g, n_fft = vo.octave_time_kernels(level, 16000.0, vo.default_gamma())
    assert n_fft == lv.n_fft
    for j in range(lv.fpr):
        a = D[:, lv.fb_a_col + 48 * j: lv.fb_a_col + 48 * j + 48]
        b = D[:, lv.fb_b_col + 24 * j: lv.fb_b_col + 24 * j + 24]
        c = (a[:, :24] + (a[:, 24:48] + b) / 2048.0) * lv.fb_scale
        got = c[:, 0::2] + 1j * c[:, 1::2]                        # [128 rows][12 bins]
        f = (row0 + np.arange(128)) * lv.fpr + j
        start = f * lv.hop - n_fft // 2
        frames = sig[start[:, None] + np.arange(n_fft)[None, :]].astype(np.float64)
        ref = frames @ g.T
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max(), (level, j)


def test_level_plan_edges_zero_extension():
    """Rows before the clip start / after its end read zeros (decimator semantics of resampy)."""
    lv, bimg = level_plan(2)
    rng = np.random.default_rng(9)
    sig = (0.5 * rng.standard_normal(64 * 128 + 11)).astype(np.float32)
    D = replay(lv, bimg, sig, 0)
    want = vo.resample_2to1_f64(sig)
    got = (D[:, :lv.dec_w] + D[:, lv.dec_b_col:lv.dec_b_col + lv.dec_w] / 2048.0) * lv.dec_scale
    t = np.arange(128)[:, None] * lv.dec_w + np.arange(lv.dec_w)[None, :]
    n_valid = sig.size // 2
    ok = t < n_valid
    assert np.abs(got[ok] - want[t[ok]]).max() <= 2e-6 * np.abs(want).max()


def test_level_geometry_table():
    rows = {0: (32, 1), 1: (16, 1), 2: (8, 1), 3: (4, 1), 4: (4, 2), 5: (4, 4), 6: (1, 2), 7: (1, 4)}
    for level, (q, fpr) in rows.items():
        lv, _ = level_plan(level)
        assert (lv.q, lv.fpr) == (q, fpr)
        assert lv.n_mma <= MAX_MMA and lv.tmem_cols in (128, 256, 512)
        assert 2 * lv.q * lv.rtot * 16 + lv.b_bytes < 220 * 1024


def test_cqt_geometry_falls_back():
    """gamma = 0 (CQT) needs n_fft = 256 at every octave: not covered by the level kernels (the plan then keeps the
    round-1 kernels); the host hook reports it instead of building a wrong plan."""
    with pytest.raises(L.ZnsError, match="not supported"):
        level_plan(0, "cqt")
