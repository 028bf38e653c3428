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

MAX_MMA = 120


class VqtMma(C.Structure):
    _fields_ = [("a_off", C.c_uint32), ("b_off", C.c_uint32), ("n", C.c_uint16), ("d_col", C.c_uint16),
                ("term", C.c_uint8), ("job", C.c_uint8), ("part", C.c_uint8), ("b_rows8", C.c_uint8)]


class VqtSeg(C.Structure):
    _fields_ = [("begin", C.c_uint16), ("count", C.c_uint16), ("job", C.c_uint8), ("part", C.c_uint8), ("flags", C.c_uint8),
                ("pad", C.c_uint8)]


N_ISSUERS = 4


class VqtLevel(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("q", "hb", "ha", "rtot", "a_lbo", "fpr", "hop", "n_fft", "bin0", "dec_w", "dec_wp",
                                       "n_pass", "fb_n1", "fb_n2", "pg", "gpt", "n_slots", "slot_term_bytes")] + \
               [("g_order", C.c_int * 8), ("ring_base", C.c_int * 2), ("ring_width", C.c_int * 2), ("ring_stages", C.c_int * 2),
                ("n_jobs", C.c_int), ("ep_job", C.c_int * 3), ("n_mma", C.c_int), ("b_bytes", C.c_int),
                ("dec_scale", C.c_float), ("fb_scale", C.c_float), ("n_seg", C.c_int),
                ("seg_begin", (C.c_int * 9) * N_ISSUERS), ("seg", VqtSeg * 64), ("mma", VqtMma * MAX_MMA),
                ("pk", C.c_uint32 * (4 * MAX_MMA))]


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


def acc_base(lv, job):
    """TMEM column of a job's accumulator stage for the first tile of a CTA."""
    if job == 0:
        return lv.ring_base[0]
    return lv.ring_base[1] + ((job - 1) % lv.ring_stages[1]) * lv.ring_width[1]


def replay(lv, bimg, sig, row0):
    """Emulate one tile the way the kernel runs it (group by group): returns the TMEM accumulator [128][512]."""
    q, R = lv.q, 8 * lv.q
    n_rows = 128 + lv.hb + lv.ha
    first = (row0 - lv.hb) * R
    idx = first + np.arange(n_rows * R)
    x = np.where((idx >= 0) & (idx < sig.size), sig[np.clip(idx, 0, sig.size - 1)], 0.0).astype(np.float32)
    h1, h2 = split(x)
    # slots: [group][term] -> halfword array of slot_term_bytes
    slots = [[np.zeros(lv.slot_term_bytes // 2, dtype=np.float16) for _ in range(2)] for _ in range(lv.gpt)]
    for i in range(n_rows * q):
        r, c = divmod(i, q)
        g, cl = divmod(c, lv.pg)
        off = (16 * r if q == 1 else cl * lv.a_lbo + 16 * r) // 2
        slots[g][0][off:off + 8] = h1[8 * i:8 * i + 8]
        slots[g][1][off:off + 8] = h2[8 * i:8 * i + 8]
    D = np.full((128, 512), np.nan)            # garbage until a clearing MMA has run
    r = np.arange(128)[:, None]
    k = np.arange(16)[None, :]
    started, finished, covered = set(), set(), set()
    # the four issuers run concurrently; their accumulator units are disjoint, so any interleaving gives the same result
    for isr in range(N_ISSUERS):
        for pos in range(lv.gpt):
            g = lv.g_order[pos]
            for si in range(lv.seg_begin[isr][pos], lv.seg_begin[isr][pos + 1]):
                sg = lv.seg[si]
                unit = (sg.job, sg.part)
                base = acc_base(lv, sg.job)
                w0, w1 = (lv.dec_wp, lv.dec_wp) if sg.job else (lv.fb_n1, lv.fb_n2)
                lo, hi = (w0, w0 + w1) if sg.part else (0, w0)
                if sg.flags & 1:               # the issuer waits for the stage, then clears the unit's columns
                    assert unit not in started
                    started.add(unit)
                    D[:, base + lo:base + hi] = 0.0
                assert unit in started and unit not in finished
                for i in range(sg.begin, sg.begin + sg.count):
                    m = lv.mma[i]
                    assert (m.job, m.part) == unit and m.term in (0, 1) and i not in covered
                    assert lo <= m.d_col and m.d_col + m.n <= hi     # an issuer only touches its own unit's columns
                    covered.add(i)
                    # the packed entry the issuer reads must describe the same MMA
                    a_lo, b_lo, idesc, col = (lv.pk[4 * i + j] for j in range(4))
                    assert a_lo == ((m.a_off + m.term * lv.slot_term_bytes) >> 4) | ((lv.a_lbo >> 4) << 16)
                    assert b_lo == (m.b_off >> 4) | (((128 * m.b_rows8) >> 4) << 16)
                    assert idesc == (1 << 4) | ((m.n >> 3) << 17) | (8 << 24)
                    assert col == lv.ring_base[1 if m.job else 0] + m.d_col
                    c0 = base + m.d_col
                    a_idx = (m.a_off + (k // 8) * lv.a_lbo + 16 * r) // 2 + (k % 8)
                    A = slots[g][m.term][a_idx].astype(np.float64)
                    n = np.arange(m.n)[:, None]
                    b_idx = (m.b_off + (k // 8) * 128 * m.b_rows8 + 16 * n) // 2 + (k % 8)
                    B = bimg[b_idx].astype(np.float64)
                    D[:, c0:c0 + m.n] += A @ B.T
                if sg.flags & 2:
                    finished.add(unit)
    assert covered == set(range(lv.n_mma))
    units = {(j, p) for j in range(lv.n_jobs) for p in (0, 1)}                         # two commits per job ...
    if lv.fb_n2 == 0:
        units.discard((0, 1))                    # ... but one for a merged filterbank accumulator (several frames per row)
    assert started == finished == units
    assert sorted(lv.ep_job[: lv.n_jobs]) == list(range(lv.n_jobs))
    return D


def dec_outputs(lv, D):
    cols = []
    for p in range(lv.n_pass):
        base = acc_base(lv, 1 + p)
        w = min(64, lv.dec_w - 64 * p)
        cols.append((D[:, base:base + w] + D[:, base + lv.dec_wp:base + lv.dec_wp + w] / 2048.0) * lv.dec_scale)
    return np.concatenate(cols, axis=1)


@pytest.mark.parametrize("mode", ["vqt", "cqt"])
@pytest.mark.parametrize("level", range(8))
def test_level_plan_replay_matches_oracle(level, mode):
    lv, bimg = level_plan(level, mode)
    R = 8 * lv.q
    rng = np.random.default_rng(level)
    n_sig = R * 128 * 3 + 37
    sig = (0.5 * rng.standard_normal(n_sig)).astype(np.float32)
    row0 = 128                    # an interior tile
    D = replay(lv, bimg, sig, row0)
    # ---- decimator ----
    if lv.dec_w:
        want = vo.resample_2to1_f64(sig)          # sqrt(2) sum h x, zero extended
        got = dec_outputs(lv, D)
        t = (row0 + np.arange(128))[:, None] * lv.dec_w + np.arange(lv.dec_w)[None, :]
        ref = want[t]
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    else:
        assert level == 7
    # ---- filterbank ----
    g, n_fft = vo.octave_time_kernels(level, 16000.0, vo.default_gamma() if mode == "vqt" else 0.0)
    assert n_fft == lv.n_fft
    for j in range(lv.fpr):
        fb0 = acc_base(lv, 0)
        if lv.fb_n2 == 0:         # merged: [g1 of every frame | g2 of every frame], x2 . g1 accumulated onto the g2 columns
            main = D[:, fb0 + 24 * j: fb0 + 24 * j + 24]
            small = D[:, fb0 + 24 * lv.fpr + 24 * j: fb0 + 24 * lv.fpr + 24 * j + 24]
            c = (main + small / 2048.0) * lv.fb_scale
        else:
            a = D[:, fb0 + 48 * j: fb0 + 48 * j + 48]
            b = D[:, fb0 + lv.fb_n1 + 24 * j: fb0 + lv.fb_n1 + 24 * j + 24]
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
    got = dec_outputs(lv, D)
    t = np.arange(128)[:, None] * lv.dec_w + np.arange(lv.dec_w)[None, :]
    n_valid = sig.size // 2
    ok = t < n_valid
    assert np.abs(got[ok] - want[t[ok]]).max() <= 2e-6 * np.abs(want).max()


@pytest.mark.parametrize("mode", ["vqt", "cqt"])
def test_issuers_work_at_every_position_or_not_at_all(mode):
    """Slot hand-off invariant (see build_level): no issuer idles at one position of a tile and works at another."""
    for level in range(8):
        lv, _ = level_plan(level, mode)
        for isr in range(N_ISSUERS):
            per_pos = [lv.seg_begin[isr][p + 1] - lv.seg_begin[isr][p] for p in range(lv.gpt)]
            assert all(per_pos) or not any(per_pos), (mode, level, isr, per_pos)


def test_level_geometry_table():
    rows = {0: (32, 1), 1: (16, 1), 2: (8, 1), 3: (4, 1), 4: (4, 2), 5: (4, 4), 6: (1, 2), 7: (1, 4)}
    for level, (q, fpr) in rows.items():
        lv, _ = level_plan(level)
        assert (lv.q, lv.fpr) == (q, fpr)
        assert lv.n_mma <= MAX_MMA and lv.n_seg <= 64 and lv.q == lv.pg * lv.gpt and lv.n_slots % lv.gpt == 0
        assert lv.ring_base[0] + lv.ring_stages[0] * lv.ring_width[0] <= 512
        assert lv.n_slots * 2 * lv.slot_term_bytes + lv.b_bytes + 256 < 220 * 1024


def test_cqt_geometry_is_covered():
    """gamma = 0 (CQT) needs n_fft = 256 at every octave: twice the filterbank k-steps and wider halos (up to 17 rows at the
    8-sample levels), still inside the level kernels' limits -- CQT runs on the tcgen05 pyramid like VQT."""
    for level in range(8):
        lv, _ = level_plan(level, "cqt")
        assert lv.n_fft == 256 and lv.n_mma <= MAX_MMA and lv.n_seg <= 64
        assert lv.n_slots * 2 * lv.slot_term_bytes + lv.b_bytes + 256 < 220 * 1024
        assert lv.ring_base[0] + lv.ring_stages[0] * lv.ring_width[0] <= 512


def test_unsupported_geometry_is_reported():
    """A hop whose top-level row would not fit the plane layout is refused by the host hook (the plan then keeps the round-1
    kernels) instead of building a wrong plan."""
    lv = VqtLevel()
    with pytest.raises(L.ZnsError, match="not supported"):
        L.check(L.lib().zns_dbg_vqt_level_plan(16000, 1024, 96, 12, vo.FMIN_C0, -1.0, 0, C.byref(lv), C.sizeof(lv), None, 0))
