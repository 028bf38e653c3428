"""CPU-only checks of the host-side launch geometry of the tensor-core convolutions (zns_dbg_conv_*_plan):
tile plans cover every output row exactly once, CTA pairs are well formed, the weight-gradient work-item table
holds every (tap row, tap group, cin block) exactly once, and everything fits shared memory / TMEM.  The same
host code runs in front of every zns_conv_fwd / zns_conv_wgrad launch (reference layers:
/root/reference/zeroNoteSamba/models/models.py:17-23)."""
import ctypes as C
import itertools

import numpy as np
import pytest

from zeronotesamba_b200 import _lib as L

SMEM_LIMIT = 232448
LAYERS = [  # c_in, c_out, kh, kw, H  (cv2..cv8)
    (64, 64, 7, 13, 96), (64, 128, 5, 15, 32), (128, 128, 9, 17, 32), (128, 256, 3, 19, 8),
    (256, 256, 5, 21, 8), (256, 128, 1, 23, 1), (128, 128, 1, 25, 1),
]


def fwd_plan(B, H, W, ci, co, kh, kw, n_br=2):
    d = L.conv_desc(B, H, W, ci, co, kh, kw)
    out = (C.c_int * 14)()
    L.check(L.lib().zns_dbg_conv_fwd_plan(C.byref(d), n_br, out))
    keys = ["kernel", "N", "ctas", "hb", "nb", "hs", "ns", "n_cols", "n_total", "slots", "stages", "grid_x", "smem", "units"]
    return dict(zip(keys, list(out)))


def wgrad_plan(B, H, W, ci, co, kh, kw, n_br=2):
    d = L.conv_desc(B, H, W, ci, co, kh, kw)
    out = (C.c_int * 16)()
    items = (C.c_uint32 * 256)()
    L.check(L.lib().zns_dbg_conv_wgrad_plan(C.byref(d), n_br, out, items))
    keys = ["NB", "ctas", "slices", "n_acc", "groups", "grp_base", "grp_rem", "rows", "stack_dy", "fold", "cin_blocks",
            "cout_blocks", "pairs", "stages", "grid_x", "smem"]
    p = dict(zip(keys, list(out)))
    p["items"] = [items[i] for i in range(2 * p["pairs"])]
    return p


def check_fwd(p, B, H, W, co, n_br):
    G = (B + 7) // 8
    n_wt = (W + 15) // 16
    assert p["n_cols"] == G * n_wt * n_br
    assert p["hb"] * p["nb"] + p["hs"] * p["ns"] == p["units"], p         # every row (pair) exactly once
    assert p["n_total"] == p["n_cols"] * (p["nb"] + p["ns"]) == p["grid_x"]
    assert 1 <= p["hs"] <= p["hb"] and p["nb"] >= 0 and p["ns"] >= 0 and p["nb"] + p["ns"] >= 1
    assert p["hb"] * p["N"] <= 512, "accumulators exceed TMEM"
    assert p["smem"] <= SMEM_LIMIT and p["slots"] <= 8 and 2 <= p["stages"] <= 8
    if p["kernel"] == 1:
        assert co == 64 and H % 2 == 0 and p["units"] == H // 2 and p["slots"] >= (p["hb"] - 1) * 2 + 2
    else:
        assert p["N"] == co and p["units"] == H and p["slots"] >= p["hb"] + 1
    if p["ctas"] == 2:
        assert (G * n_wt) % 2 == 0 and p["grid_x"] % 2 == 0 and (p["n_cols"] * p["nb"]) % 2 == 0
    else:
        assert p["ctas"] == 1


@pytest.mark.parametrize("ci,co,kh,kw,H", LAYERS)
def test_forward_and_dgrad_plans_of_the_reference_layers(ci, co, kh, kw, H):
    for B, W, n_br in [(16, 313, 2), (8, 48, 2), (1, 371, 1), (5, 37, 1), (16, 626, 2)]:
        check_fwd(fwd_plan(B, H, W, ci, co, kh, kw, n_br), B, H, W, co, n_br)      # forward
        check_fwd(fwd_plan(B, H, W, co, ci, kh, kw, n_br), B, H, W, ci, n_br)      # data gradient (roles swapped)
    p = fwd_plan(16, H, 313, ci, co, kh, kw, 2)
    assert p["ctas"] == 2, "training shape must run the CTA-pair kernels"
    assert p["kernel"] == (1 if co == 64 else 0)


def test_forward_plans_random_shapes():
    rng = np.random.default_rng(0)
    for _ in range(300):
        ci = int(rng.choice([64, 128, 256])); co = int(rng.choice([64, 128, 256]))
        kh = int(rng.choice([1, 3, 5, 7, 9])); kw = int(rng.choice([1, 3, 11, 17, 25, 41]))
        B = int(rng.integers(1, 33)); H = int(rng.integers(1, 100)); W = int(rng.integers(1, 700))
        n_br = int(rng.integers(1, 3))
        check_fwd(fwd_plan(B, H, W, ci, co, kh, kw, n_br), B, H, W, co, n_br)


def check_wgrad(p, B, H, W, ci, co, kh, kw, n_br):
    unit = 2 if ci == 64 else 1
    row_acc = (kw + unit - 1) // unit
    assert p["fold"] == (1 if ci == 64 else 0)
    assert p["groups"] * p["grp_base"] + p["grp_rem"] == row_acc            # balanced split of a filter row
    assert p["n_acc"] == p["grp_base"] + (1 if p["grp_rem"] else 0) and p["n_acc"] * p["NB"] <= 512
    assert p["rows"] == kh + p["stack_dy"] and p["smem"] <= SMEM_LIMIT and 2 <= p["stages"] <= 8
    assert p["cin_blocks"] == (1 if ci == 64 else ci // 128)
    assert p["cout_blocks"] == (1 if p["stack_dy"] else co // p["NB"])
    assert p["slices"] >= 1
    if p["ctas"] == 1:
        assert p["grid_x"] == p["rows"] * p["groups"] * p["cin_blocks"] * p["cout_blocks"] * p["slices"]
        return
    # CTA pairs: the table holds every (row item, tap group, cin block) exactly once among its valid entries
    items = p["items"]
    assert p["NB"] == 128 and len(items) == 2 * p["pairs"] and p["grid_x"] == len(items) * p["cout_blocks"] * p["slices"]
    dec = [(it & 15, (it >> 4) & 63, (it >> 10) & 15, (it >> 14) & 15, (it >> 18) & 1) for it in items]
    valid = [(r, s0, na, cib) for r, s0, na, cib, v in dec if v]
    want = set()
    for sg in range(p["groups"]):
        acc0 = sg * p["grp_base"] + min(sg, p["grp_rem"])
        na = p["grp_base"] + (1 if sg < p["grp_rem"] else 0)
        for r, cib in itertools.product(range(p["rows"]), range(p["cin_blocks"])):
            want.add((r, acc0 * unit, na, cib))
    assert len(valid) == len(set(valid)) and set(valid) == want
    assert len(dec) - len(valid) <= 2, "at most one dummy per accumulator class"
    for a, b in zip(dec[0::2], dec[1::2]):
        assert a[2] == b[2], "both CTAs of a pair issue the same number of accumulators"
        if not b[4]:
            assert a[4] and a[:4] == b[:4], "a dummy repeats its (valid) partner"
        assert a[4], "the leader of a pair is never the dummy"
    accs = [d_[2] for d_ in dec]
    assert accs == sorted(accs, reverse=True), "bigger class first (longest-first dispatch)"
    # taps of a row: the groups tile [0, kw) (the last accumulator of a folded row may hold one padding tap)
    for r, cib in itertools.product(range(p["rows"]), range(p["cin_blocks"])):
        spans = sorted((s0, s0 + na * unit) for rr, s0, na, cc in valid if rr == r and cc == cib)
        assert spans[0][0] == 0 and all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and kw <= spans[-1][1] <= kw + unit - 1


@pytest.mark.parametrize("ci,co,kh,kw,H", LAYERS)
def test_weight_gradient_plans_of_the_reference_layers(ci, co, kh, kw, H):
    for B, W, n_br in [(16, 313, 2), (8, 48, 2), (1, 371, 1), (5, 37, 1)]:
        p = wgrad_plan(B, H, W, ci, co, kh, kw, n_br)
        check_wgrad(p, B, H, W, ci, co, kh, kw, n_br)
        assert p["ctas"] == 2                           # CTA pairs everywhere; c_out = 64 runs with dy rows stacked on N = 128
        assert p["stack_dy"] == (1 if co == 64 else 0) and p["NB"] == 128


def test_weight_gradient_plans_random_shapes():
    rng = np.random.default_rng(1)
    for _ in range(300):
        ci = int(rng.choice([64, 128, 256, 512])); co = int(rng.choice([64, 128, 256]))
        kh = int(rng.choice([1, 3, 5, 7, 9, 15])); kw = int(rng.choice([1, 3, 11, 13, 17, 25, 41]))
        B = int(rng.integers(1, 33)); H = int(rng.integers(1, 100)); W = int(rng.integers(1, 700))
        n_br = int(rng.integers(1, 3))
        check_wgrad(wgrad_plan(B, H, W, ci, co, kh, kw, n_br), B, H, W, ci, co, kh, kw, n_br)


def test_plan_argument_errors():
    out = (C.c_int * 16)()
    for bad in [L.conv_desc(16, 8, 40, 96, 128, 3, 5), L.conv_desc(16, 8, 40, 64, 128, 4, 5), L.conv_desc(16, 0, 40, 64, 128, 3, 5)]:
        assert L.lib().zns_dbg_conv_fwd_plan(C.byref(bad), 1, out) != 0
        assert L.lib().zns_dbg_conv_wgrad_plan(C.byref(bad), 1, out, None) != 0
    assert L.lib().zns_dbg_conv_fwd_plan(C.byref(L.conv_desc(16, 8, 40, 64, 192, 3, 5)), 1, out) != 0   # c_out not 64/128/256
    assert L.lib().zns_dbg_conv_fwd_plan(C.byref(L.conv_desc(16, 8, 40, 64, 64, 3, 5)), 3, out) != 0    # n_br


@pytest.mark.parametrize("env", [{"ZNS_CONV_PAIR": "0"}, {"ZNS_CONV_PAIR": "1"}, {"ZNS_WGRAD_STACK": "1"}, {"ZNS_WGRAD_STACK": "0"},
                                 {"ZNS_CONV_NO_STACK": "1"}])
def test_plans_of_the_kernel_variants(env):
    """The A/B switches are read once per process, so the variants are checked in a child process."""
    import os
    import subprocess
    import sys
    code = (
        "import importlib.util, sys\n"
        f"spec = importlib.util.spec_from_file_location('plans', {os.path.abspath(__file__)!r})\n"
        "m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)\n"
        "for ci, co, kh, kw, H in m.LAYERS:\n"
        "    for B, W, n_br in [(16, 313, 2), (8, 48, 2), (1, 371, 1)]:\n"
        "        m.check_fwd(m.fwd_plan(B, H, W, ci, co, kh, kw, n_br), B, H, W, co, n_br)\n"
        "        m.check_fwd(m.fwd_plan(B, H, W, co, ci, kh, kw, n_br), B, H, W, ci, n_br)\n"
        "        m.check_wgrad(m.wgrad_plan(B, H, W, ci, co, kh, kw, n_br), B, H, W, ci, co, kh, kw, n_br)\n"
        "p = m.wgrad_plan(16, 96, 313, 64, 64, 7, 13, 2); f = m.fwd_plan(16, 32, 313, 128, 128, 9, 17, 2)\n"
        "print('RESULT', p['stack_dy'], p['ctas'], p['NB'], f['ctas'], m.fwd_plan(16, 96, 313, 64, 64, 7, 13, 2)['kernel'])\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], env={**os.environ, **env, "PYTHONPATH": root}, capture_output=True,
                         text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    stack_dy, w_ctas, nb, f_ctas, kern = map(int, res.stdout.split("RESULT")[1].split())
    if env.get("ZNS_CONV_PAIR") == "0":
        assert (w_ctas, f_ctas) == (1, 1)
    if env.get("ZNS_CONV_PAIR") == "1":
        assert (w_ctas, f_ctas) == (1, 2)
    if env.get("ZNS_WGRAD_STACK") == "1":
        assert (stack_dy, w_ctas, nb) == (1, 2, 128)
    if env.get("ZNS_WGRAD_STACK") == "0":
        assert (stack_dy, nb) == (0, 64)
    assert kern == (0 if "ZNS_CONV_NO_STACK" in env else 1)
