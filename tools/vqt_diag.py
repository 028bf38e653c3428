"""Per-octave error of the GPU VQT against the oracle (f32-faithful) and the float64 truth, for the default path and
ZNS_VQT_LEGACY=1 (run each in its own process: the switch is read once)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import vqt_oracle as vo
from zeronotesamba_b200 import synth
from zeronotesamba_b200.processing.input_rep import VQTPlan

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
for clip in (5, 0):
    y = synth.stem_pair(clip, secs)[1]
    n = y.size
    plan = VQTPlan(16000, "vqt", 1, n)
    out = plan.forward(torch.from_numpy(y[None]).cuda()).cpu().numpy()[0]
    ref = vo.vqt_ref_f32(y)
    tru = vo.vqt_truth_f64(y)
    v, r, t = (np.exp(a.astype(np.float64)) - 1e-9 for a in (out, ref, tru))
    mx = t.max()
    print(f"clip {clip} legacy={os.environ.get('ZNS_VQT_LEGACY')} max|V|={mx:.3f}")
    for o in range(8):
        rows = slice(96 - 12 * (o + 1), 96 - 12 * o)
        e_ref = np.abs(v[rows] - r[rows]).max() / mx
        e_tru = np.abs(v[rows] - t[rows]).max() / mx
        e_rt = np.abs(r[rows] - t[rows]).max() / mx
        big = t[rows] >= 1e-2 * mx
        rel_tru = (np.abs(v[rows] - t[rows])[big] / t[rows][big]).max() if big.any() else 0
        rel_rt = (np.abs(r[rows] - t[rows])[big] / t[rows][big]).max() if big.any() else 0
        print(f"  octave {o}: |gpu-oracle32|/max {e_ref:.2e}  |gpu-truth64|/max {e_tru:.2e}  |oracle32-truth64|/max {e_rt:.2e}   rel(gpu,truth) {rel_tru:.2e} rel(oracle32,truth) {rel_rt:.2e}")
