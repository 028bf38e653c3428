"""cfg2 timing of the VQT front-end alone (256 x 30 s clips); used under ncu for per-kernel times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from zeronotesamba_b200 import synth
from zeronotesamba_b200.processing.input_rep import VQTPlan

B, N = 256, 480000
base = np.stack([synth.stem_pair(i, 30.0)[i % 2] for i in range(4)])
y = torch.from_numpy(base).cuda().repeat(B // 4, 1)
y += 1e-4 * torch.randn_like(y)
plan = VQTPlan(16000, "vqt", B, N)
out = torch.empty(B, 96, 1876, device="cuda")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
timing = len(sys.argv) > 2 and sys.argv[2] == "--timing"
if timing:
    from zeronotesamba_b200 import _lib as L
    dbg = torch.zeros(8, 16, dtype=torch.int64, device="cuda")
    L.check(L.lib().zns_dbg_vqt_timing(L.ptr(dbg)))
for _ in range(2):
    plan.forward(y, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    plan.forward(y, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"cfg2 VQT: {ms:.3f} ms  ({B*30/(ms*1e-3):.3e} audio-s/s, {B*(4*N+4*96*1876)/(ms*1e-3)/1e9:.1f} GB/s algorithmic)")

if timing:
    d = dbg.cpu().numpy()
    print("level: issuer total / wait operands / wait accumulator  [tiles] | epilogue A total / wait | epilogue B total / wait | loader total / wait slot   (cycles, CTA 0)")
    for lv in range(8):
        r = d[lv]
        print(f"  L{lv}: {r[0]:9d} {r[1]:9d} {r[2]:9d} [{r[3]}] | {r[4]:9d} {r[5]:9d} | {r[6]:9d} {r[7]:9d} | {r[8]:9d} {r[9]:9d}  || in umma {r[10]} in commit {r[11]} mmas {r[12]}")
