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
    dbg = torch.zeros(8 * 32 + 64, dtype=torch.int64, device="cuda")
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
    d = dbg.cpu().numpy()[:256].reshape(8, 32)
    print('L0 loader 0 timestamps (loop top, after wait, issue start, issue end, wait_all end, convert end, fence end, arrive end):')
    print(dbg.cpu().numpy()[256:304].reshape(6, 8))
    print("per level (cycles, CTA 0): issuers [total / wait operands / wait accumulator / MMAs] x4 | epilogue A total/wait | B total/wait | loader 0 total/wait slot")
    for lv in range(8):
        r = d[lv]
        iss = "  ".join(f"{r[4*i]}/{r[4*i+1]}/{r[4*i+2]}/{r[4*i+3]}" for i in range(4))
        print(f"  L{lv}: {iss} | {r[16]}/{r[17]} | {r[18]}/{r[19]} | {r[20]}/{r[21]} (issue {r[22]} wait_all {r[23]} convert {r[24]} fence {r[25]} arrive {r[26]})")
