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
    dbg = torch.zeros(8 * 32 + 64 + 448, dtype=torch.int64, device="cuda")
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
    tl = dbg.cpu().numpy()[320:768]
    if tl.any():
        t0 = tl[tl > 0].min()
        rel = lambda v: int(v - t0) if v > 0 else -1
        print("L0 timeline of CTA 0, tiles 4..7 (clocks since the first stamp):")
        print("  loader warp 1, tile 5, work done per lane:", [rel(v) for v in tl[384:416]])
        for ti in range(4):
            print(f"  tile {4 + ti}:")
            print("    loaders  [slot wait end -> full arrive] per warp 0..7:", "  ".join(f"{rel(tl[(ti * 8 + w) * 2])}->{rel(tl[(ti * 8 + w) * 2 + 1])}" for w in range(8)))
            print("    loaders  [work done lane 0 / lane 31 / fenced] per warp:", "  ".join(
                f"{rel(tl[280 + (ti * 8 + w) * 3])}/{rel(tl[280 + (ti * 8 + w) * 3 + 1])}/{rel(tl[280 + (ti * 8 + w) * 3 + 2])}" for w in range(8)))
            for isr in range(4):
                print(f"    issuer {isr} [operands ready -> committed] per position:", "  ".join(
                    f"{rel(tl[64 + ((ti * 4 + isr) * 4 + p) * 2])}->{rel(tl[64 + ((ti * 4 + isr) * 4 + p) * 2 + 1])}" for p in range(4)))
            print("    issuer 0 [before operand wait / accumulator stage granted] per position:", "  ".join(
                f"{rel(tl[240 + (ti * 4 + p) * 2])}/{rel(tl[240 + (ti * 4 + p) * 2 + 1])}" for p in range(4)))
            for hf in range(2):
                print(f"    epilogue warp {4 * hf} [accumulator ready -> drained] per job:", "  ".join(
                    f"{rel(tl[192 + ((ti * 2 + hf) * 3 + j) * 2])}->{rel(tl[192 + ((ti * 2 + hf) * 3 + j) * 2 + 1])}" for j in range(3)))
    print("per level (cycles, CTA 0): issuers [total / wait operands / wait accumulator / MMAs] x4 | epilogue A total/wait | B total/wait | loader 0 total/wait slot")
    for lv in range(8):
        r = d[lv]
        iss = "  ".join(f"{r[4*i]}/{r[4*i+1]}/{r[4*i+2]}/{r[4*i+3]}" for i in range(4))
        print(f"  L{lv}: {iss} | {r[16]}/{r[17]} | {r[18]}/{r[19]} | {r[20]}/{r[21]} (loader work {r[22]}) | epilogue A decimator blocks: tmem ld+wait {r[27]} math {r[28]} stores {r[29]}")
