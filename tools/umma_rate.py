"""tcgen05 rate microbenchmark driver (diagnostic): cycles per 128 x N x 16 bf16 MMA."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zeronotesamba_b200 import _lib as L

def run(n, iters, per_group, mode, ctas):
    cyc = torch.zeros(ctas, dtype=torch.int64, device="cuda")
    L.check(L.lib().zns_dbg_umma_rate(n, iters, per_group, mode, ctas, L.ptr(cyc), L.current_stream()))
    torch.cuda.synchronize()
    c = cyc.float()
    return float(c.mean()) / (iters * per_group), float(c.max()) / (iters * per_group)

if __name__ == "__main__":
    print("n per_group mode ctas -> cycles/MMA (mean, max over CTAs); ideal = n/2")
    for ctas in (1, 148):
        for n in (64, 128, 256):
            for mode in (0, 1, 8, 9):
                for pg in (4, 16):
                    run(n, 50, pg, mode, ctas)
                    m, mx = run(n, 2000, pg, mode, ctas)
                    print(f"n={n:3d} per_group={pg:2d} mode={mode} ctas={ctas:3d}: {m:7.1f} {mx:7.1f}  (ideal {n/2:.0f})")
