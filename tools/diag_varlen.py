"""Localise a stale-workspace dependence: forward the same clip through a fresh engine and through one that has seen other lengths."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from zeronotesamba_b200.engine import EncoderEngine
from zeronotesamba_b200.models.checkpoint import he_normal_state_dict
from zeronotesamba_b200.models.models import Pretext_CNN
dev = torch.device("cuda")
m = Pretext_CNN().to(dev); m.load_state_dict(he_normal_state_dict(7))
names = [n for n, _ in m.anchor.named_parameters()]
params = [dict(m.anchor.named_parameters()), dict(m.postve.named_parameters())]
g = torch.Generator().manual_seed(1)
def run(eng, T, x):
    eng.set_T(T)
    eng.pack_weights(params, need_dgrad=False)
    eng.forward([x[:, 0], x[:, 1]], 2 * 96 * T, params, train=False)
    torch.cuda.synchronize()
    return {k: [t.clone() for t in getattr(eng, k)] for k in ("x1", "p2", "x3", "p4", "x5", "y6", "p6", "x7", "x8", "emb")}
B = 8
xs = {T: (torch.rand(B, 2, 96, T, generator=g) * 10 - 9).to(dev) for T in (400, 170, 333)}
with torch.no_grad():
    fresh = {T: run(EncoderEngine(B, T, 2, dev), T, xs[T]) for T in xs}
    eng = EncoderEngine(B, 400, 2, dev)
    for T in (400, 170, 333, 400, 170):
        got = run(eng, T, xs[T])
        bad = [k for k in got if not all(torch.equal(a, b) for a, b in zip(got[k], fresh[T][k]))]
        print(f"T={T}: tensors differing from a fresh engine: {bad}")
        for k in bad[:2]:
            d = (got[k][0].float() - fresh[T][k][0].float()).abs()
            idx = d.nonzero()
            print("   ", k, "n_diff", int((d > 0).sum()), "of", d.numel(), "first idx", idx[0].tolist() if len(idx) else None, "last", idx[-1].tolist() if len(idx) else None, "max", float(d.max()))
