"""Diagnostic: cfg2 VQT output of the current library against another build (ZNS_LIB_PATH of a child process), per octave."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch

def run(path):
    from zeronotesamba_b200 import synth
    from zeronotesamba_b200.processing.input_rep import VQTPlan
    y = synth.cfg2_batch("cuda")
    plan = VQTPlan(16000, "vqt", 256, 480000)
    outs = []
    for _ in range(3):
        outs.append(plan.forward(y).clone())
    torch.cuda.synchronize()
    print("run-to-run max |diff|:", float((outs[0] - outs[1]).abs().max()), float((outs[1] - outs[2]).abs().max()))
    np.save(path, outs[2].cpu().numpy())

if len(sys.argv) > 2 and sys.argv[1] == "--child":
    run(sys.argv[2])
    sys.exit(0)
other = sys.argv[1]
run("/tmp/vqt_new.npy")
env = dict(os.environ, ZNS_LIB_PATH=other)
subprocess.run([sys.executable, __file__, "--child", "/tmp/vqt_old.npy"], env=env, check=True)
a, b = np.load("/tmp/vqt_new.npy"), np.load("/tmp/vqt_old.npy")
va, vb = np.exp(a.astype(np.float64)), np.exp(b.astype(np.float64))
scale = vb.max()
for lvl in range(8):
    sl = slice(96 - 12 * (lvl + 1), 96 - 12 * lvl)
    d = np.abs(va[:, sl] - vb[:, sl]) / scale
    bad = np.argwhere(d > 1e-6)
    print(f"level {lvl}: max |dV|/max = {d.max():.3e}, entries > 1e-6: {len(bad)}")
    if len(bad):
        clips = np.unique(bad[:, 0]); frames = bad[:, 2]
        print("   clips:", clips[:20], "... frames min/max:", frames.min(), frames.max(), " first:", bad[:6].tolist())
        fr = np.unique(frames)
        print("   distinct frames:", len(fr), fr[:40])
