"""CQT (gamma = 0, n_fft 256 per octave) on the tcgen05 level kernels: 128 x 30 s clips, repeated passes bit-identical, and
within float32 rounding of the round-1 kernels (ZNS_VQT_LEGACY=1, run in a child process); time of both paths."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch

B, N = 128, 480000


def run(path):
    from zeronotesamba_b200 import synth
    from zeronotesamba_b200.processing.input_rep import VQTPlan
    y = synth.cfg2_batch("cuda")[:B].contiguous()
    plan = VQTPlan(16000, "cqt", B, N)
    outs = [plan.forward(y).clone() for _ in range(3)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        plan.forward(y, out=outs[0])
    e1.record()
    torch.cuda.synchronize()
    print(f"{'legacy' if os.environ.get('ZNS_VQT_LEGACY') else 'level kernels'}: run-to-run max |diff| "
          f"{float((outs[1] - outs[2]).abs().max())}, {e0.elapsed_time(e1) / 10:.3f} ms per pass of {B} clips")
    np.save(path, outs[2].cpu().numpy())


if len(sys.argv) > 2 and sys.argv[1] == "--child":
    run(sys.argv[2])
    sys.exit(0)
run("/tmp/cqt_new.npy")
subprocess.run([sys.executable, __file__, "--child", "/tmp/cqt_old.npy"], env=dict(os.environ, ZNS_VQT_LEGACY="1"), check=True)
a, b = np.exp(np.load("/tmp/cqt_new.npy").astype(np.float64)), np.exp(np.load("/tmp/cqt_old.npy").astype(np.float64))
d = np.abs(a - b) / b.max()
print("max |dV| / max V per octave:", [f"{d[:, 96 - 12 * (l + 1):96 - 12 * l].max():.2e}" for l in range(8)])
assert d.max() < 5e-6, d.max()
print("CQT_CHECK_OK")
