"""GPU parity at the BASELINE.json configurations against outputs of the reference's own modules
(tests/golden/config_golden.npz, written by tools/make_golden_configs.py from /root/reference):

  cfg1  sample_script.py:31-48   generate_XQT x2 -> Down_CNN three forwards at T = 1876
  cfg2  256 x 30 s batched VQT   16 clips of the exact bench batch against the oracle
  cfg3  pretext.py:475-490       one training step at batch 16, T = 313, temperature 0.25, lr 1e-6
  cond  the same step at a better-conditioned operating point (tied branches, cos+ - cos- = 0.07, lr 1e-3)

Tolerances (BASELINE.json north_star): embeddings and loss 1e-2 relative (bf16 operands, fp32 accumulation); one-step
weight updates: Adam's first step is -lr * g / (|g| + eps), so wherever |g_ref| >> eps the update must have the
reference's SIGN and its size to 1e-3 (+ one ulp of the weight); VQT 1e-4 relative.
"""
import os
import random

import numpy as np
import pytest
import torch

from helpers import check_adam_deltas

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "config_golden.npz"))


@pytest.fixture(scope="module")
def sd(gold):
    from zeronotesamba_b200.models.checkpoint import he_normal_state_dict
    return he_normal_state_dict(int(gold["ckpt_seed"]))


# Embedding tolerance: north_star's 1e-2 relative (L2), plus a per-frame bound on the (0, 1) sigmoid outputs.  With bf16
# forward activations cfg3 measured 1.09e-2 on the anchor branch (just over); the forward pass now keeps its activations and
# weights in fp16 (11 significant bits, same tensor-core rate), which is what these bounds are for.
EMB_REL = 1e-2
EMB_ABS = 1e-2


def _emb_ok(got, ref):
    got, ref = torch.as_tensor(got).double().cpu(), torch.as_tensor(ref).double().cpu()
    rel = float((got - ref).norm() / ref.norm())
    ab = float((got - ref).abs().max())
    print(f"embedding error: L2-relative {rel:.3e}, max abs {ab:.3e}")
    return rel < EMB_REL and ab < EMB_ABS


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm())


def _tied(sd):
    out = {k: v.clone() for k, v in sd.items()}
    for k in list(out):
        if k.startswith("postve."):
            out[k] = out["anchor." + k[len("postve."):]].clone()
    return out


def check_weight_update(gold, tag, model, w0_sd, min_sign=0.99, g_floor=1e-6, min_checked=1000):
    """Discriminating one-step check on the sampled entries (helpers.check_adam_deltas): the update is Adam's step on this
    implementation's gradient (every entry), and it has the reference's sign and size wherever the reference gradient
    stands clear of eps and of the tensor's reduced-precision gradient noise."""
    keys = [str(k) for k in gold["layout_keys"]]
    off = gold["sample_off"]
    new_sd = model.state_dict()
    named = dict(model.named_parameters())
    lr = float(gold[f"{tag}_lr"])
    n_checked = agree = 0.0
    for i, k in enumerate(keys):
        sl = slice(off[i], off[i + 1])
        idx = gold["sample_idx"][sl]
        w0 = gold[f"{tag}_w0_samples"][sl].astype(np.float64)
        assert np.array_equal(w0_sd[k].reshape(-1)[idx].numpy().astype(np.float64), w0), k
        d_ref = gold[f"{tag}_w1_samples"][sl].astype(np.float64) - w0
        d_got = new_sd[k].reshape(-1)[idx].double().cpu().numpy() - w0
        g_ref = gold[f"{tag}_grad_samples"][sl].astype(np.float64)
        g_got = named[k].grad.reshape(-1)[idx].double().cpu().numpy()
        n, frac = check_adam_deltas(d_got, d_ref, g_ref, w0, lr, g_got=g_got, g_floor=g_floor, min_sign=0.8)
        n_checked += n
        agree += n * frac
    assert n_checked > min_checked, n_checked
    assert agree / n_checked >= min_sign, (agree, n_checked)
    return agree / n_checked, int(n_checked)


def grad_stats(gold, tag, model):
    keys = [str(k) for k in gold["layout_keys"]]
    named = dict(model.named_parameters())
    off = gold["sample_off"]
    out = []
    for i, k in enumerate(keys):
        g = named[k].grad
        sl = slice(off[i], off[i + 1])
        gs = g.reshape(-1)[gold["sample_idx"][sl]].double().cpu().numpy()
        ref = gold[f"{tag}_grad_samples"][sl].astype(np.float64)
        cos = float(gs @ ref / (np.linalg.norm(gs) * np.linalg.norm(ref) + 1e-300))
        out.append((k, float(g.double().norm()) / float(gold[f"{tag}_grad_l2"][i]), cos))
    return out


def _batch(gold, tag):
    pair = torch.from_numpy(gold[f"{tag}_vqt"]).to(DEV)
    starts = [int(s) for s in gold["cfg3_starts"]]
    assert starts == random.Random(0).sample(range(0, 313), 16)          # pretext.py:312
    return torch.stack([pair[:, :, s:s + 313] for s in starts]).contiguous()


@pytest.mark.parametrize("use_graph", [False, True])
def test_cfg3_training_step_identical_inputs(gold, sd, use_graph):
    from zeronotesamba_b200.models.models import Pretext_CNN
    from zeronotesamba_b200.pretext import PretextTrainer
    batch = _batch(gold, "cfg3")
    assert batch.shape == (16, 2, 96, 313)
    model = Pretext_CNN().to(DEV)
    model.load_state_dict(sd)
    tr = PretextTrainer(model, batch_len=16, temperature=0.25, lr=1e-6, crop_frames=313, dropout_p=0.0, use_graph=use_graph)
    ev = tr.eval_step(batch).cpu().numpy().copy()
    assert np.allclose(ev, gold["cfg3_val_loss_cos"], rtol=1e-2, atol=2e-3), (ev, gold["cfg3_val_loss_cos"])
    res = tr.step(batch).cpu().numpy()
    want = gold["cfg3_train_loss_cos"]
    assert abs(res[0] - want[0]) <= 1e-2 * abs(want[0])
    # the loss sits 7.5e-3 below ln 16 (a constant-output network gives exactly ln 16): resolve that gap to 25 %
    assert abs((np.log(16.0) - res[0]) - (np.log(16.0) - want[0])) <= 0.25 * (np.log(16.0) - want[0]), (res, want)
    assert abs(res[1] - want[1]) <= 2e-3 and abs(res[2] - want[2]) <= 2e-3, (res, want)
    assert _emb_ok(tr.engine.emb[0], gold["cfg3_anc_emb"])
    assert _emb_ok(tr.engine.emb[1], gold["cfg3_pos_emb"])
    frac, n = check_weight_update(gold, "cfg3", model, sd, min_sign=0.97)
    print(f"cfg3: update sign agreement {frac:.4f} over {n} sampled weights")


def test_cfg3_end_to_end_from_audio(gold, sd):
    """The in-loop path: synthetic stems -> VQT on the GPU -> index-only crops -> training step."""
    from zeronotesamba_b200 import synth
    from zeronotesamba_b200.models.models import Pretext_CNN
    from zeronotesamba_b200.pretext import PretextTrainer
    drums, other = synth.stem_pair(int(gold["cfg3_clip"]), 10.0)
    model = Pretext_CNN().to(DEV)
    model.load_state_dict(sd)
    tr = PretextTrainer(model, batch_len=16, temperature=0.25, lr=1e-6, crop_frames=313, dropout_p=0.0, use_graph=True)
    st = torch.from_numpy(gold["cfg3_starts"]).to(DEV)
    res = tr.step_from_audio(torch.from_numpy(other).to(DEV), torch.from_numpy(drums).to(DEV), st).cpu().numpy()
    # the VQT the step saw equals the oracle VQT the reference was fed
    from helpers import vqt_check
    for c in range(2):
        rel, ab = vqt_check(tr._vqt_buf[c].cpu().numpy(), gold["cfg3_vqt"][c])
        assert rel < 1e-4 and ab < 2e-6, (c, rel, ab)
    want = gold["cfg3_train_loss_cos"]
    assert abs(res[0] - want[0]) <= 1e-2 * abs(want[0]) and abs(res[1] - want[1]) <= 2e-3 and abs(res[2] - want[2]) <= 2e-3
    assert _emb_ok(tr.engine.emb[0], gold["cfg3_anc_emb"])
    check_weight_update(gold, "cfg3", model, sd, min_sign=0.97)


def test_conditioned_step_gradients_and_update(gold, sd):
    """cos+ - cos- = 0.07: the softmax is far from uniform, so end-to-end gradients are well conditioned."""
    from zeronotesamba_b200.models.models import Pretext_CNN
    from zeronotesamba_b200.pretext import PretextTrainer
    tsd = _tied(sd)
    batch = _batch(gold, "cond")
    model = Pretext_CNN().to(DEV)
    model.load_state_dict(tsd)
    lr = float(gold["cond_lr"])
    tr = PretextTrainer(model, batch_len=16, temperature=0.25, lr=lr, crop_frames=313, dropout_p=0.0, use_graph=False)
    res = tr.step(batch).cpu().numpy()
    want = gold["cond_train_loss_cos"]
    assert want[1] - want[2] > 0.05
    assert abs(res[0] - want[0]) <= 1e-2 * abs(want[0]) and abs(res[1] - want[1]) <= 2e-3 and abs(res[2] - want[2]) <= 2e-3
    assert abs((res[1] - res[2]) - (want[1] - want[2])) <= 0.05 * (want[1] - want[2])
    assert _emb_ok(tr.engine.emb[0], gold["cond_anc_emb"]) and _emb_ok(tr.engine.emb[1], gold["cond_pos_emb"])
    stats = grad_stats(gold, "cond", model)
    for k, ratio, cos in stats:
        print(f"cond grad {k}: |g|/|g_ref| {ratio:.4f} cos {cos:.4f}")
    weights = [s for s in stats if s[0].endswith("weight")]
    assert min(c for _, _, c in weights) >= 0.99, stats
    assert max(abs(r - 1.0) for _, r, _ in stats) <= 0.05, stats
    frac, n = check_weight_update(gold, "cond", model, tsd, min_sign=0.99, g_floor=1e-5)
    print(f"cond: update sign agreement {frac:.4f} over {n} sampled weights")


def test_cfg1_sample_script_path(gold, sd):
    """sample_script.py:31-48 with a 30 s clip: generate_XQT on both stems, reshape, Down_CNN three forwards."""
    import zeronotesamba_b200.processing.input_rep as IR
    from helpers import vqt_check
    from zeronotesamba_b200 import synth
    from zeronotesamba_b200.models.models import Down_CNN
    drums, other = synth.stem_pair(int(gold["cfg1_clip"]), 30.0)
    vqt_postve = IR.generate_XQT(drums, 16000, "vqt")
    vqt_anchor = IR.generate_XQT(other, 16000, "vqt")
    assert vqt_anchor.shape == (96, 1876) and vqt_anchor.dtype == np.float32
    cols = gold["cfg1_vqt_anchor_cols"]
    rel, ab = vqt_check(vqt_anchor[:, ::125], cols)
    assert rel < 1e-4 and ab < 2e-6, (rel, ab)
    vqt_postve = torch.reshape(torch.from_numpy(vqt_postve), (1, 1, 96, -1)).to(DEV)
    vqt_anchor = torch.reshape(torch.from_numpy(vqt_anchor), (1, 1, 96, -1)).to(DEV)
    model = Down_CNN().to(DEV)
    model.pretext.load_state_dict(sd)
    model.eval()
    with torch.no_grad():
        postve = model.pretext.postve(vqt_postve)
        anchor = model.pretext.anchor(vqt_anchor)
        both = model(vqt_anchor, vqt_postve)
    assert postve.shape == anchor.shape == both.shape == (1, 1876)
    assert _emb_ok(postve, gold["cfg1_postve"])
    assert _emb_ok(anchor, gold["cfg1_anchor"])
    assert _emb_ok(both, gold["cfg1_max"])


def test_cfg2_bench_batch_against_oracle():
    """BASELINE.json configs[1]: the exact 256 x 30 s batch bench.py times; 16 of its clips against the oracle."""
    from helpers import vqt_check
    from oracle import vqt_oracle as vo
    from zeronotesamba_b200 import synth
    from zeronotesamba_b200.processing.input_rep import VQTPlan
    y = synth.cfg2_batch(DEV)
    assert y.shape == (256, 480000)
    plan = VQTPlan(16000, "vqt", 256, 480000)
    out = plan.forward(y)
    torch.cuda.synchronize()
    assert out.shape == (256, 96, 1876) and bool(torch.isfinite(out).all())
    # the only configuration with many tiles per persistent CTA (ring slots and accumulator stages are reused ~13 times on the
    # small levels): repeated passes must be bit-identical -- a hand-off race shows up here first
    for _ in range(2):
        again = plan.forward(y, out=torch.empty_like(out))
        assert torch.equal(again, out)
    worst = (0.0, 0.0)
    for i in range(0, 256, 16):
        ref = vo.vqt_ref_f32(y[i].cpu().numpy())
        rel, ab = vqt_check(out[i].cpu().numpy(), ref)
        worst = (max(worst[0], rel), max(worst[1], ab))
        assert rel < 1e-4 and ab < 2e-6, (i, rel, ab)
    print("cfg2 worst rel / abs-over-max:", worst)
    # size-independent property over the whole batch: clips that share a base stem differ by -80 dBFS noise only.  The
    # transform is linear before |.|, so in the magnitude domain (the log is ill-conditioned near silence) every replica is
    # within the noise's own response of its base clip: |V_a - V_b| <= |V(noise_a - noise_b)|, a few 1e-4 sqrt(L_k)-normalised
    v = torch.exp(out.double()) - 1e-9
    vmax = float(v.max())
    for r in range(1, 256 // 8):
        d = float((v[0:8] - v[8 * r:8 * r + 8]).abs().max())
        assert d < 2e-3 * vmax, (r, d, vmax)
    assert float((v[0:8] - v[8:16]).abs().max()) > 0.0       # and they are distinct clips
