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

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "config_golden.npz"))


@pytest.fixture(scope="module")
def sd(gold):
    from zeronotesamba_b200.models.checkpoint import he_normal_state_dict
    return he_normal_state_dict(int(gold["ckpt_seed"]))


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm())


def _tied(sd):
    out = {k: v.clone() for k, v in sd.items()}
    for k in list(out):
        if k.startswith("postve."):
            out[k] = out["anchor." + k[len("postve."):]].clone()
    return out


def check_weight_update(gold, tag, model, w0_sd, min_sign=0.99, g_floor=1e-6):
    """Discriminating one-step check on the sampled entries: sign of the update and its size."""
    keys = [str(k) for k in gold["layout_keys"]]
    off = gold["sample_off"]
    new_sd = model.state_dict()
    lr = float(gold[f"{tag}_lr"])
    n_checked = agree = 0
    worst = 0.0
    for i, k in enumerate(keys):
        sl = slice(off[i], off[i + 1])
        idx = gold["sample_idx"][sl]
        w0 = gold[f"{tag}_w0_samples"][sl].astype(np.float64)
        assert np.array_equal(w0_sd[k].reshape(-1)[idx].numpy().astype(np.float64), w0), k
        d_ref = gold[f"{tag}_w1_samples"][sl].astype(np.float64) - w0
        d_got = new_sd[k].reshape(-1)[idx].double().cpu().numpy() - w0
        g_ref = gold[f"{tag}_grad_samples"][sl].astype(np.float64)
        sel = np.abs(g_ref) > g_floor                      # >> eps = 1e-8 and above the bf16 noise of the gradient
        if not sel.any():
            continue
        assert np.all(np.abs(d_ref[sel]) > 0.5 * lr), k     # the reference really moved these by ~lr
        same = np.sign(d_got[sel]) == np.sign(d_ref[sel])
        n_checked += int(sel.sum())
        agree += int(same.sum())
        ulp = np.spacing(np.abs(w0[sel]).astype(np.float32)).astype(np.float64)
        err = np.abs(d_got[sel] - d_ref[sel])[same] - (1e-3 * np.abs(d_ref[sel]) + 2 * ulp)[same]
        if err.size:
            worst = max(worst, float(err.max()))
    assert n_checked > 2000, n_checked
    assert agree / n_checked >= min_sign, (agree, n_checked)
    assert worst <= 0.0, worst
    return agree / n_checked, n_checked


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
    assert _rel(tr.engine.emb[0], gold["cfg3_anc_emb"]) < 1e-2
    assert _rel(tr.engine.emb[1], gold["cfg3_pos_emb"]) < 1e-2
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
    assert _rel(tr.engine.emb[0], gold["cfg3_anc_emb"]) < 1e-2
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
    assert _rel(tr.engine.emb[0], gold["cond_anc_emb"]) < 1e-2 and _rel(tr.engine.emb[1], gold["cond_pos_emb"]) < 1e-2
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
    assert _rel(postve, gold["cfg1_postve"]) < 1e-2
    assert _rel(anchor, gold["cfg1_anchor"]) < 1e-2
    assert _rel(both, gold["cfg1_max"]) < 1e-2
    # per-frame check as well: activations are in (0, 1)
    assert float((both.cpu() - torch.from_numpy(gold["cfg1_max"])).abs().max()) < 1e-2


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
    worst = (0.0, 0.0)
    for i in range(0, 256, 16):
        ref = vo.vqt_ref_f32(y[i].cpu().numpy())
        rel, ab = vqt_check(out[i].cpu().numpy(), ref)
        worst = (max(worst[0], rel), max(worst[1], ab))
        assert rel < 1e-4 and ab < 2e-6, (i, rel, ab)
    print("cfg2 worst rel / abs-over-max:", worst)
    # size-independent property over the whole batch: clips that share a base stem differ by -80 dBFS noise only
    d = (out[0:8] - out[8:16]).abs().max()
    assert float(d) < 1.0
