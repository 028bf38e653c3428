"""One small invocation of the hot path on cuda:0, checked against the oracle
(used by __graft_entry__.smoke(); the oracle is only the checker here)."""
from __future__ import annotations

import numpy as np
import torch


def run(verbose: bool = True) -> dict:
    from oracle import encoder_oracle as eo
    from oracle import vqt_oracle as vo
    from . import synth
    from .models.checkpoint import he_normal_state_dict
    from .models.models import Pretext_CNN
    from .pretext import PretextTrainer
    from .processing import input_rep as IR

    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    out = {}
    # 1. VQT of a synthetic 2 s stem pair through the reference-facing call
    drums, other = synth.stem_pair(0, 2.0)
    v_other = IR.generate_XQT(other, 16000, "vqt")
    v_drums = IR.generate_XQT(drums, 16000, "vqt")
    ref = vo.vqt_ref_f32(other)
    assert v_other.shape == ref.shape == (96, 126) and v_other.dtype == np.float32
    mag, rmag = np.exp(v_other.astype(np.float64)), np.exp(ref.astype(np.float64))
    big = rmag >= 1e-2 * rmag.max()
    out["vqt_max_rel"] = float((np.abs(mag - rmag)[big] / rmag[big]).max())
    assert out["vqt_max_rel"] < 1e-4, out
    # 2. eight 48-frame crops of that clip, one pretext training step (dropout off) vs the oracle
    B, T = 8, 48
    sd = he_normal_state_dict(7)
    model = Pretext_CNN().to(dev)
    model.load_state_dict(sd)
    pair = torch.from_numpy(np.stack([v_other, v_drums])).to(dev)          # anchor = other, positive = drums
    starts = [0, 9, 21, 30, 44, 57, 66, 78]
    batch = torch.stack([pair[:, :, s:s + T] for s in starts]).contiguous()
    tr = PretextTrainer(model, batch_len=B, temperature=0.25, lr=1e-6, crop_frames=T, dropout_p=0.0, use_graph=False)
    res = tr.step(batch).cpu().numpy()
    torch.set_num_threads(max(1, torch.get_num_threads()))
    want = eo.pretext_step(sd, batch.cpu(), batch_len=B, temperature=0.25, lr=1e-6)
    out["loss"], out["loss_ref"] = float(res[0]), want["loss"]
    assert abs(res[0] - want["loss"]) <= 1e-2 * abs(want["loss"]), out
    emb = tr.engine.emb[0].cpu()
    out["emb_rel"] = float((emb - want["anc_emb"]).norm() / want["anc_emb"].norm())
    assert out["emb_rel"] < 1e-2, out
    # one-step weight update: Adam's first step is -lr * g / (|g| + eps); compare the DELTAS (the weights themselves
    # are ~1e-2, the deltas ~1e-6) wherever the oracle gradient is well above eps: same sign, same size
    new_sd = model.state_dict()
    named = dict(model.named_parameters())
    n_checked = n_same = 0
    for k in ("anchor.pretrained.cv4.weight", "postve.pretrained.cv6.weight", "anchor.pretrained.cv2.weight"):
        w0 = sd[k].double().reshape(-1)
        d_got = (new_sd[k].cpu().double().reshape(-1) - w0)
        d_ref = (want["new_sd"][k].double().reshape(-1) - w0)
        g_ref = want["grads"][k].reshape(-1).double()
        g_got = named[k].grad.detach().cpu().double().reshape(-1)
        ulp = torch.from_numpy(np.spacing(np.abs(sd[k].reshape(-1).numpy()))).double()
        # (A) the applied update is Adam's first step on this implementation's own gradient, every entry
        step = -1e-6 * g_got / (g_got.abs() + 1e-8)
        assert bool(((d_got - step).abs() <= 1e-3 * 1e-6 + 2 * ulp).all()), k
        # (B) against the oracle's update where its gradient stands clear of eps and of the reduced-precision gradient noise
        floor = max(1e-6, 10.0 * float((g_got - g_ref).pow(2).mean().sqrt()))
        sel = g_ref.abs() > floor
        same = torch.sign(d_got[sel]) == torch.sign(d_ref[sel])
        n_checked += int(sel.sum())
        n_same += int(same.sum())
        slack = 1e-6 * (1e-8 / g_ref[sel].abs()) * 0.5      # the update's own sensitivity to the gradient entry (|dg / g| <= 0.5)
        ok = ((d_got[sel] - d_ref[sel]).abs() <= 1e-3 * d_ref[sel].abs() + 2 * ulp[sel] + slack)[same]
        assert float(ok.double().mean()) >= 0.999, (k, float(ok.double().mean()))   # whole tensors: allow 1 in 1000 noise outliers
    out["update_sign_agreement"] = n_same / max(n_checked, 1)
    assert n_checked > 1000 and out["update_sign_agreement"] > 0.95, out
    if verbose:
        print("smoke:", out)
    return out
