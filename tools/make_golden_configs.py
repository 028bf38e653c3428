"""Generate tests/golden/config_golden.npz: the REFERENCE's own modules at the BASELINE.json configurations.

Run in the build container only (needs /root/reference):
    python tools/make_golden_configs.py
Uses the reference's ``Pretext_CNN`` / ``Down_CNN`` / ``NTXent`` / ``pretext.train_epoch`` / ``val_epoch`` unmodified
(imported as in tools/make_golden.py) on inputs made by the VQT oracle from the synthetic stems:

  cfg1  sample_script.py:31-48 -- one 30 s stem pair -> VQT (1,1,96,1876) -> Down_CNN three forwards
  cfg3  pretext.py:308-321,475-490 -- one 10 s stem pair -> VQT (2,96,626) -> 16 crops at
        random.Random(0).sample(range(313),16) -> one training step at batch 16, temperature 0.25, lr 1e-6
  cond  the same step at a better-conditioned operating point: tied branch weights and an anchor stem that
        contains the drums (cos+ - cos- > 0.1), lr 1e-3, so that the end-to-end gradient check can be tight and the
        weight deltas dominate rounding.
Dropout is p = 0 (torch's Philox stream cannot be matched).
"""
import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

N_SAMPLE = 256      # sampled entries per tensor for gradient / delta checks


def tied_state_dict(sd):
    out = {k: v.clone() for k, v in sd.items()}
    for k in list(out):
        if k.startswith("postve."):
            out[k] = out["anchor." + k[len("postve."):]].clone()
    return out


def run_step(mm, lf, pt, sd, batch, lr, tag, out, keys, samp_idx):
    model = mm.Pretext_CNN()
    model.load_state_dict(sd)
    for br in (model.anchor, model.postve):
        br.pretrained.dp.p = 0.0
    crit = lf.NTXent(batch_len=batch.shape[0], temperature=0.25)
    opt = torch.optim.Adam(params=model.parameters(), lr=lr)
    loader = [[batch]]
    t0 = time.time()
    vl, vp, vn = pt.val_epoch(model, loader, crit, opt)
    out[f"{tag}_val_loss_cos"] = np.array([vl, vp, vn])
    model.train()
    with torch.no_grad():
        a, p = model(batch[:, 0:1], batch[:, 1:2])
    out[f"{tag}_anc_emb"] = a.numpy()
    out[f"{tag}_pos_emb"] = p.numpy()
    _, tl, tp, tn = pt.train_epoch(model, loader, crit, opt)
    out[f"{tag}_train_loss_cos"] = np.array([tl, tp, tn])
    params = dict(model.named_parameters())
    out[f"{tag}_grad_l2"] = np.array([float(params[k].grad.double().norm()) for k in keys])
    out[f"{tag}_grad_samples"] = np.concatenate([params[k].grad.reshape(-1)[samp_idx[k]].numpy() for k in keys])
    out[f"{tag}_w0_samples"] = np.concatenate([sd[k].reshape(-1)[samp_idx[k]].numpy() for k in keys])
    out[f"{tag}_w1_samples"] = np.concatenate([params[k].detach().reshape(-1)[samp_idx[k]].numpy() for k in keys])
    out[f"{tag}_lr"] = np.array(lr)
    print(tag, "loss/cos", tl, tp, tn, "val", vl, vp, vn, f"{time.time() - t0:.0f} s", flush=True)


def main():
    import make_golden as mg
    from oracle import encoder_oracle as eo
    from oracle import vqt_oracle as vo
    from zeronotesamba_b200 import synth

    torch.set_num_threads(8)
    mm, lf, pt = mg.import_reference()
    out = {}
    sd = eo.he_normal_state_dict(seed=7)
    keys = list(sd.keys())
    rs = np.random.default_rng(17)
    samp_idx = {k: rs.integers(0, sd[k].numel(), size=min(N_SAMPLE, sd[k].numel())) for k in keys}
    out["ckpt_seed"] = np.array(7)
    out["sample_idx"] = np.concatenate([samp_idx[k] for k in keys])
    out["sample_off"] = np.cumsum([0] + [len(samp_idx[k]) for k in keys])
    out["layout_keys"] = np.array(keys)

    # ---- cfg3 ------------------------------------------------------------------------------
    clip = 3
    drums, other = synth.stem_pair(clip, 10.0)
    starts = random.Random(0).sample(range(0, 313), 16)            # pretext.py:312
    out["cfg3_clip"] = np.array(clip)
    out["cfg3_starts"] = np.array(starts, dtype=np.int32)
    pair = np.stack([vo.vqt_ref_f32(other), vo.vqt_ref_f32(drums)])   # anchor = other stems, positive = drums
    assert pair.shape == (2, 96, 626)
    out["cfg3_vqt"] = pair
    batch = torch.from_numpy(np.stack([pair[:, :, s:s + 313] for s in starts]))
    run_step(mm, lf, pt, sd, batch, 1e-6, "cfg3", out, keys, samp_idx)

    # ---- better-conditioned step -------------------------------------------------------------
    anchor_sig = np.clip(drums + 0.05 * other, -0.9, 0.9).astype(np.float32)
    out["cond_anchor_mix"] = np.array(0.05)
    pair_c = np.stack([vo.vqt_ref_f32(anchor_sig), vo.vqt_ref_f32(drums)])
    out["cond_vqt"] = pair_c
    batch_c = torch.from_numpy(np.stack([pair_c[:, :, s:s + 313] for s in starts]))
    run_step(mm, lf, pt, tied_state_dict(sd), batch_c, 1e-3, "cond", out, keys, samp_idx)

    # ---- cfg1: sample_script.py:31-48 at T = 1876 ---------------------------------------------
    d30, o30 = synth.stem_pair(1, 30.0)
    out["cfg1_clip"] = np.array(1)
    vq_anchor = torch.from_numpy(vo.vqt_ref_f32(o30)).reshape(1, 1, 96, -1)
    vq_postve = torch.from_numpy(vo.vqt_ref_f32(d30)).reshape(1, 1, 96, -1)
    assert vq_anchor.shape[-1] == 1876
    down = mm.Down_CNN()
    down.pretext.load_state_dict(sd)
    down.eval()
    with torch.no_grad():
        out["cfg1_postve"] = down.pretext.postve(vq_postve).numpy()
        out["cfg1_anchor"] = down.pretext.anchor(vq_anchor).numpy()
        out["cfg1_max"] = down(vq_anchor, vq_postve).numpy()
    # a few columns of the oracle VQT, so that a mismatch can be attributed
    out["cfg1_vqt_anchor_cols"] = vq_anchor[0, 0, :, ::125].numpy()

    path = os.path.join(ROOT, "tests", "golden", "config_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
