"""Generate tests/golden/encoder_golden.npz by running the REFERENCE's own modules.

Run in the build container only (needs /root/reference; the GPU box does not have it):
    python tools/make_golden.py
Imports ``zeroNoteSamba.models.models``, ``zeroNoteSamba.models.loss_functions`` and
``zeroNoteSamba.pretext.train_epoch / val_epoch`` from /root/reference unmodified (third-party
modules the reference imports at module scope but that are absent here -- librosa, matplotlib,
spleeter, soundfile -- are stubbed in ``sys.modules``; ``pretext.device0/device1`` are pointed at
the CPU), feeds them seeded inputs and a synthetic He-normal checkpoint, and stores the inputs
and the reference's outputs.  Dropout is set to p=0 for the training step (torch's Philox stream
cannot be matched by another implementation).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def import_reference():
    _stub("librosa", note_to_hz=lambda *_: 0.0)
    _stub("librosa.display")
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot", rcParams={})
    mpl.pyplot = plt
    _stub("spleeter")
    _stub("spleeter.separator", Separator=object)
    _stub("soundfile")
    import zeroNoteSamba.models.loss_functions as lf
    import zeroNoteSamba.models.models as mm
    import zeroNoteSamba.pretext as pt
    pt.device0 = pt.device1 = torch.device("cpu")
    return mm, lf, pt


def main():
    from oracle import encoder_oracle as eo

    torch.set_num_threads(8)
    mm, lf, pt = import_reference()
    out = {}

    # ---- layout ---------------------------------------------------------------------------
    ref_model = mm.Pretext_CNN()
    ref_layout = {k: tuple(v.shape) for k, v in ref_model.state_dict().items()}
    assert list(ref_layout.items()) == list(eo.state_dict_layout().items()), "state_dict layout differs"
    out["layout_keys"] = np.array(list(ref_layout.keys()))
    out["layout_numel"] = np.array([int(np.prod(s)) for s in ref_layout.values()])
    down = mm.Down_CNN()
    assert list(down.state_dict().keys()) == ["pretext." + k for k in ref_layout]

    # ---- synthetic checkpoint ----------------------------------------------------------------
    sd = eo.he_normal_state_dict(seed=7)
    ref_model.load_state_dict(sd)
    out["ckpt_seed"] = np.array(7)
    out["ckpt_checksum"] = np.array([float(sum(v.double().sum() for v in sd.values())),
                                     float(sum((v.double() ** 2).sum() for v in sd.values()))])

    # ---- training step (B=8, T=48) ------------------------------------------------------------
    B, T = 8, 48
    g = torch.Generator().manual_seed(11)
    # VQT-like log-magnitudes: mostly in [-12, 2]
    batch = (torch.rand(B, 2, 96, T, generator=g) * 10.0 - 9.0) + 2.0 * torch.randn(B, 2, 96, T, generator=g)
    batch = batch.float()
    out["step_batch"] = batch.numpy()
    for br in (ref_model.anchor, ref_model.postve):
        br.pretrained.dp.p = 0.0
    crit = lf.NTXent(batch_len=B, temperature=0.25)
    opt = torch.optim.Adam(params=ref_model.parameters(), lr=0.000001)
    loader = [[batch]]
    # forward-only numbers first (val_epoch, model.eval())
    vl, vp, vn = pt.val_epoch(ref_model, loader, crit, opt)
    out["val_loss_cos"] = np.array([vl, vp, vn])
    ref_model.train()
    with torch.no_grad():
        a, p = ref_model(batch[:, 0:1], batch[:, 1:2])
    out["step_anc_emb"] = a.numpy()
    out["step_pos_emb"] = p.numpy()
    # hook grads
    _, tl, tp, tn = pt.train_epoch(ref_model, loader, crit, opt)
    out["train_loss_cos"] = np.array([tl, tp, tn])
    keys = list(ref_layout.keys())
    params = dict(ref_model.named_parameters())
    out["grad_l2"] = np.array([float(params[k].grad.double().norm()) for k in keys])
    out["grad_sum"] = np.array([float(params[k].grad.double().sum()) for k in keys])
    rs = np.random.default_rng(5)
    samp_idx = {k: rs.integers(0, params[k].numel(), size=min(64, params[k].numel())) for k in keys}
    out["sample_idx"] = np.concatenate([samp_idx[k] for k in keys])
    out["sample_off"] = np.cumsum([0] + [len(samp_idx[k]) for k in keys])
    out["grad_samples"] = np.concatenate([params[k].grad.reshape(-1)[samp_idx[k]].numpy() for k in keys])
    out["delta_samples"] = np.concatenate(
        [(params[k].detach().reshape(-1)[samp_idx[k]].double() - sd[k].reshape(-1)[samp_idx[k]].double()).numpy()
         for k in keys])
    out["delta_l2"] = np.array([float((params[k].detach().double() - sd[k].double()).norm()) for k in keys])

    # ---- Down_CNN forward (sample_script.py:38-48), T=40, B=2 ---------------------------------
    down = mm.Down_CNN()
    down.pretext.load_state_dict(sd)
    down.eval()
    x = (torch.rand(2, 2, 96, 40, generator=g) * 10.0 - 9.0).float()
    out["down_in"] = x.numpy()
    with torch.no_grad():
        out["down_postve"] = down.pretext.postve(x[:, 1:2]).numpy()
        out["down_anchor"] = down.pretext.anchor(x[:, 0:1]).numpy()
        out["down_max"] = down(x[:, 0:1], x[:, 1:2]).numpy()
        out["down_mean"] = _with_sd(mm.Down_CNN("mean"), sd)(x[:, 0:1], x[:, 1:2]).numpy()

    # ---- downstream fine-tune step, batch size 1 (epochs.py:45-63): Down_CNN -> BCELoss -> backward -----
    T1 = 400
    x1 = (torch.rand(1, 2, 96, T1, generator=g) * 10.0 - 9.0 + 2.0 * torch.randn(1, 2, 96, T1, generator=g)).float()
    msk = torch.tensor(np.random.default_rng(3).choice([0.0, 0.5, 1.0], size=(1, T1), p=[0.8, 0.1, 0.1]), dtype=torch.float32)
    ft = mm.Down_CNN()
    ft.pretext.load_state_dict(sd)
    ft.train()
    for br in (ft.pretext.anchor, ft.pretext.postve):
        br.pretrained.dp.p = 0.0
    out1 = ft(x1[:, 0:1], x1[:, 1:2])
    loss1 = torch.nn.BCELoss()(out1, msk)
    loss1.backward()
    out["ft_in"] = x1.numpy()
    out["ft_mask"] = msk.numpy()
    out["ft_out"] = out1.detach().numpy()
    out["ft_loss"] = np.array(float(loss1))
    ftp = dict(ft.pretext.named_parameters())
    out["ft_grad_l2"] = np.array([float(ftp[k].grad.double().norm()) for k in keys])
    out["ft_grad_samples"] = np.concatenate([ftp[k].grad.reshape(-1)[samp_idx[k]].numpy() for k in keys])

    # ---- CLMR baseline step (pretext.py:494-511): one DS_CNN on both views, shared weights ---------------
    ds = mm.DS_CNN()
    ds.load_state_dict({k[len("anchor."):]: v for k, v in sd.items() if k.startswith("anchor.")})
    ds.pretrained.dp.p = 0.0
    opt_c = torch.optim.Adam(params=ds.parameters(), lr=0.00001)
    _, cl, cp_, cn_ = pt.train_epoch(ds, loader, lf.NTXent(batch_len=B, temperature=0.25), opt_c, pt_task="clmr")
    out["clmr_loss_cos"] = np.array([cl, cp_, cn_])
    dsp = dict(ds.named_parameters())
    akeys = [k[len("anchor."):] for k in keys if k.startswith("anchor.")]
    out["clmr_grad_l2"] = np.array([float(dsp[k].grad.double().norm()) for k in akeys])

    # ---- NT-Xent on its own, incl. a short last batch (loss_functions.py:30) --------------------
    e1 = torch.rand(16, 313, generator=g)
    e2 = torch.rand(16, 313, generator=g)
    out["nt_a"] = e1.numpy()
    out["nt_p"] = e2.numpy()
    l, cp, cn = lf.NTXent(16, 0.25)(e1, e2)
    out["nt_full"] = np.array([float(l), cp, cn])
    l, cp, cn = lf.NTXent(16, 0.25)(e1[:5], e2[:5])
    out["nt_short"] = np.array([float(l), cp, cn])
    l, cp, cn = lf.NTXent(16, 0.5)(e1 * 0 + 0.3, e2 * 0 + 0.7)
    out["nt_const"] = np.array([float(l), cp, cn])

    path = os.path.join(ROOT, "tests", "golden", "encoder_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k in ("val_loss_cos", "train_loss_cos", "nt_full", "nt_short", "nt_const"):
        print(k, out[k])
    print("grad_l2", out["grad_l2"][:6], "delta_l2", out["delta_l2"][:4])


def _with_sd(model, sd):
    model.pretext.load_state_dict(sd)
    model.eval()
    return model


if __name__ == "__main__":
    main()
