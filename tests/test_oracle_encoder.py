"""The encoder / loss / Adam oracle against vectors produced by the reference's own modules
(tools/make_golden.py imported /root/reference/zeroNoteSamba/{models/models.py,
models/loss_functions.py, pretext.py} in the build container)."""
import os

import numpy as np
import pytest
import torch

from oracle import encoder_oracle as eo


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "encoder_golden.npz"))


@pytest.fixture(scope="module")
def sd(gold):
    sd = eo.he_normal_state_dict(int(gold["ckpt_seed"]))
    s1 = float(sum(v.double().sum() for v in sd.values()))
    s2 = float(sum((v.double() ** 2).sum() for v in sd.values()))
    assert np.allclose([s1, s2], gold["ckpt_checksum"], rtol=1e-12)
    return sd


def test_state_dict_layout(gold):
    layout = eo.state_dict_layout()
    assert list(layout.keys()) == [str(k) for k in gold["layout_keys"]]
    assert [int(np.prod(s)) for s in layout.values()] == list(gold["layout_numel"])
    assert len(layout) == 36 and sum(int(np.prod(s)) for s in layout.values()) == 26819202


def test_ntxent_against_reference(gold):
    a, p = torch.from_numpy(gold["nt_a"]), torch.from_numpy(gold["nt_p"])
    l, cp, cn = eo.ntxent(a, p, 16, 0.25)
    assert np.allclose([float(l), cp, cn], gold["nt_full"], rtol=2e-6)
    l, cp, cn = eo.ntxent(a[:5], p[:5], 16, 0.25)
    assert np.allclose([float(l), cp, cn], gold["nt_short"], rtol=2e-6)
    l, cp, cn = eo.ntxent(a * 0 + 0.3, p * 0 + 0.7, 16, 0.5)
    assert np.allclose([float(l), cp, cn], gold["nt_const"], rtol=5e-6)  # cos == 1 up to fp32 rounding
    assert abs(float(l) - np.log(16)) < 1e-5  # all embeddings equal -> ln B
    l2, cp2, cn2 = eo.ntxent_loop_numpy(gold["nt_a"], gold["nt_p"], 16, 0.25)
    assert np.allclose([l2, cp2, cn2], gold["nt_full"], rtol=5e-6)


def test_down_cnn_forward(gold, sd):
    torch.set_num_threads(8)
    x = torch.from_numpy(gold["down_in"])
    with torch.no_grad():
        a = eo.ds_cnn_forward(sd, "anchor", x[:, 0:1]).numpy()
        p = eo.ds_cnn_forward(sd, "postve", x[:, 1:2]).numpy()
        mx = eo.down_forward(sd, x[:, 0:1], x[:, 1:2], "max").numpy()
        mean = eo.down_forward(sd, x[:, 0:1], x[:, 1:2], "mean").numpy()
    assert a.shape == (2, 40)
    for got, key in ((a, "down_anchor"), (p, "down_postve"), (mx, "down_max"), (mean, "down_mean")):
        assert np.allclose(got, gold[key], rtol=1e-5, atol=1e-6), key


def test_training_step(gold, sd):
    torch.set_num_threads(8)
    batch = torch.from_numpy(gold["step_batch"])
    res = eo.pretext_step(sd, batch, batch_len=batch.shape[0], temperature=0.25, lr=1e-6)
    assert np.allclose([res["loss"], res["cos_pos"], res["cos_neg"]], gold["train_loss_cos"], rtol=1e-5)
    assert np.allclose(gold["val_loss_cos"], gold["train_loss_cos"], rtol=1e-6)  # dropout off
    assert np.allclose(res["anc_emb"].numpy(), gold["step_anc_emb"], rtol=1e-5, atol=1e-6)
    assert np.allclose(res["pos_emb"].numpy(), gold["step_pos_emb"], rtol=1e-5, atol=1e-6)
    keys = list(eo.state_dict_layout().keys())
    off = gold["sample_off"]
    for i, k in enumerate(keys):
        g = res["grads"][k]
        assert np.isclose(float(g.double().norm()), gold["grad_l2"][i], rtol=1e-3, atol=1e-9), k
        idx = gold["sample_idx"][off[i]:off[i + 1]]
        gs = g.reshape(-1)[idx].numpy()
        ref = gold["grad_samples"][off[i]:off[i + 1]]
        assert np.allclose(gs, ref, rtol=2e-3, atol=2e-4 * np.abs(ref).max() + 1e-12), k
        delta = (res["new_sd"][k].double() - sd[k].double())
        assert np.isclose(float(delta.norm()), gold["delta_l2"][i], rtol=2e-3), k


def test_finetune_step_batch1(gold, sd):
    """Down_CNN -> BCELoss -> backward at batch size 1 (epochs.py:45-63) against the reference's vectors."""
    torch.set_num_threads(8)
    x = torch.from_numpy(gold["ft_in"])
    msk = torch.from_numpy(gold["ft_mask"])
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = eo.down_forward(params, x[:, 0:1], x[:, 1:2], "max")
    loss = torch.nn.functional.binary_cross_entropy(out, msk)
    loss.backward()
    assert np.allclose(out.detach().numpy(), gold["ft_out"], rtol=1e-5, atol=1e-6)
    assert np.isclose(float(loss), float(gold["ft_loss"]), rtol=1e-5)
    keys = list(eo.state_dict_layout().keys())
    for i, k in enumerate(keys):
        assert np.isclose(float(params[k].grad.double().norm()), gold["ft_grad_l2"][i], rtol=1e-3, atol=1e-9), k


def test_clmr_shared_weights(gold, sd):
    """CLMR baseline (pretext.py:494-511): the oracle's DS_CNN applied to both views with shared weights."""
    torch.set_num_threads(8)
    batch = torch.from_numpy(gold["step_batch"])
    single = {k[len("anchor."):]: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("anchor.")}
    shared = {("anchor." + k): v for k, v in single.items()}
    a = eo.ds_cnn_forward(shared, "anchor", batch[:, 0:1])
    p = eo.ds_cnn_forward(shared, "anchor", batch[:, 1:2])
    loss, cp, cn = eo.ntxent(a, p, batch.shape[0], 0.25)
    loss.backward()
    assert np.allclose([float(loss.detach()), cp, cn], gold["clmr_loss_cos"], rtol=1e-5)
    for i, k in enumerate(single):
        assert np.isclose(float(single[k].grad.double().norm()), gold["clmr_grad_l2"][i], rtol=1e-3, atol=1e-9), k


def test_rms_gate_oracle_properties():
    from oracle import vqt_oracle as vo
    from zeronotesamba_b200 import synth
    d, o = synth.stem_pair(0, 10.0)
    r = vo.rms_frames(o)
    assert r.shape == (313,) and np.all(r >= 0)
    assert vo.rms_fraction(o, o * 0.6) == 1.0 and vo.rms_fraction(o, o * 5.0) == 0.0
    assert 0.0 <= vo.rms_fraction(o, d) <= 1.0
