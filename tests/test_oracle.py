"""The oracle (oracle/vit_oracle.py) against the golden vectors generated from the reference's own model class
(transformers.ViTForImageClassification, tests/golden/make_golden.py) and, when transformers is importable,
against that class live. CPU only."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import vit_oracle as O

TINY = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, image_size=32, num_labels=10)
BASE = dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, image_size=224, num_labels=120)


def test_oracle_forward_matches_golden_tiny(golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny_train_step.npz"))
    sd = O.deterministic_state_dict(TINY, 0.05)
    x = O.deterministic_images(3, 32, seed=1)
    logits = O.vit_forward(sd, x, 2)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=1e-4, atol=2e-5)
    y = torch.tensor([0, 3, 7])
    assert abs(float(O.cross_entropy(logits, y)) - float(g["loss"])) < 1e-5
    soft = O.mixup_targets(y, 10, 0.3)
    assert abs(float(O.cross_entropy(logits, soft)) - float(g["loss_soft"])) < 1e-5
    conf, idx = O.serve_postprocess(logits)
    assert np.array_equal(idx.numpy(), g["idx"])
    np.testing.assert_allclose(conf.numpy(), g["conf"], rtol=1e-4)


def test_oracle_train_step_matches_golden_tiny(golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny_train_step.npz"))
    sd = O.deterministic_state_dict(TINY, 0.05)
    x = O.deterministic_images(3, 32, seed=1)
    y = torch.tensor([0, 3, 7])
    loss, logits, grads, new_sd = O.train_step(sd, {}, x, y, 2, lr=1e-3, weight_decay=0.01)
    names = [str(n) for n in g["names"]]
    assert names == list(sd.keys())
    for i, n in enumerate(names):
        gn = float(grads[n].double().norm())
        if "key.bias" in n:  # mathematically zero (softmax shift invariance)
            assert gn < 1e-6
            continue
        assert abs(gn - g["grad_norms"][i]) <= 2e-4 * g["grad_norms"][i] + 1e-7, n
        k = min(16, grads[n].numel())
        np.testing.assert_allclose(grads[n].flatten()[:k].numpy(), g["grad_heads"][i][:k], rtol=2e-3, atol=1e-6 + 2e-4 * g["grad_norms"][i] / max(1, grads[n].numel()) ** 0.5)
        # AdamW's first step moves every element by ~lr * sign(g): where g is rounding noise the sign is arbitrary,
        # so the norm gets a loose bound and the elementwise check is restricted to well-conditioned gradients.
        np.testing.assert_allclose(float(new_sd[n].double().norm()), g["after_norms"][i], rtol=1e-3)
        stable = np.abs(g["grad_heads"][i][:k]) > 1e-6
        np.testing.assert_allclose(new_sd[n].flatten()[:k].numpy()[stable], g["after_heads"][i][:k][stable], rtol=1e-4, atol=2e-5)


def test_oracle_forward_matches_golden_vitb16(golden_dir):
    g = np.load(os.path.join(golden_dir, "vitb16_forward.npz"))
    sd = O.deterministic_state_dict(BASE, 0.02)
    assert [str(k) for k in g["keys"]] == list(sd.keys())
    assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in g["shapes"]]
    assert len(sd) == 200 and sum(v.numel() for v in sd.values()) == 85_890_936
    x = O.deterministic_images(2, 224, seed=2)
    with torch.no_grad():
        logits = O.vit_forward(sd, x, 12)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=2e-4, atol=2e-5)


def test_oracle_matches_live_transformers():
    transformers = pytest.importorskip("transformers")
    torch.manual_seed(3)
    cfg = dict(TINY, image_size=48, num_labels=7)
    m = transformers.ViTForImageClassification(transformers.ViTConfig(**cfg))
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.05)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.randn(4, 3, 48, 48)
    y = torch.randint(0, 7, (4,))
    ref = m(x).logits
    F.cross_entropy(ref, y).backward()
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = O.vit_forward(leaves, x, 2)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-5)
    O.cross_entropy(out, y).backward()
    for n, p in m.named_parameters():
        if "key.bias" in n:
            continue
        torch.testing.assert_close(leaves[n].grad, p.grad, rtol=2e-3, atol=1e-6)


def test_oracle_adamw_matches_torch():
    torch.manual_seed(0)
    p = torch.randn(1000)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, weight_decay=0.01)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 5):
        g = torch.randn(1000) * 0.1
        pr.grad = g.clone()
        opt.step()
        p, m, v = O.adamw_update(p, g, m, v, step, 1e-3)
    torch.testing.assert_close(p, pr.detach(), rtol=1e-5, atol=1e-6)


def test_bf16_sim_is_close_to_fp32():
    sd = O.deterministic_state_dict(TINY, 0.05)
    x = O.deterministic_images(3, 32, seed=1)
    a = O.vit_forward(sd, x, 2)
    b = O.vit_forward(sd, x, 2, bf16_sim=True)
    rel = float((a - b).norm() / a.norm())
    assert 0 < rel < 2e-2
