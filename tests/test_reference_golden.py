"""Fixtures produced by importing the REFERENCE'S OWN modules (tests/golden/make_golden_ref.py ran
TIC.ViT.finetune.early_exit and TIC.utils.preprocess.get_transforms from /root/reference in the build container).
They pin the oracle (CPU) and the product (GPU) for the rows of SURVEY section 8 that the reference itself implements."""
import logging
import os

import numpy as np
import pytest
import torch

from oracle import augment_oracle as A


def test_early_exit_matches_the_reference_function(golden_dir):
    from touhouimageclassification_b200.finetune import early_exit
    g = np.load(os.path.join(golden_dir, "ref_early_exit.npz"))
    log = logging.getLogger("t")
    log.disabled = True
    for tl, k, want in zip(g["timelines"], g["tolerances"], g["verdicts"]):
        tl = [float(v) for v in tl if not np.isnan(v)]
        assert bool(early_exit(tl, int(k), log)) == bool(want), (tl, k)


def _unpatchify(tok, B, S):
    """bf16 patch rows [B*(S/16)^2, 768] (K ordered c, py, px) as uint16 -> float32 images [B, 3, S, S]."""
    G = S // 16
    f = (tok.astype(np.uint32) << 16).view(np.float32).reshape(B, G, G, 3, 16, 16)
    return f.transpose(0, 3, 1, 4, 2, 5).reshape(B, 3, S, S)


def test_oracle_inference_transform_matches_the_reference_get_transforms(golden_dir):
    """preprocess.py:73-77 via PIL (uint8 bilinear + antialias, ToTensor, Normalize with the dataset statistics) against
    the oracle's 'none' recipe: the resample may differ by 1 LSB of uint8 (PIL's fixed-point filter), i.e. 1/255/std."""
    g = np.load(os.path.join(golden_dir, "ref_preprocess.npz"))
    imgs, ref, mean, std = g["images"], g["outputs"], g["mean"], g["std"]
    S = int(g["size"][0])
    keep = A.MEAN.copy(), A.STD.copy()
    try:
        A.MEAN[:] = mean.astype(np.float32)
        A.STD[:] = std.astype(np.float32)
        pix, tok = A.augment_batch(imgs, 0, first_sample=0, size=S, recipe="none")
    finally:
        A.MEAN[:], A.STD[:] = keep
    # uint8 pixels after the resize: at most 1 LSB from PIL on every pixel, identical on most
    ref_u8 = np.rint((ref * std[None, :, None, None] + mean[None, :, None, None]) * 255.0).astype(int).transpose(0, 2, 3, 1)
    d = np.abs(pix.astype(int) - ref_u8)
    assert d.max() <= 1 and (d == 0).mean() > 0.7
    # normalised bf16 patch rows: 1 LSB of the resize (1/255/std) + bf16 rounding of values up to ~3
    out = _unpatchify(tok, imgs.shape[0], S)
    tol = 1.0 / 255.0 / std.min() + 3.0 * 2.0 ** -8
    assert np.abs(out - ref).max() <= tol


@pytest.mark.gpu
def test_gpu_inference_transform_matches_the_reference_get_transforms(golden_dir):
    from touhouimageclassification_b200 import serve as S_
    g = np.load(os.path.join(golden_dir, "ref_preprocess.npz"))
    imgs, ref, mean, std = g["images"], g["outputs"], g["mean"], g["std"]
    S = int(g["size"][0])
    patches = S_.preprocess_u8(torch.from_numpy(imgs).cuda(), mean.tolist(), std.tolist(), size=S)
    tok = patches.view(torch.int16).cpu().numpy().view(np.uint16)
    out = _unpatchify(tok, imgs.shape[0], S)
    tol = 1.0 / 255.0 / std.min() + 3.0 * 2.0 ** -8
    assert np.abs(out - ref).max() <= tol
    assert np.abs(out - ref).mean() < 0.2 * tol


TINY = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, image_size=32, num_labels=10)


def test_oracle_train_steps_match_the_reference_train_step(golden_dir):
    """Three calls of the reference's own finetune.train_step + validate_step (run in fp32 by the generator) against the
    oracle's restatement: losses to 1e-4, parameter norms after the third AdamW step to 1e-5 relative."""
    from oracle import vit_oracle as O
    g = np.load(os.path.join(golden_dir, "ref_finetune_steps.npz"))
    sd = O.deterministic_state_dict(TINY, 0.05)
    x = O.deterministic_images(3, 32, seed=1)
    y = torch.tensor([0, 3, 7])
    state = {}
    for want in g["losses"]:
        loss, _, _, sd = O.train_step(sd, state, x, y, TINY["num_attention_heads"], lr=float(g["lr"]), weight_decay=float(g["weight_decay"]))
        assert abs(float(loss) - float(want)) < 1e-4
    logits = O.vit_forward(sd, x, TINY["num_attention_heads"])
    assert abs(float(O.cross_entropy(logits, y)) - float(g["val_loss"])) < 1e-4
    assert int((logits.argmax(1) == y).sum()) == int(g["correct"])
    for n, want in zip(g["names"], g["param_norms"]):
        if "key.bias" in str(n):  # its gradient is rounding noise (mathematically zero), which Adam turns into +-lr steps
            continue
        assert abs(float(sd[str(n)].double().norm()) - want) <= 1e-5 * max(1.0, want), n


@pytest.mark.gpu
def test_engine_train_steps_match_the_reference_train_step(golden_dir):
    """The same three steps through finetune.train_step + FusedAdamW on the B200 engine (bf16 tensor cores): losses
    within 3e-2 of the reference's fp32 run, validate_step agrees on the correct count."""
    from oracle import vit_oracle as O
    from touhouimageclassification_b200.finetune import train_step, validate_step
    from touhouimageclassification_b200.model import ViTConfig, ViTForImageClassification
    from touhouimageclassification_b200.optim import FusedAdamW
    g = np.load(os.path.join(golden_dir, "ref_finetune_steps.npz"))
    m = ViTForImageClassification(ViTConfig(**TINY))
    m.load_state_dict(O.deterministic_state_dict(TINY, 0.05), strict=True)
    m = m.to("cuda")
    opt = FusedAdamW(m, lr=float(g["lr"]), weight_decay=float(g["weight_decay"]))
    crit = torch.nn.CrossEntropyLoss()
    x = O.deterministic_images(3, 32, seed=1)
    y = torch.tensor([0, 3, 7])
    for want in g["losses"]:
        assert abs(train_step(m, (x, y), opt, crit, None) - float(want)) < 3e-2
    vl, correct = validate_step(m, (x, y), crit)
    assert abs(vl - float(g["val_loss"])) < 3e-2 and correct == int(g["correct"])
    for (n, p), want in zip(m.named_parameters(), g["param_norms"]):
        if "key.bias" in n:  # exactly zero gradient here, +-lr random walk in the reference
            continue
        assert abs(float(p.double().norm()) - want) <= 2e-3 * max(1.0, want), n
