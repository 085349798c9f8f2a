"""Host-side logic of the drop-in boundary that needs no GPU: state_dict layout, checkpoint containers,
arena repacking, error behaviour on CPU tensors, optimizer state layout."""
import os

import numpy as np
import pytest
import torch

from oracle import vit_oracle as O
from touhouimageclassification_b200.model import ViT, ViTConfig, ViTForImageClassification
from touhouimageclassification_b200.serve import extract_state_dict, get_model

TINY = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, image_size=32, num_labels=10)


def test_state_dict_layout_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "vitb16_forward.npz"))
    m = ViTForImageClassification(ViTConfig())
    sd = m.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["keys"]]
    assert [str(tuple(v.shape)) for v in sd.values()] == [str(s) for s in g["shapes"]]
    assert all(v.dtype == torch.float32 for v in sd.values())
    assert len(list(m.buffers())) == 0
    assert [n for n, _ in m.named_parameters()] == list(sd.keys())
    assert len(list(m.base_model.parameters())) == 198  # everything except classifier.{weight,bias}


def test_strict_load_and_roundtrip():
    m = ViTForImageClassification(ViTConfig(**TINY))
    sd = O.deterministic_state_dict(TINY, 0.05)
    m.load_state_dict(sd, strict=True)
    assert m._arena_ok()
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k])
    bad = dict(sd)
    bad.pop("classifier.bias")
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad, strict=True)
    # q/k/v weights sit next to each other in the arena: [3D, D] view equals the concatenation
    D = 128
    o = m._offsets[4]
    qkv = m._arena[o:o + 3 * D * D].view(3 * D, D)
    p = "vit.encoder.layer.0.attention.attention."
    assert torch.equal(qkv, torch.cat([sd[p + "query.weight"], sd[p + "key.weight"], sd[p + "value.weight"]]))


def test_checkpoint_containers():
    sd = O.deterministic_state_dict(TINY, 0.05)
    assert extract_state_dict(sd) is sd
    assert extract_state_dict((sd, {"opt": 1})) is sd
    assert extract_state_dict({"model": sd}) is sd
    lightning = {"state_dict": {"vit." + k: v for k, v in sd.items()}, "epoch": 3}
    out = extract_state_dict(lightning)
    assert list(out.keys()) == list(sd.keys())
    m = ViTForImageClassification(ViTConfig(**TINY))
    m.load_state_dict(out, strict=True)


def test_factory_and_presets():
    m = get_model("vit_base", 120)
    assert m.config.hidden_size == 768 and m.config.num_hidden_layers == 12 and m.config.image_size == 224
    with pytest.raises(ValueError):
        get_model("nvit", 120)  # serve.py:235 offers it, get_model has no branch
    with pytest.raises(ValueError):
        ViT(120, pretrained=False, model_name="google/unknown")


def test_cpu_inputs_raise_instead_of_falling_back():
    m = ViTForImageClassification(ViTConfig(**TINY))
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(torch.zeros(1, 3, 32, 32))
    with pytest.raises(ValueError, match="channel dimension"):
        m(torch.zeros(1, 1, 32, 32))
    with pytest.raises(ValueError, match="doesn't match model"):
        m(torch.zeros(1, 3, 48, 48))


def test_requires_grad_false_on_backbone_like_ntrain():
    m = ViTForImageClassification(ViTConfig(**TINY))
    for p in m.base_model.parameters():
        p.requires_grad = False
    assert [n for n, p in m.named_parameters() if p.requires_grad] == ["classifier.weight", "classifier.bias"]


def test_lmodule_surface():
    from touhouimageclassification_b200.ntrain import ViTLModule, cutmix_or_mixup
    from oracle import vit_oracle as O
    lm = ViTLModule(120, False, "google/vit-base-patch16-224", lr=1e-5, weight_decay=0.01, full_finetune=False)
    keys = list(lm.state_dict().keys())
    assert keys[0] == "vit.vit.embeddings.cls_token" and keys[-1] == "vit.classifier.bias"  # Lightning ckpt prefix
    opt = lm.configure_optimizers()
    assert isinstance(opt, torch.optim.AdamW) and len(opt.param_groups) == 1
    torch.manual_seed(0)
    x = torch.randn(4, 3, 8, 8)
    y = torch.tensor([0, 1, 2, 3])
    x2, y2 = O.cutmix_or_mixup(x, y, 5)
    assert x2.shape == x.shape and y2.shape == (4, 5)
    torch.testing.assert_close(y2.sum(1), torch.ones(4))
    with pytest.raises(RuntimeError):   # the product path is CUDA only
        cutmix_or_mixup(x, y, 5)


def test_mixup_cutmix_match_torchvision():
    v2 = pytest.importorskip("torchvision.transforms.v2")
    from touhouimageclassification_b200.ntrain import draw_mix
    from oracle import vit_oracle as O
    x = torch.randn(6, 3, 16, 16)
    y = torch.tensor([0, 1, 2, 3, 4, 0])
    ref = v2.RandomChoice([v2.CutMix(num_classes=5), v2.MixUp(num_classes=5)])
    for seed in range(8):
        torch.manual_seed(seed)
        xr, yr = ref(x, y)
        torch.manual_seed(seed)
        xo, yo = O.cutmix_or_mixup(x, y, 5)          # the oracle's restatement is torchvision's, bit for bit
        assert torch.equal(xo, xr) and torch.equal(yo, yr)
        # the product's host-side sampler draws the same numbers in the same order and derives the same box
        torch.manual_seed(seed)
        kind, lam, r_x, r_y = O.draw_mix(16, 16)
        torch.manual_seed(seed)
        mode, lam2, box, lam_label = draw_mix(16, 16)
        assert lam2 == lam and mode == (2 if kind == "cutmix" else 1)
        if kind == "cutmix":
            assert (*box, lam_label) == O.cutmix_box(16, 16, lam, r_x, r_y)
        else:
            assert lam_label == lam
