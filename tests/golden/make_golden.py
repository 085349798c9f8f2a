"""Generates tests/golden/*.npz from the reference's own model class (transformers.ViTForImageClassification,
the class TIC/ViT/model.py:45 instantiates) on CPU in fp32, with closed-form weights and inputs
(oracle.vit_oracle.deterministic_*), so the fixtures hold outputs only.

Run in the build container:  python tests/golden/make_golden.py
Recorded versions are stored inside each file (transformers is unpinned by the reference).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import vit_oracle as O  # noqa: E402

import transformers  # noqa: E402
from transformers import ViTConfig, ViTForImageClassification  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TINY = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, image_size=32, num_labels=10)
BASE = dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, image_size=224, num_labels=120)


def hf_model(cfg, scale):
    m = ViTForImageClassification(ViTConfig(**cfg))
    m.load_state_dict(O.deterministic_state_dict(cfg, scale), strict=True)
    return m


def summarize(tensors):
    names = list(tensors)
    norms = np.array([float(tensors[n].double().norm()) for n in names])
    heads = np.stack([np.pad(tensors[n].flatten()[:16].double().numpy(), (0, max(0, 16 - tensors[n].numel()))) for n in names])
    return names, norms, heads


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    meta = dict(transformers=transformers.__version__, torch=torch.__version__)
    # ---- tiny config: forward, hard + soft CE, gradients, one AdamW step
    m = hf_model(TINY, 0.05)
    m.train()
    x = O.deterministic_images(3, 32, seed=1)
    y = torch.tensor([0, 3, 7])
    logits = m(x).logits
    loss = F.cross_entropy(logits, y)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=0.01)
    opt.step()
    after = {n: p.detach().clone() for n, p in m.named_parameters()}
    gn, gnorm, ghead = summarize(grads)
    an, anorm, ahead = summarize(after)
    soft = O.mixup_targets(y, 10, 0.3)
    m2 = hf_model(TINY, 0.05)
    loss_soft = F.cross_entropy(m2(x).logits, soft)
    conf, idx = torch.max(torch.softmax(logits.detach(), 1), 1)
    np.savez(os.path.join(HERE, "tiny_train_step.npz"), names=np.array(gn), logits=logits.detach().numpy(), loss=float(loss),
             loss_soft=float(loss_soft), grad_norms=gnorm, grad_heads=ghead, after_norms=anorm, after_heads=ahead,
             conf=conf.numpy(), idx=idx.numpy(), meta=np.array(str(meta)))
    # ---- ViT-B/16 at full size: logits only
    m = hf_model(BASE, 0.02).eval()
    x = O.deterministic_images(2, 224, seed=2)
    with torch.no_grad():
        logits = m(x).logits
    keys = list(m.state_dict().keys())
    shapes = [tuple(v.shape) for v in m.state_dict().values()]
    np.savez(os.path.join(HERE, "vitb16_forward.npz"), logits=logits.numpy(), keys=np.array(keys),
             shapes=np.array([str(s) for s in shapes]), meta=np.array(str(meta)))
    print("wrote golden fixtures", meta)


if __name__ == "__main__":
    main()
