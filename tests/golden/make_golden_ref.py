"""Generates tests/golden/ref_*.npz by IMPORTING THE REFERENCE'S OWN MODULES from /root/reference (build container only;
the fixtures travel, the reference does not):

  * ref_early_exit.npz     TIC.ViT.finetune.early_exit (finetune.py:79-91) on random validation-loss timelines
  * ref_finetune_steps.npz TIC.ViT.finetune.train_step x3 + validate_step (finetune.py:54-77), the functions themselves,
                           on the class TIC/ViT/model.py:45 instantiates with closed-form weights. The reference moves
                           every batch to "cuda"; this container has no GPU, so the generator maps "cuda" to "cpu" for
                           Tensor.to (autocast('cuda') and GradScaler then are no-ops: the step runs in fp32)
  * ref_preprocess.npz     TIC.utils.preprocess.get_transforms (preprocess.py:60-77): the inference transform
                           Resize -> ToTensor -> Normalize(dataset mean / std read from meta_mean_std.pth), applied
                           to synthetic PIL images

Run:  python tests/golden/make_golden_ref.py
"""
import logging
import os
import sys
import tempfile

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import TIC.ViT.finetune as RF             # noqa: E402
from TIC.ViT.finetune import early_exit  # noqa: E402
from TIC.utils import preprocess as P     # noqa: E402


def finetune_steps():
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import vit_oracle as O
    from transformers import ViTConfig, ViTForImageClassification
    tiny = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, image_size=32, num_labels=10)
    real_to = torch.Tensor.to

    def to(self, *a, **k):  # "cuda" -> "cpu": the only change made to the environment the reference code runs in
        a = tuple("cpu" if (isinstance(v, str) and v.startswith("cuda")) else v for v in a)
        if isinstance(k.get("device"), str) and k["device"].startswith("cuda"):
            k["device"] = "cpu"
        return real_to(self, *a, **k)

    torch.Tensor.to = to
    try:
        m = ViTForImageClassification(ViTConfig(**tiny))
        m.load_state_dict(O.deterministic_state_dict(tiny, 0.05))
        opt = torch.optim.AdamW(m.parameters(), lr=1e-3, weight_decay=0.01)
        crit = torch.nn.CrossEntropyLoss()
        scaler = torch.GradScaler()
        x = O.deterministic_images(3, 32, seed=1)
        y = torch.tensor([0, 3, 7])
        losses = [RF.train_step(m, (x, y), opt, crit, scaler) for _ in range(3)]
        val_loss, correct = RF.validate_step(m, (x, y), crit)
    finally:
        torch.Tensor.to = real_to
    names = [n for n, _ in m.named_parameters()]
    norms = np.array([float(p.detach().double().norm()) for _, p in m.named_parameters()])
    np.savez_compressed(os.path.join(HERE, "ref_finetune_steps.npz"), losses=np.array(losses), val_loss=val_loss,
                        correct=correct, names=np.array(names), param_norms=norms, lr=1e-3, weight_decay=0.01)
    print("finetune.train_step x3:", losses, "validate_step:", val_loss, correct)


def main():
    rng = np.random.default_rng(7)
    log = logging.getLogger("golden")
    log.disabled = True
    timelines, tolerances, verdicts = [], [], []
    for _ in range(200):
        n = int(rng.integers(1, 12))
        tl = np.round(rng.uniform(0.1, 2.0, size=n), 3)
        if rng.random() < 0.4:  # plateaus and exact ties exercise the >= in the rule
            tl[int(rng.integers(0, n)):] = tl[int(rng.integers(0, n))]
        k = int(rng.integers(1, 6))
        timelines.append(np.pad(tl, (0, 12 - n), constant_values=np.nan))
        tolerances.append(k)
        verdicts.append(bool(early_exit(list(map(float, tl)), k, log)))
    np.savez_compressed(os.path.join(HERE, "ref_early_exit.npz"), timelines=np.stack(timelines),
                        tolerances=np.array(tolerances), verdicts=np.array(verdicts))

    finetune_steps()
    from PIL import Image
    mean = torch.tensor([0.61, 0.55, 0.52], dtype=torch.float64)   # the reference stores float64 tensors (preprocess.py:104-127)
    std = torch.tensor([0.31, 0.30, 0.29], dtype=torch.float64)
    with tempfile.TemporaryDirectory() as d:
        torch.save({"mean": mean, "std": std}, os.path.join(d, P.META_MEAN_STD_FILENAME))
        tf = P.get_transforms(d, (64, 64))
    yy, xx = np.mgrid[0:96, 0:80].astype(np.float32)
    imgs = []
    for s in range(3):  # smooth gradients + texture, so the resample filter matters but +-1 LSB stays +-1 LSB
        base = np.stack([127 + 100 * np.sin(xx / (7 + s) + c) * np.cos(yy / (9 - s) + 2 * c) for c in range(3)], -1)
        imgs.append(np.clip(base + rng.normal(0, 6, base.shape), 0, 255).astype(np.uint8))
    outs = np.stack([tf(Image.fromarray(im)).numpy() for im in imgs])
    np.savez_compressed(os.path.join(HERE, "ref_preprocess.npz"), images=np.stack(imgs), outputs=outs.astype(np.float32),
                        mean=mean.numpy(), std=std.numpy(), size=np.array([64, 64]))
    print("early_exit:", sum(verdicts), "of", len(verdicts), "stop; preprocess outputs", outs.shape)


if __name__ == "__main__":
    main()
