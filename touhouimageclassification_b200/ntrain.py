"""Mirror of the reference's Lightning module (``TIC/ViT/ntrain.py:16-66`` [a15-a17]).

``ViTLModule`` keeps the constructor signature, the ``vit`` attribute (so Lightning checkpoints carry the
``vit.<hf key>`` prefix), ``configure_optimizers``, ``training_step`` / ``validation_step`` /
``test_step`` and the logged metric names. ``lightning`` is not installed in this image; when it is importable
the class derives from ``lightning.LightningModule``, otherwise from ``torch.nn.Module`` with a ``log`` method
that records the last values in ``self.logged``.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .finetune import fused_train_step
from .model import ViT
from .optim import FusedAdamW

try:  # pragma: no cover - lightning is absent from the build image
    import lightning as L
    _Base = L.LightningModule
except Exception:  # ImportError or a broken install
    _Base = nn.Module


def draw_mix(H: int, W: int):
    """RNG draws of ``v2.RandomChoice([CutMix, MixUp])`` (ntrain.py:30-33), same generator, same order as torchvision
    (``_container.py:152``; Beta(1,1) then the box centre, ``_augment.py:249-267,298-337``). Returns the kernel
    arguments: (mode, lam_image, box, lam_label) with mode 1 = MixUp, 2 = CutMix."""
    idx = int(torch.multinomial(torch.tensor([0.5, 0.5]), 1))
    lam = float(torch.distributions.Beta(torch.tensor([1.0]), torch.tensor([1.0])).sample(()))
    if idx == 0:
        r_x = int(torch.randint(W, size=(1,)))
        r_y = int(torch.randint(H, size=(1,)))
        r = 0.5 * (1.0 - lam) ** 0.5
        r_w_half, r_h_half = int(r * W), int(r * H)
        x1, y1 = max(r_x - r_w_half, 0), max(r_y - r_h_half, 0)
        x2, y2 = min(r_x + r_w_half, W), min(r_y + r_h_half, H)
        lam_adj = float(1.0 - (x2 - x1) * (y2 - y1) / (W * H))
        return 2, lam, (x1, y1, x2, y2), lam_adj
    return 1, lam, (0, 0, 0, 0), lam


def cutmix_or_mixup(x, y, num_classes, want_pixels=True, want_patches=False):
    """``self.cutmix_or_mixup(x, y)`` of the reference's ``training_step`` (ntrain.py:45-46) on the device: ONE kernel
    blends / pastes the rolled batch (bit-exact with torchvision's fp32 ops) and can emit the bf16 patch rows the engine
    consumes, a second tiny kernel writes the soft labels. Returns (pixels | None, soft_labels, patches | None)."""
    if not x.is_cuda:
        raise RuntimeError("cutmix_or_mixup runs on CUDA tensors only (there is no CPU fallback)")
    if y.dim() != 1:
        raise ValueError("labels must be a 1-D tensor of class indices (torchvision v2 CutMix / MixUp contract)")
    mode, lam, box, lam_label = draw_mix(*x.shape[-2:])
    return ops.mix_batch(x, y, num_classes, mode, lam, box, lam_label, want_pixels=want_pixels, want_patches=want_patches)


class ViTLModule(_Base):
    def __init__(self, num_classes: int, pretrained: bool, model_name: str, lr: float, weight_decay: float,
                 enable_mixup: bool = True, full_finetune: bool = True, fused_optimizer: bool = False):
        super().__init__()
        self.num_classes = num_classes
        self.vit = ViT(num_classes, pretrained, model_name)
        self.lr = lr
        self.weight_decay = weight_decay
        self.enable_mixup = enable_mixup
        self.fused_optimizer = fused_optimizer
        self.logged = {}
        if not full_finetune:
            for param in self.vit.base_model.parameters():
                param.requires_grad = False

    if _Base is nn.Module:
        def log(self, name, value, **kwargs):
            self.logged[name] = value

    def configure_optimizers(self):
        if self.fused_optimizer:
            return FusedAdamW(self.vit, lr=self.lr, weight_decay=self.weight_decay)
        return torch.optim.AdamW(self.parameters(), lr=self.lr, weight_decay=self.weight_decay)

    def training_step(self, batch, batch_idx):
        x, y = batch
        if self.enable_mixup:
            x, y, _ = cutmix_or_mixup(x, y, self.num_classes)
        logits = self.vit(x).logits
        loss = F.cross_entropy(logits.float(), y)
        self.log('train_loss', loss, prog_bar=True)
        return loss

    def fused_training_step(self, batch, optimizer: FusedAdamW, grad_sync=None, world_size: int = 1, augment=None):
        """training_step + backward + optimizer step in one engine pass (manual-optimization fast path).

        ``batch`` is what the reference's DataLoader yields (fp32 ``[B, 3, S, S]`` images, ntrain.py:43-44) or, with
        ``augment`` (a :class:`~.augment.GpuAugment`), the raw uint8 NHWC thumbnails: the train transform of
        ``AugmentedDataset.setup`` (ntrain.py:104-112) then runs on the device inside the step."""
        x, y = batch
        patches = None
        if x.dtype == torch.uint8:
            if augment is None:
                raise ValueError("a uint8 batch needs augment=GpuAugment(...) (the reference's train transform)")
            if self.enable_mixup:
                x = augment.tensor(x)
            else:  # no per-batch blend: the augmentation kernel writes the bf16 patch rows directly
                patches, x = augment(x), None
        if self.enable_mixup:  # the mixed image is never materialised: the kernel writes the bf16 patch rows directly
            _, y, patches = cutmix_or_mixup(x, y, self.num_classes, want_pixels=False, want_patches=True)
            x = None
        loss = fused_train_step(self.vit, optimizer, x, y, patches=patches, grad_sync=grad_sync, world_size=world_size)
        self.log('train_loss', loss, prog_bar=True)
        return loss

    def validation_step(self, batch, batch_idx):
        x, y = batch
        logits = self.vit(x).logits
        loss = F.cross_entropy(logits.float(), y)
        self.log('val_loss', loss, prog_bar=True)
        pred = logits.argmax(dim=1)
        acc = (pred == y).float().mean()
        self.log('val_acc', acc, prog_bar=True)

    def test_step(self, batch, batch_idx):
        x, y = batch
        logits = self.vit(x).logits
        pred = logits.argmax(dim=1)
        acc = (pred == y).float().mean()
        self.log('test_acc', acc, prog_bar=True)
