"""Host-side mirror of the reference's plain-PyTorch training step (``TIC/ViT/finetune.py:54-77`` [a13, a14]).

Same names, argument order and return values as the reference, so its ``train_model`` loop can call these
unchanged. Differences that follow from the B200 engine: bf16 tensor-core arithmetic instead of fp16 autocast
(so the ``GradScaler`` argument is accepted and ignored -- bf16 needs no loss scaling), and when the optimizer
is :class:`FusedAdamW` and the criterion is a plain ``CrossEntropyLoss`` the whole step (forward, fused
softmax-CE, backward, AdamW) runs inside the native engine without autograd.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .model import ViTForImageClassification
from .optim import FusedAdamW


def _plain_ce(criterion) -> bool:
    return (criterion is None or (isinstance(criterion, nn.CrossEntropyLoss) and criterion.weight is None
                                  and criterion.label_smoothing == 0.0 and criterion.reduction == "mean"
                                  and criterion.ignore_index == -100))


def fused_train_step(model: ViTForImageClassification, optimizer: FusedAdamW, inputs=None, target=None, *,
                     patches=None, grad_sync=None, world_size: int = 1) -> torch.Tensor:
    """forward -> softmax-CE (hard or soft targets) -> backward -> AdamW, all on the engine. Returns loss[1] (device).

    ``grad_sync(model, stage_begin, stage_end)`` -- optional hook used by the data-parallel wrapper to launch the
    bucketed gradient all-reduce as backward stages complete.
    """
    batch = inputs.shape[0] if inputs is not None else patches.shape[0] // ((model.config.image_size // 16) ** 2)
    logits = model.engine_forward(inputs, patches=patches, training=True)
    loss, dlogits, _ = ops.softmax_xent(logits, target, grad_scale=1.0 / (batch * world_size), round_grad=True)
    if not optimizer.arena_clean:
        model.grad_arena().zero_()
    optimizer.arena_clean = False
    params = model._params_in_order()
    head_only = not any(p.requires_grad for p in params[:-2])
    if grad_sync is None:
        model.engine_backward(dlogits, batch, head_only=head_only)
    else:
        grad_sync(model, dlogits, batch, head_only)
    optimizer.grads_in_arena = True
    optimizer.step()
    return loss


def train_step(model, data, optimizer, criterion, scaler=None, scheduler=None):
    """``finetune.train_step`` (finetune.py:54-67): returns ``loss.item()`` (a host sync, as in the reference)."""
    model.train()
    optimizer.zero_grad()
    inputs, labels = map(lambda x: x.to("cuda", non_blocking=True), data)
    if isinstance(model, ViTForImageClassification) and isinstance(optimizer, FusedAdamW) and _plain_ce(criterion):
        loss = fused_train_step(model, optimizer, inputs, labels)
    else:
        outputs = model(inputs)
        loss = criterion(outputs.logits.float(), labels)
        loss.backward()
        optimizer.step()
    if scheduler:
        scheduler.step()
    return loss.item()


def validate_step(model, data, criterion):
    """``finetune.validate_step`` (finetune.py:69-77): returns ``(loss.item(), correct)``."""
    model.eval()
    with torch.no_grad():
        inputs, labels = map(lambda x: x.to("cuda", non_blocking=True), data)
        if isinstance(model, ViTForImageClassification) and _plain_ce(criterion):
            logits = model.engine_forward(inputs, training=False)
            loss, _, correct = ops.softmax_xent(logits, labels, need_grad=False)
            return loss.item(), int(correct.item())
        logits = model(inputs).logits
        loss = criterion(logits.float(), labels)
        correct = (logits.argmax(dim=1) == labels).sum().item()
    return loss.item(), correct
