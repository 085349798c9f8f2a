"""Host-side mirror of the reference's plain-PyTorch training step (``TIC/ViT/finetune.py:54-77`` [a13, a14]).

Same names, argument order and return values as the reference, so its ``train_model`` loop can call these
unchanged. Differences that follow from the B200 engine: bf16 tensor-core arithmetic instead of fp16 autocast
(so the ``GradScaler`` argument is accepted and ignored -- bf16 needs no loss scaling), and when the optimizer
is :class:`FusedAdamW` and the criterion is a plain ``CrossEntropyLoss`` the whole step (forward, fused
softmax-CE, backward, AdamW) runs inside the native engine without autograd.
"""
from __future__ import annotations

import logging
import math
import os
from typing import Optional

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

    ``grad_sync(model, dlogits, batch, head_only)`` -- the data-parallel wrapper's schedule (backward stage by stage,
    bucketed gradient all-reduce, AdamW per bucket); defaults to the same schedule without the exchange.
    """
    batch = inputs.shape[0] if inputs is not None else patches.shape[0] // ((model.config.image_size // 16) ** 2)
    logits = model.engine_forward(inputs, patches=patches, training=True)
    loss, dlogits, _ = ops.softmax_xent(logits, target, grad_scale=1.0 / (batch * world_size), round_grad=True)
    if not optimizer.arena_clean:
        model.grad_arena().zero_()
    optimizer.arena_clean = False
    params = model._params_in_order()
    head_only = not any(p.requires_grad for p in params[:-2])
    optimizer.grads_in_arena = True
    if grad_sync is None:  # single process: same bucketed schedule, no exchange
        grad_sync = _local_schedule(model, optimizer)
    # runs the backward stage by stage; as each bucket of gradients becomes final (after its all-reduce when there is
    # one) AdamW is applied to that slice on a side stream, under the backward of the earlier layers
    grad_sync(model, dlogits, batch, head_only)
    return loss


def _local_schedule(model, optimizer):
    sched = getattr(optimizer, "_local_schedule", None)
    if sched is None or sched.model is not model:
        from .parallel import DataParallelTrainer
        sched = DataParallelTrainer(model, optimizer, local=True)
        optimizer._local_schedule = sched
    return sched._grad_sync


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


def early_exit(timeline, max_tolerant_epoch, logger):
    """``finetune.early_exit`` (finetune.py:79-91): stop when the validation loss has not improved on the value at the
    start of the last ``max_tolerant_epoch + 1``-epoch window."""
    if len(timeline) < max_tolerant_epoch:
        return False
    window = timeline[-(max_tolerant_epoch + 1):]
    if all(loss >= window[0] for loss in window[1:]):
        logger.info(f"Validation loss has not improved for {max_tolerant_epoch} epochs. Stopping training.")
        return True
    return False


def train_model(model: torch.nn.Module, dataset, optimizer: torch.optim.Optimizer, scheduler, criterion: torch.nn.Module,
                batch_size: int, num_epochs: int, max_tolerant_epoch: int, save_path: str,
                logger: Optional[logging.Logger] = None, skip_optimizer_load: bool = False,
                scheduler_per_epoch: bool = True, num_workers: int = 8):
    """``finetune.train_model`` (finetune.py:93-268) [section 8f rank 1]: the epoch loop around train_step / validate_step.

    Same behaviour as the reference: resume from the newest ``save_path.format(epoch=i)`` (tuple checkpoints
    ``(model_sd, optim_sd[, sched_sd])`` or a bare state_dict), 90/10 ``random_split`` under ``torch.manual_seed(0)``,
    NaN losses replaced by the running mean, a tuple checkpoint per epoch, early stop, per-epoch or per-step scheduler.
    Checkpoints are interchangeable with the reference's in both directions (HF key layout; ``FusedAdamW.state_dict()``
    uses torch AdamW's index order). Returns the list of per-epoch validation losses."""
    from torch.utils.data import DataLoader, random_split
    logger = logger or logging.getLogger("tic_b200.finetune")
    latest = 0
    for i in range(num_epochs, 0, -1):
        if os.path.exists(save_path.format(epoch=i)):
            latest = i
            break
    if latest > 0:
        logger.info(f"Resuming from epoch {latest}")
        ckpt = torch.load(save_path.format(epoch=latest), map_location="cuda", weights_only=False)
        if isinstance(ckpt, tuple) and len(ckpt) >= 2:
            model.load_state_dict(ckpt[0])
            sched_state = ckpt[2] if len(ckpt) > 2 else None
            if not skip_optimizer_load:
                optimizer.load_state_dict(ckpt[1])
                if scheduler and sched_state and scheduler_per_epoch:
                    scheduler.load_state_dict(sched_state)
                    logger.info("Loaded scheduler state.")
                elif scheduler and not scheduler_per_epoch:
                    logger.warning("Resuming per-step scheduler state not fully implemented, may restart LR schedule.")
            elif scheduler and scheduler_per_epoch:
                logger.info(f"Skipping optimizer load, manually advancing scheduler to epoch {latest}")
                for _ in range(latest):
                    scheduler.step()
        else:
            model.load_state_dict(ckpt)
            logger.warning("Loaded checkpoint only contains model state_dict. Optimizer and scheduler state not loaded.")
    else:
        logger.info("Starting training from scratch.")
    start_epoch = latest

    val_size = len(dataset) // 10
    torch.manual_seed(0)  # split consistency across runs (finetune.py:150)
    train_set, val_set = random_split(dataset, [len(dataset) - val_size, val_size])
    train_loader = DataLoader(train_set, batch_size=batch_size, shuffle=True, pin_memory=True, num_workers=num_workers)
    val_loader = DataLoader(val_set, batch_size=batch_size, shuffle=False, pin_memory=True, num_workers=num_workers)
    timeline = []

    def run_train(epoch):
        if scheduler and scheduler_per_epoch:
            logger.info(f"LR for epoch {epoch + 1}: {scheduler.get_last_lr()[0]:.6e}")
        model.train()
        running = 0.0
        for i, data in enumerate(train_loader):
            loss = train_step(model, data, optimizer, criterion, None, scheduler if not scheduler_per_epoch else None)
            if math.isnan(loss):
                logger.warning(f"NaN loss detected at training step {i} in epoch {epoch + 1}. Replacing with avg loss.")
                loss = running / (i + 1) if i > 0 else 0.0
            running += loss
        return running / len(train_loader) if len(train_loader) > 0 else 0.0

    def run_val(epoch):
        optimizer.zero_grad(set_to_none=True)
        running, correct, total = 0.0, 0, 0
        for i, data in enumerate(val_loader):
            loss, ok = validate_step(model, data, criterion)
            if math.isnan(loss):
                logger.warning(f"NaN loss detected during validation step {i} in epoch {epoch + 1}. Replacing with avg loss.")
                loss = running / (i + 1) if i > 0 else 0.0
            running += loss
            correct += ok
            total += len(data[1])
        return (running / len(val_loader) if len(val_loader) > 0 else 0.0), (correct / total * 100 if total else 0.0)

    if start_epoch > 0:
        logger.info(f"Validating model from loaded checkpoint (Epoch {start_epoch}) before resuming training...")
        vl, acc = run_val(start_epoch - 1)
        logger.info(f"Epoch [{start_epoch}], Validation Loss: {vl:.4f}, Accuracy: {acc:.2f}%")
    for epoch in range(start_epoch, num_epochs):
        tl = run_train(epoch)
        vl, acc = run_val(epoch)
        timeline.append(vl)
        ckpt = (model.state_dict(), optimizer.state_dict())
        if scheduler and scheduler_per_epoch:
            ckpt += (scheduler.state_dict(),)
        path = save_path.format(epoch=epoch + 1)
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        torch.save(ckpt, path)
        logger.info(f"Checkpoint saved to {path}")
        logger.info(f"Epoch [{epoch + 1}/{num_epochs}], Training Loss: {tl:.4f}, Validation Loss: {vl:.4f}, Accuracy: {acc:.2f}%")
        if early_exit(timeline, max_tolerant_epoch, logger):
            break
        if scheduler and scheduler_per_epoch:
            scheduler.step()
    return timeline
