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
        loss = ops.cross_entropy(logits, y)  # F.cross_entropy (ntrain.py:48) on the fused softmax-CE kernel
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
        loss, _, correct = ops.softmax_xent(logits, y, need_grad=False)  # loss + arg-max hits in one launch
        self.log('val_loss', loss.view(()), prog_bar=True)
        acc = (correct.float() / y.shape[0]).view(())
        self.log('val_acc', acc, prog_bar=True)

    def test_step(self, batch, batch_idx):
        x, y = batch
        logits = self.vit(x).logits
        _, _, correct = ops.softmax_xent(logits, y, need_grad=False)
        acc = (correct.float() / y.shape[0]).view(())
        self.log('test_acc', acc, prog_bar=True)

    if _Base is nn.Module:
        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, **ctor):
            """Lightning's classmethod, as the reference calls it (ntrain.py:190): every constructor argument is passed
            again because the reference never calls ``save_hyperparameters()``."""
            ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
            module = cls(**ctor)
            module.load_state_dict(ckpt["state_dict"], strict=True)
            return module


# ----------------------------------------------------------------------------------------------------
# The loop ``L.Trainer(...).fit / .test`` runs for the reference (ntrain.py:219-248) [section 8f rank 1]
# ----------------------------------------------------------------------------------------------------
class FitState:
    """What ``fit`` leaves behind (and what a checkpoint's ``callbacks`` entry restores)."""

    def __init__(self):
        self.epoch = -1              # last finished epoch (0-based)
        self.global_step = 0
        self.best = []               # [(val_acc, path)] of ModelCheckpoint(monitor='val_acc', mode='max', save_top_k)
        self.periodic = []           # paths of ModelCheckpoint(monitor='epoch', every_n_epochs, save_top_k)
        self.best_score = None       # EarlyStopping(monitor='val_acc', mode='max')
        self.wait_count = 0
        self.stopped_early = False
        self.history = []            # (epoch, train_loss, val_loss, val_acc)

    def callbacks_state(self):
        return dict(best=list(self.best), periodic=list(self.periodic), best_score=self.best_score,
                    wait_count=self.wait_count)


def _to_device(batch, device):
    return tuple(t.to(device, non_blocking=True) if torch.is_tensor(t) else t for t in batch)


def _module_device(module):
    return next(module.parameters()).device


def _all_reduce_sums(values):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor(values, dtype=torch.float64, device="cuda" if dist.get_backend() == "nccl" else "cpu")
        dist.all_reduce(t)
        return [float(v) for v in t]
    return values


@torch.no_grad()
def evaluate(lmodule, loader, step: str = "validation_step", device=None, augment=None):
    """One pass of ``validation_step`` / ``test_step`` over ``loader``: the logged metrics averaged the way Lightning
    reduces ``self.log(..., on_epoch=True)`` values -- a mean over batches weighted by batch size. uint8 thumbnail
    batches go through ``augment.tensor`` first (the transform the reference's Dataset applies per sample)."""
    if not hasattr(lmodule, "logged"):
        raise TypeError("evaluate() reads the values of self.log(...) from `lmodule.logged` (the Lightning-free ViTLModule); "
                        "with Lightning installed use its own Trainer")
    device = device or _module_device(lmodule)
    was_training = lmodule.training
    lmodule.eval()
    sums, count = {}, 0
    for i, batch in enumerate(loader):
        batch = _to_device(batch, device)
        n = len(batch[1])
        if batch[0].dtype == torch.uint8:
            if augment is None:
                raise ValueError("a uint8 batch needs augment=GpuAugment(...) (the reference's Dataset transform)")
            batch = (augment.tensor(batch[0]),) + tuple(batch[1:])
        lmodule.logged.clear()
        getattr(lmodule, step)(batch, i)
        for k, v in lmodule.logged.items():
            sums[k] = sums.get(k, 0.0) + float(v) * n
        count += n
    lmodule.train(was_training)
    keys = sorted(sums)
    red = _all_reduce_sums([sums[k] for k in keys] + [float(count)])
    total = red[-1]
    return {k: (red[j] / total if total else 0.0) for j, k in enumerate(keys)}


LIGHTNING_LAYOUT_VERSION = "2.1.0"   # the checkpoint layout written below is Lightning 2.x's


def save_checkpoint(path, lmodule, optimizer, state: FitState):
    """Lightning's checkpoint layout, reduced to what the reference's tooling reads: ``state_dict`` with ``vit.`` keys
    (``load_from_checkpoint``, ``--transform``), ``optimizer_states``, ``epoch`` / ``global_step`` for resuming."""
    import os
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    # "pytorch-lightning_version" must parse as a version: Lightning's migrate_checkpoint (load_from_checkpoint,
    # Trainer(ckpt_path)) runs packaging.version.Version over it. Lightning keys "callbacks" by callback state_key, so the
    # state of this loop's policies lives under its own key and "callbacks" stays empty (Lightning then starts its
    # callbacks fresh, which is what it does for any checkpoint written without them).
    torch.save({"epoch": state.epoch, "global_step": state.global_step, "pytorch-lightning_version": LIGHTNING_LAYOUT_VERSION,
                "state_dict": lmodule.state_dict(), "optimizer_states": [optimizer.state_dict()], "lr_schedulers": [],
                "loops": {}, "callbacks": {}, "tic_callbacks": state.callbacks_state()}, path)


def transform_checkpoint(checkpoint_path, out_path):
    """``ntrain.py --restore CKPT --transform OUT`` (ntrain.py:186-192): the bare inner ``vit`` state_dict (HF keys), the
    ``nViT_epoch*.pth`` files ``serve.load_model`` reads."""
    ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    sd = ckpt["state_dict"]
    inner = {k[len("vit."):]: v for k, v in sd.items() if k.startswith("vit.")}
    if not inner:
        raise ValueError(f"{checkpoint_path} holds no 'vit.' parameters")
    torch.save(inner, out_path)
    return inner


def fit(lmodule, train_loader, val_loader, *, max_epochs: int, patience: int = 3, checkpoint_dir=None,
        train_id: str = "run", save_top_k: int = 3, every_n_epochs: int = 3, ckpt_path=None, optimizer=None, augment=None,
        val_augment=None, data_parallel=None, device=None, logger=None) -> FitState:
    """``L.Trainer(max_epochs, callbacks=[ModelCheckpoint(val_acc, max, top 3), ModelCheckpoint(epoch, every 3, top 3),
    EarlyStopping(val_acc, max, patience)], precision='bf16-mixed').fit(lmodule, datamodule, ckpt_path)``
    (ntrain.py:219-245) without Lightning.

    Per epoch: every training batch goes through ``fused_training_step`` (engine forward, loss, backward, AdamW; the
    train transform and CutMix / MixUp run on the device when ``augment`` is given and the loader yields uint8
    thumbnails) -- or, with a stock optimizer, ``training_step`` + ``backward`` + ``step`` -- then one validation pass,
    the two checkpoint policies and the early-stopping rule, all on the epoch's ``val_acc``. ``data_parallel`` (a
    ``DataParallelTrainer`` around ``lmodule.vit``) adds the gradient exchange; metrics are reduced over the ranks and
    rank 0 writes the files. ``ckpt_path`` resumes (weights, optimizer state, epoch, callback state)."""
    import logging
    import os
    import torch.distributed as dist
    log = logger or logging.getLogger("tic_b200.ntrain")
    device = device or _module_device(lmodule)
    optimizer = optimizer or lmodule.configure_optimizers()
    rank0 = not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
    fused = isinstance(optimizer, FusedAdamW)
    if data_parallel is not None and not fused:
        raise ValueError("data_parallel needs the FusedAdamW optimizer (the gradient exchange runs inside the fused step); "
                         "a stock optimizer would train every rank on its own shard without any exchange")
    state = FitState()
    if ckpt_path is not None:
        ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
        lmodule.load_state_dict(ckpt["state_dict"], strict=True)
        if ckpt.get("optimizer_states"):
            optimizer.load_state_dict(ckpt["optimizer_states"][0])
        state.epoch = int(ckpt.get("epoch", -1))
        state.global_step = int(ckpt.get("global_step", 0))
        cb = ckpt.get("tic_callbacks") or {}   # absent in a checkpoint written by Lightning itself: policies start fresh
        state.best = [tuple(b) for b in cb.get("best", [])]
        state.periodic = list(cb.get("periodic", []))
        state.best_score = cb.get("best_score")
        state.wait_count = int(cb.get("wait_count", 0))
        log.info(f"Restored {ckpt_path}: resuming after epoch {state.epoch}")

    def ckpt_name(epoch, val_acc):
        return os.path.join(checkpoint_dir, f"checkpoint_{train_id}_epoch={epoch:02d}_val_acc={val_acc:.4f}.ckpt")

    if data_parallel is not None and data_parallel.world_size > 1:
        # the loss gradient is pre-scaled by 1 / (local batch * world) and every bucket is all-reduced once per step: the
        # ranks must see the same number of equally sized batches (shard with drop_last / equal shard lengths)
        shape = [len(train_loader), len(getattr(train_loader, "dataset", ())), int(getattr(train_loader, "batch_size", 0) or 0)]
        shapes = [None] * data_parallel.world_size
        dist.all_gather_object(shapes, shape)
        if any(s != shapes[0] for s in shapes):
            raise ValueError(f"data-parallel fit needs identical shard shapes on every rank (batches, samples, batch size): {shapes}")
        if augment is not None:   # (seed, global sample index): rank r's i-th local sample of a step is sample r * B + i
            augment.shard(dist.get_rank(), data_parallel.world_size)

    for epoch in range(state.epoch + 1, max_epochs):
        lmodule.train()
        running, steps = None, 0
        for i, batch in enumerate(train_loader):
            batch = _to_device(batch, device)
            if fused:
                kw = {}
                if data_parallel is not None:
                    kw = dict(grad_sync=data_parallel._grad_sync, world_size=data_parallel.world_size)
                loss = lmodule.fused_training_step(batch, optimizer, augment=augment, **kw)
            else:
                optimizer.zero_grad()
                loss = lmodule.training_step(batch, i)
                loss.backward()
                optimizer.step()
            loss = loss.detach().float().reshape(())
            running = loss if running is None else running + loss   # stays on the device: no sync per step
            steps += 1
            state.global_step += 1
        train_loss = float(running) / steps if steps else 0.0
        metrics = {}
        if val_loader is not None:  # the reference's validation split carries the training transform (ntrain.py:141-144)
            metrics = evaluate(lmodule, val_loader, "validation_step", device, val_augment or augment)
        val_acc, val_loss = metrics.get("val_acc", 0.0), metrics.get("val_loss", 0.0)
        state.epoch = epoch
        state.history.append((epoch, train_loss, val_loss, val_acc))
        log.info(f"epoch {epoch}: train_loss {train_loss:.4f} val_loss {val_loss:.4f} val_acc {val_acc:.4f}")

        # EarlyStopping(monitor='val_acc', mode='max', patience): counted before the files are written so that a
        # checkpoint restores the same decision state
        if state.best_score is None or val_acc > state.best_score:
            state.best_score, state.wait_count = val_acc, 0
        else:
            state.wait_count += 1
        stop = patience > 0 and state.wait_count >= patience

        if checkpoint_dir is not None and rank0:
            path = ckpt_name(epoch, val_acc)
            # ModelCheckpoint(monitor='val_acc', mode='max', save_top_k)
            keep = len(state.best) < save_top_k or val_acc > min(s for s, _ in state.best)
            # ModelCheckpoint(monitor='epoch', mode='max', every_n_epochs, save_top_k): the newest periodic files
            periodic = every_n_epochs > 0 and (epoch + 1) % every_n_epochs == 0
            if keep:
                state.best.append((val_acc, path))
                state.best.sort(key=lambda sp: -sp[0])
            if periodic:
                state.periodic.append(path)
            dropped = []
            if len(state.best) > save_top_k:
                dropped.append(state.best.pop()[1])
            if len(state.periodic) > save_top_k:
                dropped.append(state.periodic.pop(0))
            if keep or periodic:
                save_checkpoint(path, lmodule, optimizer, state)
            alive = {p for _, p in state.best} | set(state.periodic)
            for p in dropped:
                if p not in alive and os.path.exists(p):
                    os.remove(p)
        if stop:
            state.stopped_early = True
            log.info(f"val_acc has not improved for {patience} epochs: stopping after epoch {epoch}")
            break
    return state


def test(lmodule, test_loader, device=None, augment=None):
    """``trainer.test(lmodule, datamodule)`` (ntrain.py:248): the ``test_acc`` of ``test_step`` over the loader."""
    return evaluate(lmodule, test_loader, "test_step", device, augment)


# the reference's TIC/utils/parameter.py:1-8
NUM_CLASSES = 120
VIT_IMAGE_SIZE = 224
CHECKPOINT_DIR = "checkpoint"
TEST_DIR = "data/testset"


def train_main(PRETRAINED: bool, MODEL_NAME: str, LR: float, WEIGHT_DECAY: float, FULL_FINETUNE: bool, BATCH_SIZE: int,
               NUM_WORKERS: int, TRAIN_SPLIT: float, DATA_DIR: str, MAX_EPOCHS: int, ENABLE_MIX_UP: bool,
               ENABLE_AUGMENTATION: bool, TRAIN_ID: str, PATIENCE: int = 3, ONLY_GREY_AUGMENTATION: bool = False,
               ENABLE_DIVERSITY: bool = True, ENABLE_GENERALIZATION: bool = True, argv=None, num_classes: int = NUM_CLASSES,
               test_dir: str = TEST_DIR, checkpoint_dir: str = CHECKPOINT_DIR):
    """``ntrain.train_main`` (ntrain.py:160-248) with the same arguments and command line (``--restore``, ``--test``,
    ``--transform``): module + data module + the fit / test loop above, on the engine, training from uint8 thumbnails
    with the transform on the device. Returns what it produced (the bare state_dict for ``--transform``, else the
    ``FitState`` and the test metrics) instead of exiting."""
    import argparse
    import os
    from .data import AugmentedDataset
    parser = argparse.ArgumentParser()
    parser.add_argument("--restore", "-r", type=str, default=None, help="Path to the checkpoint to restore")
    parser.add_argument("--test", "-t", action="store_true", help="Only test model without training")
    parser.add_argument("--transform", "-tr", type=str, default=None, help="Transform the checkpoint")
    args = parser.parse_args(argv)
    torch.manual_seed(42)  # L.seed_everything(42)
    if args.transform:
        if not args.restore:
            raise SystemExit("No checkpoint to transform")
        return transform_checkpoint(args.restore, args.transform)
    lmodel = ViTLModule(num_classes=num_classes, pretrained=PRETRAINED, model_name=MODEL_NAME, lr=LR,
                        weight_decay=WEIGHT_DECAY, enable_mixup=ENABLE_MIX_UP, full_finetune=FULL_FINETUNE,
                        fused_optimizer=True).cuda()
    data = AugmentedDataset(train_path=DATA_DIR, test_path=test_dir, batch_size=BATCH_SIZE, train_split=TRAIN_SPLIT,
                            num_workers=NUM_WORKERS, image_size=VIT_IMAGE_SIZE, enable_augmentation=ENABLE_AUGMENTATION,
                            enable_diversity=ENABLE_DIVERSITY, enable_generalization=ENABLE_GENERALIZATION,
                            only_grey_augmentation=ONLY_GREY_AUGMENTATION)
    state = None
    if not args.test:
        data.setup("fit")
        state = fit(lmodel, data.train_dataloader(), data.val_dataloader(), max_epochs=MAX_EPOCHS, patience=PATIENCE,
                    checkpoint_dir=os.path.join(checkpoint_dir, TRAIN_ID), train_id=TRAIN_ID, ckpt_path=args.restore,
                    augment=data.augment(seed=42))
    elif args.restore:
        ckpt = torch.load(args.restore, map_location="cpu", weights_only=False)
        lmodel.load_state_dict(ckpt["state_dict"], strict=True)
    data.setup("test")
    return state, test(lmodel, data.test_dataloader(), augment=data.test_augment())
