"""Host logic of the Lightning-Trainer replacement (ntrain.fit / evaluate / test / transform_checkpoint), SURVEY section 8f
rank 1: the callbacks the reference configures at ntrain.py:219-245 (top-3 by val_acc, every-3-epochs, early stopping),
Lightning's batch-size-weighted epoch means, resume, and the checkpoint formats. CPU only: a scripted stand-in module plays
the LightningModule; the engine-backed module is exercised by the GPU suite."""
import os

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from touhouimageclassification_b200 import ntrain


class Scripted(nn.Module):
    """Same surface as ViTLModule (steps, log names, configure_optimizers); val_acc follows a script, one entry per epoch."""

    def __init__(self, script):
        super().__init__()
        self.lin = nn.Linear(4, 3)
        self.script, self.val_epochs, self.logged = list(script), 0, {}

    def log(self, name, value, **kw):
        self.logged[name] = value

    def configure_optimizers(self):
        return torch.optim.AdamW(self.parameters(), lr=1e-2, weight_decay=0.01)

    def training_step(self, batch, batch_idx):
        x, y = batch
        loss = F.cross_entropy(self.lin(x), y)
        self.log("train_loss", loss)
        return loss

    def validation_step(self, batch, batch_idx):
        x, y = batch
        self.log("val_loss", F.cross_entropy(self.lin(x), y))
        self.log("val_acc", torch.tensor(self.script[min(self.val_epochs, len(self.script) - 1)]))
        self.val_epochs += 1            # the validation loader below has exactly one batch per epoch

    def test_step(self, batch, batch_idx):
        x, y = batch
        self.log("test_acc", (self.lin(x).argmax(1) == y).float().mean())


def loaders():
    g = torch.Generator().manual_seed(0)
    x, y = torch.randn(12, 4, generator=g), torch.randint(0, 3, (12,), generator=g)
    train = [(x[i:i + 4], y[i:i + 4]) for i in range(0, 12, 4)]
    return train, [(x, y)]


def test_checkpoint_policies_and_early_stopping(tmp_path):
    script = [0.50, 0.60, 0.55, 0.70, 0.65, 0.64, 0.63, 0.90]
    m = Scripted(script)
    train, val = loaders()
    st = ntrain.fit(m, train, val, max_epochs=10, patience=3, checkpoint_dir=str(tmp_path), train_id="t")
    # best so far: .5 .6 (.55: wait 1) .7 (.65: 1) (.64: 2) (.63: 3 -> stop): seven epochs, the .90 is never reached
    assert st.stopped_early and st.epoch == 6 and len(st.history) == 7 and st.global_step == 7 * 3
    assert [round(s, 2) for s, _ in st.best] == [0.70, 0.65, 0.64]
    assert [os.path.basename(p) for p in st.periodic] == ["checkpoint_t_epoch=02_val_acc=0.5500.ckpt",
                                                          "checkpoint_t_epoch=05_val_acc=0.6400.ckpt"]
    # top-3 files + the periodic one that fell out of the top 3 (epoch 2); epochs 0, 1 were deleted, 6 never written
    assert sorted(os.listdir(tmp_path)) == ["checkpoint_t_epoch=02_val_acc=0.5500.ckpt", "checkpoint_t_epoch=03_val_acc=0.7000.ckpt",
                                            "checkpoint_t_epoch=04_val_acc=0.6500.ckpt", "checkpoint_t_epoch=05_val_acc=0.6400.ckpt"]
    ck = torch.load(st.best[0][1], weights_only=False)
    assert ck["epoch"] == 3 and ck["global_step"] == 12 and set(ck["state_dict"]) == {"lin.weight", "lin.bias"}
    assert len(ck["optimizer_states"]) == 1 and ck["tic_callbacks"]["best_score"] == pytest.approx(0.70)
    # what Lightning's own loader (migrate_checkpoint) needs from the file: a parseable version and callbacks keyed by state_key
    from packaging.version import Version
    assert Version(ck["pytorch-lightning_version"]) >= Version("2.0.0") and ck["callbacks"] == {} and ck["lr_schedulers"] == []


def test_patience_zero_disables_early_stopping_and_no_dir_writes_nothing(tmp_path):
    m = Scripted([0.9, 0.1, 0.1, 0.1, 0.1])
    train, val = loaders()
    st = ntrain.fit(m, train, val, max_epochs=5, patience=0)
    assert not st.stopped_early and st.epoch == 4 and os.listdir(tmp_path) == []


def test_resume_continues_epochs_steps_optimizer_and_callback_state(tmp_path):
    train, val = loaders()
    script = [0.5, 0.6, 0.4, 0.3, 0.2, 0.1]
    a = Scripted(script)
    whole = ntrain.fit(a, train, val, max_epochs=4, patience=5, checkpoint_dir=str(tmp_path / "a"), train_id="a", every_n_epochs=1)
    after_epoch_1 = os.path.join(tmp_path / "a", "checkpoint_a_epoch=01_val_acc=0.6000.ckpt")
    resumed = Scripted(script)          # different initial weights: everything must come from the checkpoint
    resumed.val_epochs = 2              # the script continues where the interrupted run stopped
    st = ntrain.fit(resumed, train, val, max_epochs=4, patience=5, checkpoint_dir=str(tmp_path / "b"), train_id="b",
                    every_n_epochs=1, ckpt_path=after_epoch_1)
    assert [h[0] for h in st.history] == [2, 3] and st.epoch == 3 and st.global_step == whole.global_step
    assert st.best_score == pytest.approx(0.6) and st.wait_count == 2
    # same data order, same optimizer state -> the resumed run lands on the uninterrupted run's weights
    for (n, p), (_, q) in zip(a.state_dict().items(), resumed.state_dict().items()):
        assert torch.allclose(p, q, atol=1e-7), n


def test_epoch_means_are_weighted_by_batch_size():
    class M(Scripted):
        def validation_step(self, batch, batch_idx):
            self.log("val_acc", torch.tensor(1.0 if batch_idx == 0 else 0.0))
            self.log("val_loss", torch.tensor(2.0))
    m = M([0.0])
    x, y = torch.zeros(10, 4), torch.zeros(10, dtype=torch.long)
    out = ntrain.evaluate(m, [(x[:8], y[:8]), (x[8:], y[8:])])
    assert out["val_acc"] == pytest.approx(0.8) and out["val_loss"] == pytest.approx(2.0)
    assert ntrain.test(m, [(x, y)])["test_acc"] in (0.0, 1.0)
    with pytest.raises(TypeError):
        ntrain.evaluate(nn.Linear(2, 2), [])


def test_transform_and_load_from_checkpoint_roundtrip(tmp_path):
    """--transform writes the bare HF-key state_dict serve.load_model reads; load_from_checkpoint restores the module."""
    lm = ntrain.ViTLModule(7, False, "google/vit-base-patch16-224", lr=1e-5, weight_decay=0.01)
    st = ntrain.FitState()
    st.epoch, st.global_step = 4, 99
    path = str(tmp_path / "c.ckpt")
    ntrain.save_checkpoint(path, lm, torch.optim.AdamW(lm.parameters(), lr=1e-5), st)
    ck = torch.load(path, weights_only=False)
    assert all(k.startswith("vit.") for k in ck["state_dict"]) and ck["epoch"] == 4
    inner = ntrain.transform_checkpoint(path, str(tmp_path / "nViT_epoch5.pth"))
    assert set(inner) == set(lm.vit.state_dict()) and "classifier.weight" in inner
    bare = torch.load(str(tmp_path / "nViT_epoch5.pth"), weights_only=False)
    assert torch.equal(bare["classifier.weight"], lm.vit.classifier.weight)
    again = ntrain.ViTLModule.load_from_checkpoint(path, num_classes=7, pretrained=False, model_name="google/vit-base-patch16-224",
                                                   lr=1e-5, weight_decay=0.01)
    assert torch.equal(again.vit.classifier.weight, lm.vit.classifier.weight)


def _write_folder(root, per_class, size, seed):
    from PIL import Image
    import numpy as np
    rng = np.random.default_rng(seed)
    for c, n in per_class.items():
        os.makedirs(os.path.join(root, c), exist_ok=True)
        for i in range(n):
            Image.fromarray(rng.integers(0, 256, (size[1], size[0], 3), dtype=np.uint8)).save(os.path.join(root, c, f"{i}.png"))


def test_data_module_mirrors_the_reference_surface(tmp_path):
    """AugmentedDataset (ntrain.py:68-158): ImageFolder class order, 80/20 split, loaders of uint8 thumbnails; the flag
    chain picks the recipe the reference's setup() composes."""
    from touhouimageclassification_b200.data import AugmentedDataset, ThumbnailFolder, recipe_from_flags
    _write_folder(str(tmp_path / "train"), {"reimu": 6, "marisa": 4}, (256, 256), 0)
    _write_folder(str(tmp_path / "test"), {"reimu": 2, "marisa": 1}, (300, 200), 1)       # odd size: brought to 256 x 256
    ds = ThumbnailFolder(str(tmp_path / "train"))
    assert ds.classes == ["marisa", "reimu"] and ds.class_to_idx == {"marisa": 0, "reimu": 1} and len(ds) == 10
    x, y = ds[0]
    assert x.dtype == torch.uint8 and x.shape == (256, 256, 3) and y == 0
    dm = AugmentedDataset(str(tmp_path / "train"), str(tmp_path / "test"), batch_size=4, train_split=0.8, num_workers=0)
    dm.setup("fit")
    dm.setup("test")
    assert len(dm.train_dataset) == 8 and len(dm.val_dataset) == 2 and len(dm.test_dataset) == 3
    xb, yb = next(iter(dm.train_dataloader()))
    assert xb.shape == (4, 256, 256, 3) and xb.dtype == torch.uint8 and yb.dtype == torch.int64
    assert next(iter(dm.test_dataloader()))[0].shape == (3, 256, 256, 3)
    dm2 = AugmentedDataset(str(tmp_path / "train"), str(tmp_path / "test"), train_split=0.8, num_workers=0)
    dm2.setup("fit")
    assert dm2.val_dataset.indices == dm.val_dataset.indices                             # the split is reproducible
    assert dm.recipe == "full"
    assert recipe_from_flags(False) == "none" and recipe_from_flags(True, True, True, True) == "grey"
    assert recipe_from_flags(True, True, False) == "diversity" and recipe_from_flags(True, False, True) == "generalization"
    with pytest.raises(Exception, match="Must select diversity or generalization"):
        recipe_from_flags(True, False, False)
    with pytest.raises(FileNotFoundError):
        ThumbnailFolder(str(tmp_path / "train" / "reimu"))
    with pytest.raises(ValueError, match="uint8 batch needs augment"):
        ntrain.evaluate(Scripted([0.0]), [(xb, yb)])


def test_train_main_command_line_transform_path(tmp_path):
    """`ntrain.py --restore CKPT --transform OUT` (ntrain.py:186-192) needs no GPU: it only rewrites the checkpoint."""
    lm = ntrain.ViTLModule(3, False, "google/vit-base-patch16-224", lr=1e-5, weight_decay=0.01)
    st = ntrain.FitState()
    path, out = str(tmp_path / "c.ckpt"), str(tmp_path / "nViT.pth")
    ntrain.save_checkpoint(path, lm, torch.optim.AdamW(lm.parameters(), lr=1e-5), st)
    kw = dict(PRETRAINED=False, MODEL_NAME="google/vit-base-patch16-224", LR=1e-5, WEIGHT_DECAY=0.01, FULL_FINETUNE=True,
              BATCH_SIZE=8, NUM_WORKERS=0, TRAIN_SPLIT=0.8, DATA_DIR="data", MAX_EPOCHS=1, ENABLE_MIX_UP=True,
              ENABLE_AUGMENTATION=True, TRAIN_ID="t")
    inner = ntrain.train_main(**kw, argv=["--restore", path, "--transform", out])
    assert os.path.exists(out) and set(inner) == set(lm.vit.state_dict())
    with pytest.raises(SystemExit, match="No checkpoint to transform"):
        ntrain.train_main(**kw, argv=["--transform", out])


def test_data_parallel_fit_refuses_a_stock_optimizer():
    """The gradient exchange lives in the fused step (FusedAdamW + DataParallelTrainer): with a stock optimizer every rank
    would train on its own shard without any exchange, silently. fit() says so instead."""
    m = Scripted([0.5])
    train, val = loaders()

    class FakeDP:
        world_size = 2

        def _grad_sync(self, *a):
            raise AssertionError("must not be reached")

    with pytest.raises(ValueError, match="FusedAdamW"):
        ntrain.fit(m, train, val, max_epochs=1, data_parallel=FakeDP())
