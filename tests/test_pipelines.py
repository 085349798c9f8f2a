"""Callers either side of the hot path: the finetune epoch loop with resume, the batched full_judge /
filter pipeline and the web daemon adapter. Host logic runs on CPU; everything that forwards the model is a GPU test."""
import csv
import logging
import os
import threading

import numpy as np
import pytest
import torch

from oracle import vit_oracle as O

TINY = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, image_size=32, num_labels=4)


# ---------------------------------------------------------------------------------------------- CPU
def test_early_exit_window_matches_reference_rule():
    from touhouimageclassification_b200.finetune import early_exit
    log = logging.getLogger("t")
    assert not early_exit([1.0], 3, log)                      # shorter than the window
    assert not early_exit([1.0, 0.9, 0.95, 0.97], 3, log)     # improved inside the window
    assert early_exit([1.0, 0.9, 0.95, 0.97, 0.91], 3, log)   # 0.9 then three epochs without beating it
    assert not early_exit([1.0, 0.9, 0.95, 0.97, 0.89], 3, log)


def test_batching_predictor_coalesces_concurrent_requests():
    from touhouimageclassification_b200.serve import BatchingPredictor
    gate = threading.Event()

    def forward(items):
        gate.wait(2.0)                 # hold the first forward so the other requests pile up behind it
        return [x * 2 for x in items]

    bp = BatchingPredictor(forward, max_batch_size=8, max_wait_s=0.05)
    outs = {}

    def call(i, n):
        outs[i] = bp(list(range(i * 100, i * 100 + n)))

    threads = [threading.Thread(target=call, args=(i, 3)) for i in range(5)]
    for t in threads:
        t.start()
    gate.set()
    for t in threads:
        t.join(5.0)
    bp.close()
    for i in range(5):
        assert outs[i] == [2 * x for x in range(i * 100, i * 100 + 3)]   # every caller gets exactly its own slice
    assert sum(bp.batches) == 15 and max(bp.batches) <= 8 and len(bp.batches) < 5   # fewer forwards than requests
    # errors reach every waiting caller
    bad = BatchingPredictor(lambda items: 1 / 0, max_batch_size=4)
    with pytest.raises(ZeroDivisionError):
        bad([1])
    bad.close()


def test_filter_csv_copies_correct_rows(tmp_path):
    from touhouimageclassification_b200.serve import filter_csv
    src = tmp_path / "src" / "reimu"
    src.mkdir(parents=True)
    for n in ("a.png", "b.png"):
        (src / n).write_bytes(b"x")
    f = tmp_path / "judge.csv"
    with open(f, "w") as fh:
        fh.write("filename,predicted_class,confidence,actual_class,correct,path\n")
        fh.write(f"a.png,reimu,0.9000,reimu,True,{src / 'a.png'}\n")
        fh.write(f"b.png,marisa,0.6000,reimu,False,{src / 'b.png'}\n")
    assert filter_csv(str(f), str(tmp_path / "out")) == (2, 1)
    assert os.path.exists(tmp_path / "out" / "reimu" / "a.png") and not os.path.exists(tmp_path / "out" / "reimu" / "b.png")


# ---------------------------------------------------------------------------------------------- GPU
def _model():
    from touhouimageclassification_b200.model import ViTConfig, ViTForImageClassification
    m = ViTForImageClassification(ViTConfig(**TINY))
    m.load_state_dict(O.deterministic_state_dict(TINY, 0.05), strict=True)
    return m.to("cuda")


class _Blobs(torch.utils.data.Dataset):
    """40 separable 32x32 images: class = which quadrant is bright."""

    def __len__(self):
        return 40

    def __getitem__(self, i):
        c = i % 4
        x = torch.full((3, 32, 32), -0.5)
        x[:, (c // 2) * 16:(c // 2) * 16 + 16, (c % 2) * 16:(c % 2) * 16 + 16] = 1.0 + 0.01 * (i // 4)
        return x, c


@pytest.mark.gpu
def test_train_model_checkpoints_resume_and_interchange_with_torch_adamw(tmp_path):
    from touhouimageclassification_b200.finetune import train_model
    from touhouimageclassification_b200.optim import FusedAdamW
    save = str(tmp_path / "ck" / "tiny_epoch{epoch}.pth")
    m = _model()
    opt = FusedAdamW(m, lr=2e-3, weight_decay=0.01)
    tl = train_model(m, _Blobs(), opt, None, torch.nn.CrossEntropyLoss(), batch_size=12, num_epochs=2, max_tolerant_epoch=5,
                     save_path=save, num_workers=0)
    assert len(tl) == 2 and os.path.exists(save.format(epoch=2))
    ck = torch.load(save.format(epoch=2), map_location="cpu", weights_only=False)
    assert isinstance(ck, tuple) and len(ck) == 2                      # (model_sd, optim_sd), finetune.py:249-255
    assert list(ck[0].keys())[0] == "vit.embeddings.cls_token" and len(ck[1]["state"]) == len(ck[0])
    # resume: a fresh model + optimizer continue at epoch 3 from the tuple checkpoint
    m2 = _model()
    opt2 = FusedAdamW(m2, lr=2e-3, weight_decay=0.01)
    tl2 = train_model(m2, _Blobs(), opt2, None, torch.nn.CrossEntropyLoss(), batch_size=12, num_epochs=4, max_tolerant_epoch=5,
                      save_path=save, num_workers=0)
    assert len(tl2) == 2 and opt2._step > opt._step and os.path.exists(save.format(epoch=4))
    assert all(np.isfinite(v) for v in tl + tl2)
    assert opt2._step == 2 * opt._step                                 # the step counter continued from the checkpoint
    # the optimizer state is torch AdamW's format: a stock AdamW over the same parameters loads it
    m3 = _model()
    ref_opt = torch.optim.AdamW(m3.parameters(), lr=2e-3, weight_decay=0.01)
    ref_opt.load_state_dict(torch.load(save.format(epoch=4), map_location="cuda", weights_only=False)[1])
    st = ref_opt.state[next(iter(m3.parameters()))]
    assert float(st["step"]) == opt2._step and st["exp_avg"].abs().sum() > 0


@pytest.mark.gpu
def test_full_judge_csv_and_daemon(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    from touhouimageclassification_b200 import serve as S
    m = _model().eval()
    classes = {"a": 0, "b": 1, "c": 2, "d": 3}
    rng = np.random.default_rng(0)
    root = tmp_path / "data"
    for cname in classes:
        (root / cname).mkdir(parents=True)
        for j in range(3):
            Image.fromarray(rng.integers(0, 255, (40, 48, 3), dtype=np.uint8)).save(root / cname / f"{cname}{j}.png")
    (root / "a" / "notes.txt").write_text("skip me")
    mean, std = [0.5, 0.5, 0.5], [0.25, 0.25, 0.25]

    def transforms(img):  # the reference's Resize((S,S)) -> ToTensor -> Normalize on the host (preprocess.py:73-77)
        arr = np.asarray(img.resize((32, 32), Image.BILINEAR), dtype=np.float32) / 255.0
        return torch.from_numpy(((arr - mean) / std).astype(np.float32)).permute(2, 0, 1)

    out_csv = tmp_path / "judge.csv"
    acc = S.full_judge(m, transforms, classes, image=str(root), device="cuda", output=str(out_csv), batch_size=5)
    rows = list(csv.DictReader(open(out_csv)))
    assert list(rows[0].keys()) == ["filename", "predicted_class", "confidence", "actual_class", "correct", "path"]
    assert len(rows) == 12 and abs(acc - sum(r["correct"] == "True" for r in rows) / 12) < 1e-9
    # batched forward == the reference's one-image-at-a-time serve() on the same tensors
    for r in rows[:4]:
        cls, conf = S.serve(m, transforms(Image.open(r["path"]).convert("RGB")).unsqueeze(0), classes, "cuda")
        assert cls == r["predicted_class"] and abs(conf - float(r["confidence"])) < 2e-3
    # GPU resize path (no CPU transform): same CSV shape, predictions agree with the host transform on most pictures
    out2 = tmp_path / "judge_gpu.csv"
    S.full_judge(m, None, classes, image=str(root), device="cuda", output=str(out2), batch_size=5, mean=mean, std=std)
    rows2 = {r["path"]: r for r in csv.DictReader(open(out2))}
    assert len(rows2) == 12
    # web daemon: single image -> pair, list -> list, concurrent callers share forwards
    d = S.ModelDaemon(m, classes, transforms=transforms, max_batch_size=8)
    imgs = [Image.open(r["path"]) for r in rows]
    single = d.predict(imgs[0])
    assert single[0] == rows[0]["predicted_class"] and abs(single[1] - float(rows[0]["confidence"])) < 2e-3
    res = {}
    ts = [threading.Thread(target=lambda i=i: res.__setitem__(i, d.predict(imgs[3 * i:3 * i + 3]))) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(20.0)
    d.stop()
    flat = [p for i in range(4) for p in res[i]]
    assert [p[0] for p in flat] == [r["predicted_class"] for r in rows]


@pytest.mark.gpu
def test_ntrain_fit_from_uint8_thumbnails_checkpoints_and_reload(tmp_path):
    """The Lightning-Trainer replacement on the engine (section 8f rank 1): uint8 thumbnails -> fused augmentation ->
    CutMix/MixUp -> fused step, validation / test through the module's own steps, Lightning-layout checkpoints that
    load_from_checkpoint and --transform read, resume continuing the step count."""
    from touhouimageclassification_b200 import ntrain
    from touhouimageclassification_b200.augment import GpuAugment
    rng = np.random.default_rng(0)
    thumbs = torch.from_numpy(rng.integers(0, 256, (8, 64, 64, 3), dtype=np.uint8))
    labels = torch.tensor([0, 1, 2, 3, 4, 0, 1, 2])
    train = [(thumbs[:4], labels[:4]), (thumbs[4:], labels[4:])]
    xv = torch.randn(6, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    yv = torch.tensor([0, 1, 2, 3, 4, 0])
    val = [(xv[:4], yv[:4]), (xv[4:], yv[4:])]
    torch.manual_seed(3)
    lm = ntrain.ViTLModule(5, False, "google/vit-base-patch16-224", lr=1e-4, weight_decay=0.01, enable_mixup=True,
                           fused_optimizer=True).cuda()
    opt = lm.configure_optimizers()
    st = ntrain.fit(lm, train, val, max_epochs=2, patience=0, checkpoint_dir=str(tmp_path), train_id="g", optimizer=opt,
                    augment=GpuAugment(seed=5))
    assert st.epoch == 1 and st.global_step == 4 and opt._step == 4 and len(st.best) == 2
    assert all(np.isfinite(h[1]) and np.isfinite(h[2]) and 0.0 <= h[3] <= 1.0 for h in st.history)
    acc = ntrain.test(lm, val)["test_acc"]
    u8 = ntrain.test(lm, [(thumbs, labels)], augment=GpuAugment(seed=0, recipe="none"))["test_acc"]   # uint8 test loader
    assert 0.0 <= u8 <= 1.0
    newest = max(st.best, key=lambda sp: sp[1])[1]
    again = ntrain.ViTLModule.load_from_checkpoint(newest, num_classes=5, pretrained=False,
                                                   model_name="google/vit-base-patch16-224", lr=1e-4, weight_decay=0.01).cuda()
    if newest.endswith(f"epoch=01_val_acc={st.history[1][3]:.4f}.ckpt"):
        assert ntrain.test(again, val)["test_acc"] == pytest.approx(acc)
    inner = ntrain.transform_checkpoint(newest, str(tmp_path / "nViT_epoch2.pth"))
    assert set(inner) == set(lm.vit.state_dict())
    # resume for one more epoch: epoch, step counters and AdamW moments continue
    lm2 = ntrain.ViTLModule(5, False, "google/vit-base-patch16-224", lr=1e-4, weight_decay=0.01, enable_mixup=True,
                            fused_optimizer=True).cuda()
    opt2 = lm2.configure_optimizers()
    last = [p for _, p in st.best if "epoch=01" in p][0]
    st2 = ntrain.fit(lm2, train, val, max_epochs=3, patience=0, checkpoint_dir=str(tmp_path / "r"), train_id="g",
                     optimizer=opt2, augment=GpuAugment(seed=5), ckpt_path=last)
    assert [h[0] for h in st2.history] == [2] and st2.global_step == 6 and opt2._step == 6


@pytest.mark.gpu
def test_train_main_end_to_end_on_image_folders(tmp_path):
    """ntrain.train_main (ntrain.py:160-248) from image folders: data module -> fit (device-side transform, CutMix/MixUp,
    fused steps) -> checkpoints -> test; then --test --restore on the written checkpoint reproduces the test metric."""
    from PIL import Image
    from touhouimageclassification_b200 import ntrain
    rng = np.random.default_rng(0)
    for split, n in (("train", 5), ("test", 2)):
        for c in ("a", "b"):
            os.makedirs(tmp_path / split / c)
            for i in range(n):
                Image.fromarray(rng.integers(0, 256, (256, 256, 3), dtype=np.uint8)).save(tmp_path / split / c / f"{i}.png")
    kw = dict(PRETRAINED=False, MODEL_NAME="google/vit-base-patch16-224", LR=1e-4, WEIGHT_DECAY=0.01, FULL_FINETUNE=True,
              BATCH_SIZE=4, NUM_WORKERS=0, TRAIN_SPLIT=0.8, DATA_DIR=str(tmp_path / "train"), MAX_EPOCHS=2, ENABLE_MIX_UP=True,
              ENABLE_AUGMENTATION=True, TRAIN_ID="m", PATIENCE=0, num_classes=2, test_dir=str(tmp_path / "test"),
              checkpoint_dir=str(tmp_path / "ck"))
    state, metrics = ntrain.train_main(**kw, argv=[])
    assert state.epoch == 1 and state.global_step == 4 and 0.0 <= metrics["test_acc"] <= 1.0
    files = sorted(os.listdir(tmp_path / "ck" / "m"))
    assert len(files) == 2 and files[0].startswith("checkpoint_m_epoch=00_val_acc=")
    last = str(tmp_path / "ck" / "m" / files[1])
    _, again = ntrain.train_main(**kw, argv=["--test", "--restore", last])
    assert again["test_acc"] == pytest.approx(metrics["test_acc"])
