"""Mirror of the reference's inference library for the ViT path (``TIC/utils/serve.py:35-114`` [a20, a21] and
the forward part of ``web/runtime.py:97-128`` [a22]). Same function names, arguments and return values.

The reference runs these forwards in fp32 with batch 1 and a ``.item()`` per image; here the forward is the
bf16 tensor-core engine and ``predict_batch`` does one device->host copy per batch.
"""
from __future__ import annotations

import os
import queue
import threading
from typing import Callable, List

import torch

from . import ops
from .model import ViT, ViTForImageClassification


def get_model(model_type: str, num_classes: int):
    """serve.get_model (serve.py:35-45) for the ViT model types."""
    model_type = model_type.lower().replace('_', '-')
    if model_type == 'vit-base':
        return ViT(num_classes=num_classes, pretrained=False, model_name='google/vit-base-patch16-224-in21k', wrap_model_name=False)
    elif model_type == 'vit-large':
        return ViT(num_classes=num_classes, pretrained=False, model_name='google/vit-large-patch16-224-in21k', wrap_model_name=False)
    raise ValueError(f"Unsupported model type: {model_type}")


def extract_state_dict(ckpt):
    """Checkpoint containers seen in the reference: bare dict (ntrain.py:193), tuple
    ``(model_sd, optim_sd[, sched_sd])`` (finetune.py:249-258), ``{"model": ...}`` (extract_ckpt.py:19-26) and
    Lightning ``{"state_dict": {"vit.<key>": ...}}`` (ntrain.py:16-27)."""
    if isinstance(ckpt, (tuple, list)):
        ckpt = ckpt[0]
    if isinstance(ckpt, dict) and "state_dict" in ckpt and isinstance(ckpt["state_dict"], dict):
        ckpt = ckpt["state_dict"]
    if isinstance(ckpt, dict) and "model" in ckpt and isinstance(ckpt["model"], dict):
        ckpt = ckpt["model"]
    if isinstance(ckpt, dict) and ckpt and all(k.startswith("vit.vit.") or k.startswith("vit.classifier.") for k in ckpt):
        ckpt = {k[len("vit."):]: v for k, v in ckpt.items()}
    return ckpt


def load_model(model_type: str, num_classes: int, weights_path: str = None, device: str = 'cuda',
               precision: str = 'bf16'):
    """serve.load_model (serve.py:47-81): tuple checkpoints -> ``[0]``, strict ``load_state_dict``, ``.to(device)``.
    ``precision='fp32'`` selects the fp32-accurate engine (what the reference's no-autocast forward computes,
    serve.py:99-101) instead of the bf16 tensor-core one."""
    model_type = model_type.lower().replace('_', '-')
    model = get_model(model_type, num_classes)
    if weights_path is None:
        raise ValueError(f"No default checkpoint found for model type: {model_type}")
    ckpt = torch.load(weights_path, map_location="cpu", weights_only=False)
    model.load_state_dict(extract_state_dict(ckpt))
    model.to(device)
    model.set_precision(precision)
    return model


def serve(model, image_tensor, class_to_idx, device: str = 'cuda'):
    """serve.serve (serve.py:83-114): one preprocessed image (batch dimension added) -> (class name, confidence).

    Precision: the reference runs this forward in plain fp32 (no autocast, serve.py:99-101). Here the arithmetic is the
    model's ``precision``: ``'bf16'`` (tensor-core engine, the default of ``load_model``) or ``'fp32'`` (logits within 1e-4
    of the reference's; ``load_model(..., precision='fp32')``). softmax + max run in one kernel (``tic_softmax_top1``)."""
    model.eval()
    idx_to_class = {v: k for k, v in class_to_idx.items()}
    with torch.no_grad():
        image_tensor = image_tensor.to(device)
        output = model(image_tensor)
        logits = output.logits if hasattr(output, 'logits') else output
        confidence, predicted_idx, _ = ops.softmax_top1(logits)
        pair = torch.stack([confidence, predicted_idx.to(confidence.dtype)]).cpu().tolist()  # one device->host copy
        predicted_class = idx_to_class[int(pair[1][0])]
    return predicted_class, pair[0][0]


def _model_lock(model):
    """Serialises the forwards of ONE model (its workspaces and graphs are per model); replicas on other GPUs have their
    own lock and run concurrently."""
    lock = getattr(model, "_lock", None)
    if lock is None:
        lock = model._lock = threading.RLock()
    return lock


def _forward_logits(model, pixels=None, patches=None):
    """Engine forward of a preprocessed batch in the model's precision (``set_precision``): bf16 tensor-core engine, or the
    fp32-accurate engine, which takes pixels only."""
    if getattr(model, "precision", "bf16") == "fp32":
        if pixels is None:
            raise RuntimeError("precision='fp32' serves from preprocessed fp32 pixels; the uint8 -> bf16 patch pipeline "
                               "(preprocess_u8) rounds its output to bf16 -- use preprocess_u8_pixels or precision='bf16'")
        return model.engine_forward_f32(pixels)
    if patches is not None:
        return model.engine_forward(patches=patches)
    return model.engine_forward(pixels, training=False)


def _top1_to_host(logits, idx_to_class, results):
    conf, idx, _ = ops.softmax_top1(logits)
    for c, k in torch.stack([conf, idx.to(conf.dtype)], 1).cpu().tolist():  # one device->host copy per chunk
        k = int(k)
        results.append((idx_to_class[k] if idx_to_class is not None else k, c))


def preprocess_u8(images_u8: torch.Tensor, mean, std, size: int = 224) -> torch.Tensor:
    """The reference's inference transform (``utils/preprocess.py:73-77`` [a23]: ``Resize((224,224))`` -> ``ToTensor`` ->
    ``Normalize(mean, std)`` with the dataset statistics from ``meta_mean_std.pth``) on the GPU: uint8 NHWC batch ->
    bf16 patch rows for ``engine_forward(patches=...)``. Same fused kernel as the training augmentation, recipe "none"."""
    from .augment import GpuAugment
    mean = [float(m) for m in mean]
    std = [float(s) for s in std]
    with torch.cuda.device(images_u8.device):
        return GpuAugment(seed=0, size=size, recipe="none", mean=mean, std=std)(images_u8, first_sample=0)


def preprocess_u8_pixels(images_u8: torch.Tensor, mean, std, size: int = 224) -> torch.Tensor:
    """Same transform, yielding what the reference's transform yields: the normalised fp32 tensor [B, 3, size, size]
    (input of the fp32-accurate engine)."""
    from .augment import GpuAugment
    mean = [float(m) for m in mean]
    std = [float(s) for s in std]
    with torch.cuda.device(images_u8.device):
        return GpuAugment(seed=0, size=size, recipe="none", mean=mean, std=std).tensor(images_u8, first_sample=0)


def predict_batch_u8(model: ViTForImageClassification, images_u8: torch.Tensor, mean, std, idx_to_class=None,
                     max_batch_size: int = 1024):
    """``serve_batch`` (runtime.py:235-251) end to end on the device: raw uint8 NHWC thumbnails in, (class, confidence)
    out -- resize + normalise + patchify in one kernel, bf16 engine forward, softmax/max, one host copy per chunk."""
    model.eval()
    results = []
    dev = model._arena.device
    fp32 = getattr(model, "precision", "bf16") == "fp32"
    with torch.no_grad(), _model_lock(model):
        for i in range(0, images_u8.shape[0], max_batch_size):
            chunk = images_u8[i:i + max_batch_size].to(dev, non_blocking=True)
            if fp32:
                logits = _forward_logits(model, pixels=preprocess_u8_pixels(chunk, mean, std, model.config.image_size))
            else:
                logits = _forward_logits(model, patches=preprocess_u8(chunk, mean, std, model.config.image_size))
            _top1_to_host(logits, idx_to_class, results)
    return results


def predict_batch(model: ViTForImageClassification, image_batch: torch.Tensor, idx_to_class=None,
                  max_batch_size: int = 1024):
    """Forward part of ``ModelDaemon.predict`` / ``serve_batch`` (runtime.py:113-124, 243-246): a stacked batch of
    preprocessed images -> list of (class, confidence). Chunks by ``max_batch_size`` and is safe to call from
    several threads (the Flask app calls predict outside its lock, runtime.py:237-246). Runs in the model's precision
    (``load_model(..., precision=...)``): the reference's is fp32, the default here is the bf16 tensor-core engine."""
    model.eval()
    results = []
    dev = model._arena.device
    with torch.no_grad(), _model_lock(model):
        for i in range(0, image_batch.shape[0], max_batch_size):
            chunk = image_batch[i:i + max_batch_size].to(dev, non_blocking=True)
            _top1_to_host(_forward_logits(model, pixels=chunk), idx_to_class, results)
    return results


class ReplicaPool:
    """Per-GPU inference replicas (BASELINE config 4 / SURVEY 8e: "one module per GPU, requests round-robined, no
    collective"). One copy of the model per device, one worker thread per replica; ``predict_u8`` / ``predict`` split a
    request into per-replica chunks, run them concurrently and return the results in request order."""

    def __init__(self, model: ViTForImageClassification, devices=None):
        if devices is None:
            devices = [f"cuda:{i}" for i in range(torch.cuda.device_count())]
        if not devices:
            raise RuntimeError("ReplicaPool needs at least one CUDA device (there is no CPU path)")
        self.devices = [torch.device(d) for d in devices]
        self.replicas = []
        for d in self.devices:
            if model._arena.device == d:
                rep = model
            else:
                # a fresh module from the same config + the same weights (the model's arenas, graphs, workspaces and
                # locks are per device and are not copied)
                rep = type(model)(model.config)
                rep.load_state_dict(model.state_dict(), strict=True)
                rep = rep.to(d)
                rep._lock = threading.RLock()
                rep.set_precision(model.precision)
            self.replicas.append(rep.eval())
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(max_workers=len(self.replicas), thread_name_prefix="tic-replica")
        self._next = 0

    def __len__(self):
        return len(self.replicas)

    def _split(self, n: int, chunk: int):
        """Contiguous chunks of at most ``chunk`` items, dealt round-robin to the replicas (continuing from the last call)."""
        plan = []
        for start in range(0, n, chunk):
            plan.append((self._next % len(self.replicas), start, min(n, start + chunk)))
            self._next += 1
        return plan

    def _run(self, fn, n: int, chunk: int):
        if chunk is None:
            chunk = max(1, -(-n // len(self.replicas)))
        plan = self._split(n, chunk)
        futures = [self._pool.submit(fn, self.replicas[r], a, b) for r, a, b in plan]
        out = []
        for f in futures:
            out.extend(f.result())
        return out

    def predict_u8(self, images_u8: torch.Tensor, mean, std, idx_to_class=None, chunk: int = None):
        """uint8 NHWC thumbnails (host, ideally pinned) -> list of (class, confidence), all replicas working at once."""
        def work(rep, a, b):
            with torch.cuda.device(rep._arena.device):
                return predict_batch_u8(rep, images_u8[a:b], mean, std, idx_to_class, max_batch_size=b - a)
        return self._run(work, images_u8.shape[0], chunk)

    def predict(self, image_batch: torch.Tensor, idx_to_class=None, chunk: int = None):
        """Preprocessed fp32 NCHW batch -> list of (class, confidence)."""
        def work(rep, a, b):
            with torch.cuda.device(rep._arena.device):
                return predict_batch(rep, image_batch[a:b], idx_to_class, max_batch_size=b - a)
        return self._run(work, image_batch.shape[0], chunk)

    def close(self):
        self._pool.shutdown(wait=True)


# ----------------------------------------------------------------------------------------------------------------------
# Batched full_judge / filter pipeline (TIC/utils/serve.py:158-230, TIC/utils/filter.py:17-27) [section 8f rank 2]
# ----------------------------------------------------------------------------------------------------------------------
IMAGE_EXTENSIONS = ('.jpg', '.jpeg', '.png', '.bmp', '.gif')


def full_judge(model, transforms, class_to_idx, args=None, image=None, device=None, output=None, batch_size: int = 1024,
               mean=None, std=None):
    """``serve.full_judge``: walk ``image`` (a file or a class-per-directory tree), predict every picture, write the
    reference's CSV (``filename,predicted_class,confidence,actual_class,correct,path``) and return the accuracy.

    The reference forwards one image at a time with a host sync per image; here pictures are decoded on the host,
    stacked into batches of up to ``batch_size`` and forwarded together (one device->host copy per batch). With
    ``transforms=None`` and ``mean`` / ``std`` given, same-sized pictures skip the CPU transform entirely: their uint8
    pixels go to the GPU and ``preprocess_u8`` resizes, normalises and patchifies them in one kernel."""
    from PIL import Image
    if args:
        image, device, output = args.image, args.device, args.output
    device = device or 'cuda'
    idx_to_class = {v: k for k, v in class_to_idx.items()}
    if os.path.isfile(image):
        tensor = transforms(Image.open(image).convert('RGB')).unsqueeze(0)
        predicted_class, confidence = serve(model, tensor, class_to_idx, device)
        if not output:
            print(f"Prediction: {predicted_class} (Confidence: {confidence:.4f})")
        return None
    files = []
    for root, _dirs, names in os.walk(image):
        for name in names:
            if os.path.splitext(name)[1].lower() in IMAGE_EXTENSIONS:
                files.append((name, os.path.basename(root), os.path.join(root, name)))
    print(f"Total images to process: {len(files)}")
    out = open(output, 'w') if output else None
    if out:
        print("filename,predicted_class,confidence,actual_class,correct,path", file=out)
    cnt = correct_cnt = 0

    def flush(entries, tensors=None, u8=None):
        nonlocal cnt, correct_cnt
        if not entries:
            return
        if u8 is not None:
            import numpy as np
            batch = torch.from_numpy(np.stack(u8))
            results = predict_batch_u8(model, batch, mean, std, idx_to_class, max_batch_size=batch_size)
        else:
            results = predict_batch(model, torch.stack(tensors), idx_to_class, max_batch_size=batch_size)
        for (name, label, path), (pred, conf) in zip(entries, results):
            cnt += 1
            correct_cnt += (pred == label)
            if out:
                out.write(f"{name},{pred},{conf:.4f},{label},{pred == label},{path}\n")
            else:
                print(f"--- {name}: {pred} (Confidence: {conf:.4f}) Correct: {pred == label}")

    entries, tensors = [], []
    by_shape = {}
    for name, label, path in files:
        try:
            img = Image.open(path).convert('RGB')
        except Exception as e:  # same policy as the reference: report and continue
            print(f"Error processing image {name}: {e}")
            continue
        if transforms is None:
            import numpy as np
            arr = np.asarray(img, dtype=np.uint8)
            key = arr.shape[:2]
            ents, arrs = by_shape.setdefault(key, ([], []))
            ents.append((name, label, path))
            arrs.append(arr)
            if len(ents) >= batch_size:
                flush(ents, u8=arrs)
                by_shape[key] = ([], [])
        else:
            entries.append((name, label, path))
            tensors.append(transforms(img))
            if len(entries) >= batch_size:
                flush(entries, tensors=tensors)
                entries, tensors = [], []
    flush(entries, tensors=tensors)
    for ents, arrs in by_shape.values():
        flush(ents, u8=arrs)
    if out:
        out.close()
    if cnt == 0:
        print("Total images processed: 0")
        return 0.0
    print(f"Total images processed: {cnt}, Correct predictions: {correct_cnt}, Accuracy: {correct_cnt / cnt * 100:.2f}%")
    return correct_cnt / cnt


def filter_csv(csv_file: str, output_directory: str):
    """``filter.filter`` (TIC/utils/filter.py:17-27): copy every correctly predicted picture of a full_judge CSV into
    ``output_directory/<class>/``. Returns (total rows, copied)."""
    import csv
    import shutil
    tot = cnt = 0
    with open(csv_file, 'r') as f:
        for row in csv.DictReader(f):
            tot += 1
            if row['predicted_class'].strip() == row['actual_class'].strip():
                cnt += 1
                dst = os.path.join(output_directory, row['actual_class'].strip(), os.path.basename(row['path'].strip()))
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copy(row['path'].strip(), dst)
    print(f"Tot:{tot}, Copy cnt:{cnt}, Rate:{cnt / tot if tot else 0.0}")
    return tot, cnt


# ----------------------------------------------------------------------------------------------------------------------
# Web daemon adapter (web/runtime.py:34-128,235-251) [section 8f rank 3]
# ----------------------------------------------------------------------------------------------------------------------
class BatchingPredictor:
    """Dynamic batching across concurrent callers: requests from any number of threads are queued, one worker thread
    drains up to ``max_batch_size`` items per forward and hands every caller its own slice back. ``forward`` maps a
    list of items to a list of results (e.g. a closure over ``predict_batch``)."""

    def __init__(self, forward: Callable[[List], List], max_batch_size: int = 64, max_wait_s: float = 0.002):
        self.forward, self.max_batch_size, self.max_wait_s = forward, max_batch_size, max_wait_s
        self._q: "queue.Queue" = queue.Queue()
        self._stop = threading.Event()
        self.batches = []  # sizes of the forwards issued so far (observability / tests)
        self._worker = threading.Thread(target=self._run, daemon=True)
        self._worker.start()

    def _run(self):
        while not self._stop.is_set():
            try:
                first = self._q.get(timeout=0.05)
            except queue.Empty:
                continue
            pending = [first]
            n = len(first[0])
            while n < self.max_batch_size:
                try:
                    nxt = self._q.get(timeout=self.max_wait_s)
                except queue.Empty:
                    break
                pending.append(nxt)
                n += len(nxt[0])
            items = [it for req in pending for it in req[0]]
            try:
                results = []
                for i in range(0, len(items), self.max_batch_size):
                    chunk = items[i:i + self.max_batch_size]
                    self.batches.append(len(chunk))
                    results.extend(self.forward(chunk))
                err = None
            except Exception as e:  # hand the failure to every waiting caller
                results, err = None, e
            off = 0
            for req_items, box, done in pending:
                box.append(err if err is not None else results[off:off + len(req_items)])
                off += len(req_items)
                done.set()

    def __call__(self, items: List) -> List:
        box, done = [], threading.Event()
        self._q.put((list(items), box, done))
        done.wait()
        if isinstance(box[0], Exception):
            raise box[0]
        return box[0]

    def close(self):
        self._stop.set()
        self._worker.join(timeout=1.0)


class ModelDaemon:
    """Drop-in for ``web/runtime.py:ModelDaemon`` (``predict`` on a PIL image or a list of them -> (class, confidence)
    or a list of pairs), with the forward on the engine and concurrent requests coalesced into shared batches.
    ``transforms`` is the reference's CPU transform (``get_transforms``); pass ``mean`` / ``std`` instead to run the
    resize + normalise on the GPU for same-sized pictures."""

    MAX_BATCH_SIZE = 64

    def __init__(self, model, class_to_idx, transforms=None, mean=None, std=None, max_batch_size: int = None):
        if transforms is None and (mean is None or std is None):
            raise ValueError("ModelDaemon needs either the reference's transforms or the dataset mean / std")
        self.model = model.eval()
        self.class_to_idx = class_to_idx
        self.idx_to_class = {v: k for k, v in class_to_idx.items()}
        self.transforms, self.mean, self.std = transforms, mean, std
        self.lock = threading.Lock()
        self._batcher = BatchingPredictor(self._forward, max_batch_size or self.MAX_BATCH_SIZE)

    def _forward(self, images):
        if self.transforms is not None:
            batch = torch.stack([self.transforms(im) for im in images])
            return predict_batch(self.model, batch, self.idx_to_class, max_batch_size=len(images))
        import numpy as np
        arrs = [np.asarray(im, dtype=np.uint8) for im in images]
        out = [None] * len(arrs)
        groups = {}
        for i, a in enumerate(arrs):
            groups.setdefault(a.shape[:2], []).append(i)
        for idxs in groups.values():
            res = predict_batch_u8(self.model, torch.from_numpy(np.stack([arrs[i] for i in idxs])), self.mean, self.std,
                                   self.idx_to_class, max_batch_size=len(idxs))
            for i, r in zip(idxs, res):
                out[i] = r
        return out

    def predict(self, images):
        is_single = not isinstance(images, list)
        if is_single:
            images = [images]
        images = [im.convert('RGB') if getattr(im, 'mode', 'RGB') != 'RGB' else im for im in images]
        results = self._batcher(images)
        return results[0] if is_single else results

    def stop(self):
        self._batcher.close()
