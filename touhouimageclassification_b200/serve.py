"""Mirror of the reference's inference library for the ViT path (``TIC/utils/serve.py:35-114`` [a20, a21] and
the forward part of ``web/runtime.py:97-128`` [a22]). Same function names, arguments and return values.

The reference runs these forwards in fp32 with batch 1 and a ``.item()`` per image; here the forward is the
bf16 tensor-core engine and ``predict_batch`` does one device->host copy per batch.
"""
from __future__ import annotations

import threading
from typing import Sequence

import torch

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
    """serve.serve (serve.py:83-114): one preprocessed image (batch dimension added) -> (class name, confidence)."""
    model.eval()
    idx_to_class = {v: k for k, v in class_to_idx.items()}
    with torch.no_grad():
        image_tensor = image_tensor.to(device)
        output = model(image_tensor)
        logits = output.logits if hasattr(output, 'logits') else output
        probabilities = torch.softmax(logits.float(), dim=1)
        confidence, predicted_idx = torch.max(probabilities, 1)
        predicted_class = idx_to_class[predicted_idx.item()]
    return predicted_class, confidence.item()


_predict_lock = threading.Lock()


def preprocess_u8(images_u8: torch.Tensor, mean, std, size: int = 224) -> torch.Tensor:
    """The reference's inference transform (``utils/preprocess.py:73-77`` [a23]: ``Resize((224,224))`` -> ``ToTensor`` ->
    ``Normalize(mean, std)`` with the dataset statistics from ``meta_mean_std.pth``) on the GPU: uint8 NHWC batch ->
    bf16 patch rows for ``engine_forward(patches=...)``. Same fused kernel as the training augmentation, recipe "none"."""
    from .augment import GpuAugment
    mean = [float(m) for m in mean]
    std = [float(s) for s in std]
    return GpuAugment(seed=0, size=size, recipe="none", mean=mean, std=std)(images_u8, first_sample=0)


def predict_batch_u8(model: ViTForImageClassification, images_u8: torch.Tensor, mean, std, idx_to_class=None,
                     max_batch_size: int = 1024):
    """``serve_batch`` (runtime.py:235-251) end to end on the device: raw uint8 NHWC thumbnails in, (class, confidence)
    out -- resize + normalise + patchify in one kernel, bf16 engine forward, softmax/max, one host copy per chunk."""
    model.eval()
    results = []
    with torch.no_grad(), _predict_lock:
        for i in range(0, images_u8.shape[0], max_batch_size):
            chunk = images_u8[i:i + max_batch_size].to(model._arena.device, non_blocking=True)
            logits = model.engine_forward(patches=preprocess_u8(chunk, mean, std, model.config.image_size))
            prob = torch.softmax(logits, dim=1)
            conf, idx = torch.max(prob, 1)
            for c, k in torch.stack([conf, idx.to(conf.dtype)], 1).cpu().tolist():
                k = int(k)
                results.append((idx_to_class[k] if idx_to_class is not None else k, c))
    return results


def predict_batch(model: ViTForImageClassification, image_batch: torch.Tensor, idx_to_class=None,
                  max_batch_size: int = 1024):
    """Forward part of ``ModelDaemon.predict`` / ``serve_batch`` (runtime.py:113-124, 243-246): a stacked batch of
    preprocessed images -> list of (class, confidence). Chunks by ``max_batch_size`` and is safe to call from
    several threads (the Flask app calls predict outside its lock, runtime.py:237-246)."""
    model.eval()
    results = []
    with torch.no_grad(), _predict_lock:
        for i in range(0, image_batch.shape[0], max_batch_size):
            chunk = image_batch[i:i + max_batch_size].to(model._arena.device, non_blocking=True)
            logits = model.engine_forward(chunk, training=False)
            prob = torch.softmax(logits, dim=1)
            conf, idx = torch.max(prob, 1)
            both = torch.stack([conf, idx.to(conf.dtype)], 1).cpu()
            for c, k in both.tolist():
                k = int(k)
                results.append((idx_to_class[k] if idx_to_class is not None else k, c))
    return results
