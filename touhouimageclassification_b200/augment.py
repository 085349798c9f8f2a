"""Host side of the fused GPU augmentation (``csrc/augment.cu``): stands where the reference's
``AugmentedDataset`` train transform stands (``TIC/ViT/ntrain.py:104-112`` [a18]) -- but on uint8 batches that are
already on the device, producing the bf16 patch rows the engine's patch-projection GEMM consumes.

``GpuAugment(seed)(images_u8_nhwc)`` -> ``patches`` (bf16 ``[B*196, 768]``). Parameters are sampled by the native
host sampler from a counter-based RNG keyed by ``(seed, global sample index)``, so any data-parallel rank or a CPU
restatement can reproduce them.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import c_i64, c_int, c_void_p

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
RECIPES = {"full": 0, "generalization": 1, "diversity": 2, "grey": 3, "none": 4}


def sample_params(seed: int, first_sample: int, batch: int, height: int, width: int, size: int = 224,
                  recipe: str = "full"):
    """Per-sample parameter records: int32 ``[B,16]`` and float32 ``[B,4]`` (host numpy arrays)."""
    ints = np.zeros((batch, 16), dtype=np.int32)
    floats = np.zeros((batch, 4), dtype=np.float32)
    _lib.check(_lib.load().tic_augment_sample_params(
        c_i64(seed), c_i64(first_sample), c_int(batch), c_int(height), c_int(width), c_int(size), c_int(RECIPES[recipe]),
        ints.ctypes.data_as(ctypes.c_void_p), floats.ctypes.data_as(ctypes.c_void_p)))
    return ints, floats


class GpuAugment:
    def __init__(self, seed: int = 0, size: int = 224, recipe: str = "full", mean=IMAGENET_MEAN, std=IMAGENET_STD):
        self.seed, self.size, self.recipe = int(seed), int(size), recipe
        self._mean = (ctypes.c_float * 3)(*[float(np.float32(m)) for m in mean])
        self._std = (ctypes.c_float * 3)(*[float(np.float32(s)) for s in std])
        self.samples_seen = 0
        self.rank, self.world_size = 0, 1

    def shard(self, rank: int, world_size: int):
        """Data-parallel use: every rank's batch of B is the slice [rank * B, (rank + 1) * B) of a global batch of
        world_size * B, so the augmentation parameters stay a function of (seed, GLOBAL sample index) and no two ranks
        draw the same ones."""
        self.rank, self.world_size = int(rank), int(world_size)
        return self

    def _prepare(self, images: torch.Tensor, first_sample):
        if not images.is_cuda or images.dtype != torch.uint8 or images.dim() != 4 or images.shape[-1] != 3:
            raise ValueError("GpuAugment takes a CUDA uint8 NHWC batch [B, H, W, 3] (there is no CPU fallback)")
        images = images.contiguous()
        B, H, W, _ = images.shape
        if first_sample is None:
            first_sample = self.samples_seen + self.rank * B
            self.samples_seen += B * self.world_size
        ints, floats = sample_params(self.seed, first_sample, B, H, W, self.size, self.recipe)
        dev = images.device
        ints_d = torch.from_numpy(ints).to(dev, non_blocking=True)
        floats_d = torch.from_numpy(floats).to(dev, non_blocking=True)
        return images, ints, floats, ints_d, floats_d

    def __call__(self, images: torch.Tensor, first_sample: int = None, return_pixels: bool = False):
        with torch.cuda.device(images.device):  # the launch goes to the CURRENT device: make it the images' device
            return self._patchify(images, first_sample, return_pixels)

    def _patchify(self, images, first_sample, return_pixels):
        images, ints, floats, ints_d, floats_d = self._prepare(images, first_sample)
        B, H, W, _ = images.shape
        dev = images.device
        G = self.size // 16
        patches = torch.empty((B * G * G, 768), dtype=torch.bfloat16, device=dev)
        pixels = torch.empty((B, self.size, self.size, 3), dtype=torch.uint8, device=dev) if return_pixels else None
        _lib.check(_lib.load().tic_augment_patchify(
            c_void_p(images.data_ptr()), c_int(B), c_int(H), c_int(W), c_void_p(ints_d.data_ptr()),
            c_void_p(floats_d.data_ptr()), c_int(self.size), self._mean, self._std, c_void_p(patches.data_ptr()),
            c_void_p(0 if pixels is None else pixels.data_ptr()), c_void_p(torch.cuda.current_stream().cuda_stream)))
        if return_pixels:
            return patches, pixels, (ints, floats)
        return patches

    def tensor(self, images: torch.Tensor, first_sample: int = None) -> torch.Tensor:
        """The batch the reference's DataLoader would hand to ``training_step``: fp32 ``[B, 3, size, size]``, normalised
        (ntrain.py:104-112 ends in ToTensor + Normalize) -- the input of the per-batch CutMix / MixUp (ntrain.py:45-46)."""
        with torch.cuda.device(images.device):
            return self._tensor(images, first_sample)

    def _tensor(self, images, first_sample):
        images, ints, floats, ints_d, floats_d = self._prepare(images, first_sample)
        B, H, W, _ = images.shape
        out = torch.empty((B, 3, self.size, self.size), dtype=torch.float32, device=images.device)
        _lib.check(_lib.load().tic_augment_tensor(
            c_void_p(images.data_ptr()), c_int(B), c_int(H), c_int(W), c_void_p(ints_d.data_ptr()),
            c_void_p(floats_d.data_ptr()), c_int(self.size), self._mean, self._std, c_void_p(out.data_ptr()),
            c_void_p(0), c_void_p(torch.cuda.current_stream().cuda_stream)))
        return out
