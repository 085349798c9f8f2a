"""The reference's LightningDataModule (``AugmentedDataset``, ``TIC/ViT/ntrain.py:68-158``) with the transform moved
off the host: the loaders deliver decoded **uint8 thumbnails** ``[B, H, W, 3]`` and labels; the transform the
reference composes per sample on the CPU (``setup``, ntrain.py:93-147) is the matching :class:`~.augment.GpuAugment`
recipe, applied to the whole batch on the device inside the step (``ntrain.fit(..., augment=dm.augment())``).

Flag -> recipe (same precedence as the reference's ``if`` chain): ``enable_augmentation=False`` -> ``"none"``;
``only_grey_augmentation`` -> ``"grey"``; diversity + generalization -> ``"full"``; diversity -> ``"diversity"``;
generalization -> ``"generalization"``; neither raises, as the reference does.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset, random_split

IMG_EXTENSIONS = (".jpg", ".jpeg", ".png", ".ppm", ".bmp", ".pgm", ".tif", ".tiff", ".webp")


def recipe_from_flags(enable_augmentation: bool = True, enable_diversity: bool = True, enable_generalization: bool = True,
                      only_grey_augmentation: bool = False) -> str:
    if not enable_augmentation:
        return "none"
    if only_grey_augmentation:
        return "grey"
    if enable_diversity and enable_generalization:
        return "full"
    if enable_diversity:
        return "diversity"
    if enable_generalization:
        return "generalization"
    raise Exception("Must select diversity or generalization!")  # ntrain.py:136


class ThumbnailFolder(Dataset):
    """``torchvision.datasets.ImageFolder`` layout (one directory per class, classes sorted by name, ``class_to_idx``),
    yielding ``(uint8 [size, size, 3], label)``. The reference's sources are 256x256 thumbnails (report section 4.4);
    an image of another size is brought to ``size`` x ``size`` with PIL's bilinear filter so that a batch stacks."""

    def __init__(self, root: str, size: int = 256):
        self.root, self.size = root, int(size)
        self.classes = sorted(d.name for d in os.scandir(root) if d.is_dir())
        if not self.classes:
            raise FileNotFoundError(f"Couldn't find any class folder in {root}.")
        self.class_to_idx = {c: i for i, c in enumerate(self.classes)}
        self.samples: List[Tuple[str, int]] = []
        for c in self.classes:
            for dirpath, _, files in sorted(os.walk(os.path.join(root, c), followlinks=True)):
                for f in sorted(files):
                    if f.lower().endswith(IMG_EXTENSIONS):
                        self.samples.append((os.path.join(dirpath, f), self.class_to_idx[c]))
        self.targets = [t for _, t in self.samples]

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, i):
        from PIL import Image
        path, target = self.samples[i]
        with Image.open(path) as im:
            im = im.convert("RGB")
            if im.size != (self.size, self.size):
                im = im.resize((self.size, self.size), Image.BILINEAR)
            arr = np.asarray(im, dtype=np.uint8)
        return torch.from_numpy(arr.copy()), target


class AugmentedDataset:
    """Same constructor, ``setup(stage)`` and ``*_dataloader()`` methods as the reference's data module. The
    validation split shares the training transform, as in the reference (``random_split`` of one dataset,
    ntrain.py:141-144): ``augment()`` serves both; ``test_augment()`` is the plain resize + normalise of ``stage='test'``."""

    def __init__(self, train_path: str, test_path: str, batch_size: int = 8, train_split: float = 0.8, num_workers: int = 8,
                 image_size: int = 224, enable_augmentation: bool = True, enable_diversity: bool = True,
                 enable_generalization: bool = True, only_grey_augmentation: bool = False, thumbnail_size: int = 256,
                 seed: int = 42):
        self.train_path, self.test_path = train_path, test_path
        self.batch_size, self.train_split, self.num_workers = batch_size, train_split, num_workers
        self.image_size, self.thumbnail_size, self.seed = image_size, thumbnail_size, seed
        self.recipe = recipe_from_flags(enable_augmentation, enable_diversity, enable_generalization, only_grey_augmentation)
        self.dataset = self.train_dataset = self.val_dataset = self.test_dataset = None

    def setup(self, stage: str):
        if stage == "fit":
            self.dataset = ThumbnailFolder(self.train_path, self.thumbnail_size)
            train_size = int(len(self.dataset) * self.train_split)
            val_size = len(self.dataset) - train_size
            gen = torch.Generator().manual_seed(self.seed)  # the reference splits under L.seed_everything(42)
            self.train_dataset, self.val_dataset = random_split(self.dataset, [train_size, val_size], generator=gen)
        if stage == "test":
            self.test_dataset = ThumbnailFolder(self.test_path, self.thumbnail_size)

    def _loader(self, ds, shuffle):
        return DataLoader(ds, batch_size=self.batch_size, shuffle=shuffle, num_workers=self.num_workers,
                          pin_memory=torch.cuda.is_available())

    def train_dataloader(self):
        return self._loader(self.train_dataset, True)

    def val_dataloader(self):
        return self._loader(self.val_dataset, False)

    def test_dataloader(self):
        return self._loader(self.test_dataset, False)

    def augment(self, seed: int = 0):
        from .augment import GpuAugment
        return GpuAugment(seed=seed, size=self.image_size, recipe=self.recipe)

    def test_augment(self):
        from .augment import GpuAugment
        return GpuAugment(seed=0, size=self.image_size, recipe="none")
