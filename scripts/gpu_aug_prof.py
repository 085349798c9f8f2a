"""GPU probe: kernel time (tic_prof events, no host sampler) of the fused augmentation kernel per recipe, B thumbnails 256x256."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from touhouimageclassification_b200 import _lib
from touhouimageclassification_b200.augment import GpuAugment
lib = _lib.load()
lib.tic_prof_collect.restype = ctypes.c_int64
for B in (256, 1024):
    imgs = torch.randint(0, 256, (B, 256, 256, 3), dtype=torch.uint8, device="cuda")
    for recipe, mode in (("full", "patches"), ("full", "tensor"), ("none", "patches")):
        aug = GpuAugment(seed=1, size=224, recipe=recipe)
        fn = (lambda: aug(imgs, first_sample=0)) if mode == "patches" else (lambda: aug.tensor(imgs, first_sample=0))
        for _ in range(3): fn()
        torch.cuda.synchronize()
        lib.tic_prof_enable(1)
        for _ in range(10): fn()
        torch.cuda.synchronize()
        buf = ctypes.create_string_buffer(1 << 16)
        n = lib.tic_prof_collect(buf, ctypes.c_int64(len(buf)))
        lib.tic_prof_enable(0)
        for ln in buf.raw[:n].decode().splitlines():
            name, cnt, ms, fl, by = ln.split("\t")
            ms = float(ms) / int(cnt)
            print(f"B={B:5d} recipe={recipe:5s} out={mode:8s} {name:20s} {ms:.4f} ms  {B / ms * 1e3:9.0f} img/s  {float(by) / int(cnt) / ms / 1e6:7.1f} GB/s (algorithmic)")
