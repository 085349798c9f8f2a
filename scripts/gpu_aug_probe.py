"""GPU probe: the augmentation / mix / patchify kernels at the training shape (256 thumbnails of 256x256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from touhouimageclassification_b200.augment import GpuAugment
from touhouimageclassification_b200 import ops
B = 256
imgs = torch.randint(0, 256, (B, 256, 256, 3), dtype=torch.uint8, device="cuda")
aug = GpuAugment(seed=1, size=224, recipe="full")
x = torch.randn(B, 3, 224, 224, device="cuda")
y = torch.randint(0, 120, (B,), device="cuda")
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
t = timeit(lambda: aug(imgs, first_sample=0))
print(f"augment_patchify (full recipe): {t:.4f} ms  {B/t*1e3:.0f} img/s  {(B*256*256*3 + B*196*768*2)/t/1e6:.0f} GB/s (alg bytes; includes host sampler + H2D of params)")
t = timeit(lambda: ops.mix_batch(x, y, 120, 1, 0.3, (0, 0, 0, 0), 0.3, want_pixels=False, want_patches=True))
print(f"mix_patchify (mixup): {t:.4f} ms  {(2*x.numel()*4 + B*196*768*2)/t/1e6:.0f} GB/s")
t = timeit(lambda: ops.patchify_f32(x))
print(f"patchify_f32: {t:.4f} ms  {(x.numel()*4 + B*196*768*2)/t/1e6:.0f} GB/s")
