"""GPU probe: attention kernels timing (ViT-L/16 224 shapes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from touhouimageclassification_b200 import ops
B, N, H = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (256, 197, 16)))
D = H * 64
torch.manual_seed(0)
qkv = torch.randn(B * N, 3 * D, device="cuda").bfloat16()
dctx = torch.randn(B * N, D, device="cuda").bfloat16()
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
ctx, lse = ops.attention_fwd(qkv, B, N, H)
t = timeit(lambda: ops.attention_fwd(qkv, B, N, H))
fl = 4.0 * B * H * N * N * 64
print(f"fwd B{B} N{N} H{H}: {t:.3f} ms  {fl / t / 1e9:.0f} TFLOP/s (useful)")
t = timeit(lambda: ops.attention_bwd(qkv, ctx, dctx, lse, B, N, H))
print(f"bwd B{B} N{N} H{H}: {t:.3f} ms  {2.5 * fl / t / 1e9:.0f} TFLOP/s (useful, 5 matmuls)")
bg = torch.zeros(3 * D, device="cuda")
t = timeit(lambda: ops.attention_bwd(qkv, ctx, dctx, lse, B, N, H, bias_grad=bg))
print(f"bwd+bias B{B} N{N} H{H}: {t:.3f} ms  {2.5 * fl / t / 1e9:.0f} TFLOP/s (useful, 5 matmuls)")
