"""GPU probe: LayerNorm / AdamW / elementwise kernel bandwidth at ViT-L/16 224 shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from touhouimageclassification_b200 import ops
M, D = 256 * 197, 1024
x = torch.randn(M, D, device="cuda")
g = torch.randn(D, device="cuda"); b = torch.randn(D, device="cuda")
dy = torch.randn(M, D, device="cuda").bfloat16()
dres = torch.randn(M, D, device="cuda")
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
y, _, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12)
t = timeit(lambda: ops.layernorm_fwd(x, g, b, 1e-12))
print(f"ln_fwd: {t:.4f} ms  {M*D*6/t/1e6:.0f} GB/s (alg 6 B/elem; includes output allocation)")
t = timeit(lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, dres))
print(f"ln_bwd: {t:.4f} ms  {M*D*16/t/1e6:.0f} GB/s (alg 16 B/elem; includes output allocation)")
