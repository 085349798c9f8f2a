"""GPU probe: per-kernel times of the attention forward / backward at one shape (tic_prof events on the launching stream)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from touhouimageclassification_b200 import _lib, ops
B, N, H = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (128, 577, 16)))
D = H * 64
torch.manual_seed(0)
qkv = torch.randn(B * N, 3 * D, device="cuda").bfloat16()
dctx = torch.randn(B * N, D, device="cuda").bfloat16()
bg = torch.zeros(3 * D, device="cuda")
lib = _lib.load()
lib.tic_prof_collect.restype = ctypes.c_int64
for _ in range(3):
    ctx, lse = ops.attention_fwd(qkv, B, N, H)
    ops.attention_bwd(qkv, ctx, dctx, lse, B, N, H, bias_grad=bg)
torch.cuda.synchronize()
lib.tic_prof_enable(1)
iters = 10
for _ in range(iters):
    ctx, lse = ops.attention_fwd(qkv, B, N, H)
    ops.attention_bwd(qkv, ctx, dctx, lse, B, N, H, bias_grad=bg)
torch.cuda.synchronize()
buf = ctypes.create_string_buffer(1 << 16)
n = lib.tic_prof_collect(buf, ctypes.c_int64(len(buf)))
lib.tic_prof_enable(0)
print(f"B{B} N{N} H{H} variant={os.environ.get('TIC_LIB_VARIANT', '')}")
for ln in buf.raw[:n].decode().splitlines():
    name, cnt, ms, fl, by = ln.split("\t")
    ms = float(ms) / int(cnt)
    print(f"  {name:22s} {ms:.4f} ms  {float(fl) / int(cnt) / ms / 1e9:8.1f} TFLOP/s  {float(by) / int(cnt) / ms / 1e6:8.1f} GB/s")
