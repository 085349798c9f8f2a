import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from touhouimageclassification_b200 import ops
B, N, H = 64, 577, 16
D = H * 64
qkv = torch.randn(B * N, 3 * D, device="cuda").bfloat16()
dctx = torch.randn(B * N, D, device="cuda").bfloat16()
ctx, lse = ops.attention_fwd(qkv, B, N, H)
for _ in range(2):
    ops.attention_bwd(qkv, ctx, dctx, lse, B, N, H)
torch.cuda.synchronize()
