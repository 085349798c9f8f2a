"""For ncu: one warm-up + one launch of each GEMM variant of a ViT-L step at B=256 (M = 50432)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from touhouimageclassification_b200 import ops
M, D, F = 50432, 1024, 4096
dev = "cuda"
torch.manual_seed(0)
h = (torch.randn(M, D, device=dev) * 0.5).bfloat16()
w1 = (torch.randn(F, D, device=dev) * 0.03).bfloat16(); b1 = torch.randn(F, device=dev) * 0.1
dy = (torch.randn(M, D, device=dev) * 0.1).bfloat16()
w2 = (torch.randn(D, F, device=dev) * 0.03).bfloat16()
wq = (torch.randn(3 * D, D, device=dev) * 0.03).bfloat16(); bq = torch.randn(3 * D, device=dev) * 0.1
wo = (torch.randn(D, D, device=dev) * 0.03).bfloat16(); bo = torch.randn(D, device=dev) * 0.1
xres = torch.randn(M, D, device=dev)
act, dact = ops.gemm_bf16(h, w1, bias=b1, epilogue=ops.EPI_BF16_GELU)            # launch 0 (also the GELU warm-up)
dW = torch.zeros(F, D, device=dev)
variants = [
    ("fc1+GELU", lambda: ops.gemm_bf16(h, w1, bias=b1, epilogue=ops.EPI_BF16_GELU)),
    ("fc2 dgrad x GELU'", lambda: ops.gemm_bf16(dy, w2, b_mn_major=True, aux=dact, epilogue=ops.EPI_BF16_DGELU)),
    ("QKV forward", lambda: ops.gemm_bf16(h, wq, bias=bq, epilogue=ops.EPI_BF16)),
    ("out-proj + residual K=1024", lambda: ops.gemm_bf16(h, wo, bias=bo, aux=xres, epilogue=ops.EPI_F32_RESID)),
    ("fc2 + residual K=4096", lambda: ops.gemm_bf16(act, w2, bias=bo, aux=xres, epilogue=ops.EPI_F32_RESID)),
    ("fc1 wgrad split-K", lambda: ops.gemm_bf16(act, h, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32_ATOMIC, out=dW, splits=7)),
]
for name, fn in variants:   # launches 1..12: warm-up, then the one to read
    fn(); fn()
torch.cuda.synchronize()
print("order:", [n for n, _ in variants])
