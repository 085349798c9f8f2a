"""GPU probe for ncu / timing: the epilogue-bound GEMMs of a ViT-L step at B=256 (M = 50432): fc1 + GELU (forward),
fc2-dgrad x GELU', plain QKV forward for reference. CUDA-event timing per variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from touhouimageclassification_b200 import ops
M, D, F = 50432, 1024, 4096
dev = "cuda"
torch.manual_seed(0)
h = (torch.randn(M, D, device=dev) * 0.5).bfloat16()
w1 = (torch.randn(F, D, device=dev) * 0.03).bfloat16()
b1 = torch.randn(F, device=dev) * 0.1
dy = (torch.randn(M, D, device=dev) * 0.1).bfloat16()
w2 = (torch.randn(D, F, device=dev) * 0.03).bfloat16()   # fc2 weight [D, F]: dgrad B operand, MN-major
wq = (torch.randn(3 * D, D, device=dev) * 0.03).bfloat16()
bq = torch.randn(3 * D, device=dev) * 0.1
act, dact = ops.gemm_bf16(h, w1, bias=b1, epilogue=ops.EPI_BF16_GELU)
def t(fn, flops, name, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:28s} {ms:.4f} ms  {flops / ms / 1e9:7.1f} TFLOP/s")
t(lambda: ops.gemm_bf16(h, w1, bias=b1, epilogue=ops.EPI_BF16_GELU), 2.0 * M * D * F, "fc1+GELU (value+grad)")
t(lambda: ops.gemm_bf16(dy, w2, b_mn_major=True, aux=dact, epilogue=ops.EPI_BF16_DGELU), 2.0 * M * D * F, "fc2 dgrad x GELU'")
t(lambda: ops.gemm_bf16(h, wq, bias=bq, epilogue=ops.EPI_BF16), 2.0 * M * D * 3 * D, "QKV forward (plain bf16)")
t(lambda: ops.gemm_bf16(h, w1, bias=b1, epilogue=ops.EPI_BF16), 2.0 * M * D * F, "fc1 plain bf16 (no GELU)")
wo = (torch.randn(D, D, device=dev) * 0.03).bfloat16()
bo = torch.randn(D, device=dev) * 0.1
xres = torch.randn(M, D, device=dev)
w2f = (torch.randn(D, F, device=dev) * 0.03).bfloat16()  # fc2 forward: [N = D, K = F]
t(lambda: ops.gemm_bf16(h, wo, bias=bo, aux=xres, epilogue=ops.EPI_F32_RESID), 2.0 * M * D * D, "out-proj + residual (K=1024)")
t(lambda: ops.gemm_bf16(act, w2f, bias=bo, aux=xres, epilogue=ops.EPI_F32_RESID), 2.0 * M * D * F, "fc2 + residual (K=4096)")
