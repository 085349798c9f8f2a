"""GPU probe: tcgen05 GEMM variants against torch fp32 matmul (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from touhouimageclassification_b200 import ops

torch.manual_seed(0)
dev = "cuda"

def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()

def run(name, fn):
    try:
        r = fn()
        torch.cuda.synchronize()
        print(f"[{name}] {r}", flush=True)
    except Exception as e:
        print(f"[{name}] EXC {type(e).__name__}: {e}", flush=True)
        raise

def case_tn(M, N, K):
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    out = ops.gemm_bf16(a, b, bias=bias)
    ref = a.float() @ b.float().t() + bias
    return f"M{M} N{N} K{K} rel={rel(out, ref):.3e}"

def case_nn(M, N, K):  # dgrad: A [M,K] K-major, B stored [K,N]
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(K, N, device=dev).bfloat16()
    out = ops.gemm_bf16(a, b, b_mn_major=True, epilogue=ops.EPI_F32)
    ref = a.float() @ b.float()
    return f"M{M} N{N} K{K} rel={rel(out, ref):.3e}"

def case_nt(M, N, K, splits):  # wgrad: A stored [K,M], B stored [K,N]
    a = torch.randn(K, M, device=dev).bfloat16(); b = torch.randn(K, N, device=dev).bfloat16()
    ref = a.float().t() @ b.float()
    if splits == 1:
        out = ops.gemm_bf16(a, b, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32)
    else:
        out = torch.zeros(M, N, device=dev)
        ops.gemm_bf16(a, b, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32_ATOMIC, out=out, splits=splits)
    return f"M{M} N{N} K{K} splits{splits} rel={rel(out, ref):.3e}"

def case_gelu(M, N, K):
    a = (torch.randn(M, K, device=dev) * 0.3).bfloat16(); b = (torch.randn(N, K, device=dev) * 0.1).bfloat16()
    bias = torch.randn(N, device=dev) * 0.1
    out, dact = ops.gemm_bf16(a, b, bias=bias, epilogue=ops.EPI_BF16_GELU)
    pre_ref = (a.float() @ b.float().t() + bias).bfloat16()
    act_ref = torch.nn.functional.gelu(pre_ref.float()).bfloat16()
    x = pre_ref.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    return f"gelu' rel={rel(dact, x.grad):.3e} act rel={rel(out, act_ref):.3e}"

def case_resid(M, N, K):
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev)
    out = ops.gemm_bf16(a, b, bias=bias, aux=res, epilogue=ops.EPI_F32_RESID)
    ref = (a.float() @ b.float().t() + bias).bfloat16().float() + res
    return f"rel={rel(out, ref):.3e}"

def case_dgelu(M, N, K):
    a = torch.randn(M, K, device=dev).bfloat16(); b = (torch.randn(K, N, device=dev) * 0.1).bfloat16()
    pre = torch.randn(M, N, device=dev).bfloat16()
    x = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    out = ops.gemm_bf16(a, b, b_mn_major=True, aux=x.grad.bfloat16(), epilogue=ops.EPI_BF16_DGELU)
    g = (a.float() @ b.float()).bfloat16().float()
    x = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(x).backward(g)
    return f"rel={rel(out, x.grad):.3e}"

def bench(M, N, K, a_mn=False, b_mn=False, epi=ops.EPI_BF16, splits=1, iters=20):
    a = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    f32 = epi in (ops.EPI_F32, ops.EPI_F32_ATOMIC)
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    kw = {}
    if epi == ops.EPI_BF16_DGELU:
        kw["aux"] = torch.rand(M, N, device=dev).bfloat16()
    if epi == ops.EPI_BF16_GELU:
        kw["out2"] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        kw["bias"] = torch.randn(N, device=dev)
    for _ in range(3):
        ops.gemm_bf16(a, b, a_mn_major=a_mn, b_mn_major=b_mn, epilogue=epi, out=out, splits=splits, **kw)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm_bf16(a, b, a_mn_major=a_mn, b_mn_major=b_mn, epilogue=epi, out=out, splits=splits, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return f"M{M} N{N} K{K} amn{int(a_mn)} bmn{int(b_mn)} epi{epi} splits{splits}: {ms:.3f} ms {2*M*N*K/ms/1e9:.1f} TFLOP/s"

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    run("tn small", lambda: case_tn(128, 256, 64))
    run("tn 1tile", lambda: case_tn(128, 256, 512))
    run("tn multi", lambda: case_tn(1024, 1024, 1024))
    run("tn ragged", lambda: case_tn(197 * 3, 768, 768))
    run("tn ragged2", lambda: case_tn(1000, 120 + 8, 200))
    run("nn small", lambda: case_nn(128, 256, 64))
    run("nn multi", lambda: case_nn(1024, 1024, 3072))
    run("nt small", lambda: case_nt(128, 256, 64, 1))
    run("nt multi", lambda: case_nt(1024, 768, 4096, 1))
    run("nt split", lambda: case_nt(1024, 768, 197 * 64, 7))
    run("gelu", lambda: case_gelu(512, 1024, 256))
    run("resid", lambda: case_resid(512, 1024, 256))
    run("dgelu", lambda: case_dgelu(512, 1024, 256))
    M = 256 * 197
    run("bench qkv", lambda: bench(M, 3072, 1024))
    run("bench fc1", lambda: bench(M, 4096, 1024, epi=ops.EPI_BF16_GELU))
    run("bench fc2", lambda: bench(M, 1024, 4096))
    run("bench dgrad fc1", lambda: bench(M, 1024, 4096, b_mn=True))
    run("bench dgrad fc2", lambda: bench(M, 4096, 1024, b_mn=True))
    run("bench dgrad fc2 dgelu", lambda: bench(M, 4096, 1024, b_mn=True, epi=ops.EPI_BF16_DGELU))
    for sp in (4, 8, 15):
        run("bench wgrad fc1", lambda: bench(4096, 1024, M, a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC, splits=sp))
    for sp in (4, 8):
        run("bench wgrad fc2", lambda: bench(1024, 4096, M, a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC, splits=sp))
    for sp in (3, 6, 9):
        run("bench wgrad qkv", lambda: bench(3072, 1024, M, a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC, splits=sp))
    for sp in (9, 18):
        run("bench wgrad o", lambda: bench(1024, 1024, M, a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC, splits=sp))
    a = torch.randn(8192, 8192, device=dev).bfloat16(); b = torch.randn(8192, 8192, device=dev).bfloat16()
    for _ in range(3): a @ b.t()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(10): a @ b.t()
    torch.cuda.synchronize(); dt = (time.time() - t0) / 10
    print(f"cublas 8192^3: {2*8192**3/dt/1e12:.1f} TFLOP/s", flush=True)
    run("bench 8192", lambda: bench(8192, 8192, 8192))
