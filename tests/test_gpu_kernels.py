"""Per-kernel parity on the B200, through the C-ABI, against plain fp32 torch restatements of each op
(floating-point kernels: tolerances are written in each assert; bf16 outputs carry ~2^-9 relative rounding)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

dev = "cuda"


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1024, 1024, 1024), (591, 768, 768), (1000, 128, 200), (50432, 1024, 256)])
def test_gemm_forward_bias(M, N, K):
    from touhouimageclassification_b200 import ops
    a = torch.randn(M, K, device=dev).bfloat16()
    b = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    out = ops.gemm_bf16(a, b, bias=bias, epilogue=ops.EPI_F32)
    ref = a.float() @ b.float().t() + bias
    assert rel(out, ref) < 2e-5          # fp32 accumulate, fp32 out
    out16 = ops.gemm_bf16(a, b, bias=bias)
    assert rel(out16, ref) < 3e-3        # bf16 output rounding


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1024, 1024, 3072), (300, 1024, 520)])
def test_gemm_dgrad_mn_major_b(M, N, K):
    from touhouimageclassification_b200 import ops
    a = torch.randn(M, K, device=dev).bfloat16()
    b = torch.randn(K, N, device=dev).bfloat16()
    out = ops.gemm_bf16(a, b, b_mn_major=True, epilogue=ops.EPI_F32)
    assert rel(out, a.float() @ b.float()) < 2e-5


@pytest.mark.parametrize("M,N,K,splits", [(128, 256, 64, 1), (1024, 768, 4096, 1), (1024, 768, 12608, 7), (3072, 1024, 1970, 3)])
def test_gemm_wgrad_splitk(M, N, K, splits):
    from touhouimageclassification_b200 import ops
    a = torch.randn(K, M, device=dev).bfloat16()
    b = torch.randn(K, N, device=dev).bfloat16()
    ref = a.float().t() @ b.float()
    out = torch.zeros(M, N, device=dev)
    ops.gemm_bf16(a, b, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32_ATOMIC, out=out, splits=splits)
    assert rel(out, ref) < 2e-5
    # linearity: accumulating the same product twice doubles the result (size-independent property)
    ops.gemm_bf16(a, b, a_mn_major=True, b_mn_major=True, epilogue=ops.EPI_F32_ATOMIC, out=out, splits=splits)
    assert rel(out, 2 * ref) < 2e-5


def test_gemm_fused_epilogues():
    from touhouimageclassification_b200 import ops
    M, N, K = 512, 1024, 256
    a = (torch.randn(M, K, device=dev) * 0.3).bfloat16()
    b = (torch.randn(N, K, device=dev) * 0.1).bfloat16()
    bias = torch.randn(N, device=dev) * 0.1
    act, dact = ops.gemm_bf16(a, b, bias=bias, epilogue=ops.EPI_BF16_GELU)
    pre_ref = (a.float() @ b.float().t() + bias).bfloat16()
    xr = pre_ref.float().requires_grad_(True)
    F.gelu(xr).sum().backward()
    # GELU of the bf16-rounded pre-activation (Appendix B); the second output is GELU'(pre), all the backward needs
    assert rel(act, F.gelu(pre_ref.float()).bfloat16()) < 1e-3
    assert rel(dact, xr.grad.bfloat16()) < 1e-3
    assert (act.float() - F.gelu(pre_ref.float())).abs().max() < 2e-2 and (dact.float() - xr.grad).abs().max() < 1e-2
    res = torch.randn(M, N, device=dev)
    out = ops.gemm_bf16(a, b, bias=bias, aux=res, epilogue=ops.EPI_F32_RESID)
    assert rel(out, pre_ref.float() + res) < 1e-3
    # dgrad * saved GELU': together with the forward epilogue this is autograd's gelu_backward
    bt = (torch.randn(K, N, device=dev) * 0.1).bfloat16()
    pre2 = torch.randn(M, N, device=dev).bfloat16()
    x = pre2.float().requires_grad_(True)
    F.gelu(x).sum().backward()
    dgelu = x.grad.bfloat16()
    out = ops.gemm_bf16(a, bt, b_mn_major=True, aux=dgelu, epilogue=ops.EPI_BF16_DGELU)
    g = (a.float() @ bt.float()).bfloat16().float()
    assert rel(out, g * dgelu.float()) < 3e-3
    x = pre2.float().requires_grad_(True)
    F.gelu(x).backward(g)
    assert rel(out, x.grad) < 6e-3
    pre2 = dgelu
    # fused bias gradient: column sums of the bf16 output, accumulated
    cs = torch.zeros(N, device=dev)
    _lib_out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    from touhouimageclassification_b200 import _lib as L
    import ctypes
    L.check(L.load().tic_gemm_bf16_colsum(ctypes.c_void_p(a.data_ptr()), ctypes.c_int64(K), 0, ctypes.c_void_p(bt.data_ptr()),
                                          ctypes.c_int64(N), 1, M, N, K, ops.EPI_BF16_DGELU, ctypes.c_void_p(_lib_out.data_ptr()),
                                          ctypes.c_int64(N), ctypes.c_void_p(pre2.data_ptr()), ctypes.c_int64(N),
                                          ctypes.c_void_p(cs.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert torch.equal(_lib_out, out)
    assert rel(cs, out.float().sum(0)) < 1e-4


@pytest.mark.parametrize("M,N,K", [(197, 1024, 1024), (197, 3072, 1024), (197, 1024, 4096), (394, 4096, 1024), (1576, 1024, 4096),
                                   (1576, 4096, 1024), (8, 1024, 1024), (130, 136, 72), (20000, 1024, 256)])
def test_gemm_forward_epilogues_small_and_wide_tiles(M, N, K):
    """The forward epilogues pick between two tile configurations by problem size (CTA pairs with 256 x 256 tiles, or
    single CTAs with 128 x 128 tiles when the wide tiles would leave most of the machine idle: small-batch inference).
    Shapes on both sides of the switch, ragged edges included, against fp32 matmuls of the same bf16 operands."""
    from touhouimageclassification_b200 import ops
    g = torch.Generator(device=dev).manual_seed(M + N + K)
    a = (torch.randn(M, K, device=dev, generator=g) * 0.3).bfloat16()
    b = (torch.randn(N, K, device=dev, generator=g) * 0.1).bfloat16()
    bias = torch.randn(N, device=dev, generator=g) * 0.1
    pre = a.float() @ b.float().t() + bias
    out = ops.gemm_bf16(a, b, bias=bias, epilogue=ops.EPI_BF16)
    assert rel(out, pre) < 3e-3 and torch.equal(out, pre.bfloat16()) or rel(out.float(), pre.bfloat16().float()) < 1e-3
    act, dact = ops.gemm_bf16(a, b, bias=bias, epilogue=ops.EPI_BF16_GELU)
    assert rel(act, F.gelu(pre.bfloat16().float())) < 4e-3
    res = torch.randn(M, N, device=dev, generator=g)
    out = ops.gemm_bf16(a, b, bias=bias, aux=res, epilogue=ops.EPI_F32_RESID)
    assert rel(out, pre.bfloat16().float() + res) < 1e-3
    # run to run, bit for bit
    assert torch.equal(ops.gemm_bf16(a, b, bias=bias, aux=res, epilogue=ops.EPI_F32_RESID), out)
    act2, _ = ops.gemm_bf16(a, b, bias=bias, epilogue=ops.EPI_BF16_GELU)
    assert torch.equal(act2, act)


def test_gemm_rejects_bad_arguments():
    from touhouimageclassification_b200 import ops, _lib
    a = torch.randn(128, 64, device=dev).bfloat16()
    b = torch.randn(100, 64, device=dev).bfloat16()  # N = 100 is not a multiple of 8
    with pytest.raises(_lib.TicError):
        ops.gemm_bf16(a, b)
    with pytest.raises(ValueError):
        ops.gemm_bf16(a.cpu(), b.cpu())


@pytest.mark.parametrize("D", [128, 768, 1024])
def test_layernorm_fwd_bwd(D):
    from touhouimageclassification_b200 import ops
    rows = 1000
    x = torch.randn(rows, D, device=dev) * 2 + 0.5
    g = torch.randn(D, device=dev)
    b = torch.randn(D, device=dev)
    y, yf, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12, out_f32=True)
    xr, gr, br = (t.clone().requires_grad_(True) for t in (x, g, b))
    ref = F.layer_norm(xr, (D,), gr, br, 1e-12)
    assert rel(yf, ref) < 1e-6 and rel(y, ref) < 3e-3
    dy = torch.randn(rows, D, device=dev).bfloat16()
    dres = torch.randn(rows, D, device=dev)
    ref.backward(dy.float())
    dx, dxb, dg, db, dxsum = ops.layernorm_bwd(dy, x, mean, rstd, g, dres)
    assert rel(dx - dres, xr.grad) < 1e-5
    assert rel(dxb, dx) < 3e-3
    assert rel(dxsum, dxb.float().sum(0)) < 1e-5      # fused bias-gradient column sums of the bf16 dx
    assert rel(dg, gr.grad) < 1e-5 and rel(db, br.grad) < 1e-5


def test_layernorm_eps_is_1e_12_not_fixed():
    """A constant row gives rsqrt(0 + 1e-12) = 1e6: the output equals beta, rstd is 1e6."""
    from touhouimageclassification_b200 import ops
    x = torch.full((8, 128), 3.0, device=dev)
    g = torch.ones(128, device=dev)
    b = torch.full((128,), 0.25, device=dev)
    _, yf, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12, out_f32=True)
    assert torch.allclose(rstd, torch.full_like(rstd, 1e6), rtol=1e-3)
    assert torch.allclose(yf, torch.full_like(yf, 0.25))


@pytest.mark.parametrize("B,N,H", [(2, 197, 12), (2, 577, 4), (2, 64, 2), (1, 1, 1), (2, 130, 3), (1, 5, 2), (3, 256, 2),
                                   (2, 128, 1), (2, 129, 2), (2, 192, 2), (2, 193, 1), (2, 65, 2), (1, 257, 2), (64, 197, 16),
                                   (2, 225, 2), (3, 300, 2), (2, 385, 3), (1, 512, 2), (2, 513, 1), (1, 640, 2), (1, 1000, 1),
                                   (9, 577, 16)])
def test_attention_fwd_bwd(B, N, H):
    from touhouimageclassification_b200 import ops
    D = H * 64
    qkv = torch.randn(B * N, 3 * D, device=dev).bfloat16()
    ctx, lse = ops.attention_fwd(qkv, B, N, H)
    q, k, v = [t.view(B, N, H, 64).transpose(1, 2).float().requires_grad_(True) for t in qkv.float().split(D, dim=1)]
    s = q @ k.transpose(-1, -2) * 0.125
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * N, D)
    assert rel(ctx, ref) < 4e-3 or N == 1
    assert rel(lse, torch.logsumexp(s, -1)) < 1e-5 or N == 1
    dctx = torch.randn(B * N, D, device=dev).bfloat16()
    ref.backward(dctx.float())
    dqkv = ops.attention_bwd(qkv, ctx, dctx, lse, B, N, H)
    assert torch.isfinite(dqkv.float()).all()
    dq, dk, dv = dqkv.float().split(D, dim=1)
    for ours, t in ((dq, q), (dk, k), (dv, v)):
        r = t.grad.transpose(1, 2).reshape(B * N, D)
        if N == 1:
            assert (ours - r).abs().max() < 2e-2
        else:
            assert rel(ours, r) < 6e-3
    # the fused QKV bias gradient accumulates the column sums of the bf16 dqkv it wrote, and dqkv itself is unchanged
    bg = torch.full((3 * D,), 0.5, device=dev)
    dqkv2 = ops.attention_bwd(qkv, ctx, dctx, lse, B, N, H, bias_grad=bg)
    assert torch.equal(dqkv2, dqkv)
    want = dqkv.float().sum(0) + 0.5
    assert (bg - want).abs().max() <= 1e-4 * max(1.0, want.abs().max().item()) + 1e-5 * B * N


@pytest.mark.parametrize("B,N,H", [(2, 577, 2), (1, 300, 1), (2, 197, 2)])
def test_attention_large_score_range(B, N, H):
    """Scores spread over hundreds of units with the row maximum moving from key block to key block: exercises the online
    softmax of the long-sequence forward (lazy rescaling of the running output) and the saved logsumexp in the backward."""
    from touhouimageclassification_b200 import ops
    D = H * 64
    qkv = torch.randn(B * N, 3 * D, device=dev)
    qkv[:, :D] *= 6.0                                                  # |q.k| / 8 reaches ~ +-150
    qkv.view(B, N, 3 * D)[:, :, D:2 * D] *= torch.linspace(0.2, 4.0, N, device=dev).view(1, N, 1)  # later keys score higher
    qkv = qkv.bfloat16()
    ctx, lse = ops.attention_fwd(qkv, B, N, H)
    q, k, v = [t.view(B, N, H, 64).transpose(1, 2).float().requires_grad_(True) for t in qkv.float().split(D, dim=1)]
    s = q @ k.transpose(-1, -2) * 0.125
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * N, D)
    assert torch.isfinite(ctx.float()).all() and rel(ctx, ref) < 5e-3
    assert (lse - torch.logsumexp(s, -1)).abs().max() < 2e-3 * torch.logsumexp(s, -1).abs().max()
    dctx = torch.randn(B * N, D, device=dev).bfloat16()
    ref.backward(dctx.float())
    dqkv = ops.attention_bwd(qkv, ctx, dctx, lse, B, N, H)
    assert torch.isfinite(dqkv.float()).all()
    for ours, t in zip(dqkv.float().split(D, dim=1), (q, k, v)):
        assert rel(ours, t.grad.transpose(1, 2).reshape(B * N, D)) < 1.5e-2


def test_softmax_xent_hard_and_soft():
    from touhouimageclassification_b200 import ops
    B, C = 37, 120
    logits = torch.randn(B, C, device=dev) * 3
    y = torch.randint(0, C, (B,), device=dev)
    lr_ = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(lr_, y)
    ref.backward()
    loss, dl, correct = ops.softmax_xent(logits, y)
    assert abs(loss.item() - ref.item()) < 1e-5 and rel(dl, lr_.grad) < 1e-5
    assert correct.item() == (logits.argmax(1) == y).sum().item()
    soft = torch.softmax(torch.randn(B, C, device=dev), 1)
    lr_ = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(lr_, soft)
    ref.backward()
    loss, dl, _ = ops.softmax_xent(logits, soft)
    assert abs(loss.item() - ref.item()) < 1e-5 and rel(dl, lr_.grad) < 1e-5
    # gradients of a cross-entropy sum to zero per row (property, any size)
    assert dl.sum(1).abs().max() < 1e-6


def test_adamw_matches_torch_and_writes_shadow():
    from touhouimageclassification_b200 import ops
    n = 1 << 20
    p = torch.randn(n, device=dev)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, weight_decay=0.01)
    m = torch.zeros(n, device=dev)
    v = torch.zeros(n, device=dev)
    sh = torch.empty(n, device=dev, dtype=torch.bfloat16)
    for step in range(1, 5):
        g = torch.randn(n, device=dev) * 0.1
        pr.grad = g.clone()
        opt.step()
        ops.adamw_step(p, g, m, v, sh, 1e-3, 0.9, 0.999, 1e-8, 0.01, step)
    assert (p - pr).abs().max() < 1e-6
    assert torch.equal(sh, p.bfloat16())


def test_patchify_is_exact_and_colsum():
    from touhouimageclassification_b200 import ops
    x = torch.randn(3, 3, 224, 224, device=dev)
    ref = x.unfold(2, 16, 16).unfold(3, 16, 16).permute(0, 2, 3, 1, 4, 5).reshape(3 * 196, 768).bfloat16()
    assert torch.equal(ops.patchify_f32(x), ref)       # pure data movement + one rounding: bit-exact
    dy = torch.randn(5000, 3072, device=dev).bfloat16()
    assert rel(ops.colsum_bf16(dy), dy.float().sum(0)) < 1e-5


@pytest.mark.parametrize("B,N,H,Nq", [(3, 197, 4, 1), (2, 197, 2, 70), (2, 64, 2, 1), (2, 130, 2, 129), (4, 197, 16, 1),
                                      (2, 577, 2, 1), (1, 300, 3, 1), (2, 240, 2, 1), (2, 577, 2, 130), (1, 640, 1, 300),
                                      (2, 385, 2, 64), (2, 300, 2, 299)])
def test_attention_query_subset(B, N, H, Nq):
    """Only the first Nq tokens of every image are queries (Nq = 1: the CLS-only last layer): forward rows, logsumexp and
    all three gradients equal the full computation with the other queries' upstream gradient set to zero."""
    from touhouimageclassification_b200 import ops
    D = H * 64
    qkv = torch.randn(B * N, 3 * D, device=dev).bfloat16()
    ctx_full, lse_full = ops.attention_fwd(qkv, B, N, H)
    ctx, lse = ops.attention_fwd(qkv, B, N, H, num_queries=Nq)
    cf, c = ctx_full.view(B, N, D), ctx.view(B, N, D)
    assert c[:, Nq:].abs().max() == 0
    if N <= 224:   # one key block: a row's arithmetic does not depend on how many rows are queries
        assert torch.equal(c[:, :Nq], cf[:, :Nq])
        assert torch.equal(lse, lse_full[:, :, :Nq].contiguous())
    else:          # long sequences: a tile without a partner splits its key blocks over both softmax groups and merges
        assert rel(c[:, :Nq], cf[:, :Nq]) < 3e-3
        assert (lse - lse_full[:, :, :Nq]).abs().max() < 1e-4
    dctx = torch.randn(B * N, D, device=dev).bfloat16()
    dctx.view(B, N, D)[:, Nq:] = 0
    ref = ops.attention_bwd(qkv, ctx_full, dctx, lse_full, B, N, H)
    bg = torch.zeros(3 * D, device=dev)
    got = ops.attention_bwd(qkv, ctx, dctx, lse, B, N, H, bias_grad=bg, num_queries=Nq)
    assert torch.isfinite(got.float()).all()
    assert rel(got, ref) < 2e-3 and got.view(B, N, 3 * D)[:, Nq:, :D].abs().max() == 0
    assert (bg - got.float().sum(0)).abs().max() <= 1e-4 * max(1.0, got.float().sum(0).abs().max().item()) + 1e-5 * B * N
