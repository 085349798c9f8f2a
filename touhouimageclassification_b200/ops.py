"""Per-kernel Python wrappers over the C-ABI (used by the parity tests and by the host-side glue).

Every function takes CUDA torch tensors, passes raw device pointers + sizes + the current stream to
``libtic_b200.so`` and returns torch tensors it allocated itself. torch is plumbing here (device
memory and streams); no torch operator does arithmetic on this path.
"""
from __future__ import annotations


import torch

from . import _lib
from ._lib import c_float, c_i64, c_int, c_void_p

EPI_BF16, EPI_BF16_GELU, EPI_F32_RESID, EPI_BF16_DGELU, EPI_F32, EPI_F32_ATOMIC, EPI_F32_POSEMB = range(7)


def _p(t):
    return c_void_p(0 if t is None else t.data_ptr())


def _s():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise ValueError("tic_b200 ops take CUDA tensors only (no CPU fallback)")


def gemm_bf16(a, b, *, a_mn_major=False, b_mn_major=False, epilogue=EPI_BF16, bias=None, aux=None, aux_int=0,
              out=None, out2=None, splits=1):
    """``D = A @ B^T`` on tcgen05. ``a``: [M,K] (or [K,M] when ``a_mn_major``); ``b``: [N,K] (or [K,N])."""
    _require_cuda(a, b, bias, aux, out, out2)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.stride(-1) == 1 and b.stride(-1) == 1
    if a_mn_major:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn_major:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, (a.shape, b.shape)
    f32_out = epilogue in (EPI_F32_RESID, EPI_F32, EPI_F32_ATOMIC, EPI_F32_POSEMB)
    if out is None:
        assert epilogue not in (EPI_F32_ATOMIC, EPI_F32_POSEMB)
        out = torch.empty((M, N), device=a.device, dtype=torch.float32 if f32_out else torch.bfloat16)
    if epilogue == EPI_BF16_GELU and out2 is None:
        out2 = torch.empty((M, N), device=a.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().tic_gemm_bf16(
        _p(a), c_i64(a.stride(0)), c_int(int(a_mn_major)), _p(b), c_i64(b.stride(0)), c_int(int(b_mn_major)),
        c_int(M), c_int(N), c_int(K), c_int(epilogue), _p(out), c_i64(out.stride(0)),
        _p(out2), c_i64(0 if out2 is None else out2.stride(0)), _p(bias), _p(aux),
        c_i64(0 if aux is None else aux.stride(0)), c_int(aux_int), c_int(splits), _s()))
    if epilogue == EPI_BF16_GELU:
        return out, out2
    return out


def layernorm_fwd(x, gamma, beta, eps, out_bf16=True, out_f32=False):
    _require_cuda(x, gamma, beta)
    rows, D = x.shape
    y = torch.empty((rows, D), device=x.device, dtype=torch.bfloat16) if out_bf16 else None
    yf = torch.empty((rows, D), device=x.device, dtype=torch.float32) if out_f32 else None
    mean = torch.empty(rows, device=x.device, dtype=torch.float32)
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().tic_layernorm_fwd(_p(x), c_i64(x.stride(0)), _p(gamma), _p(beta), c_float(eps), c_int(rows),
                                             c_int(D), _p(y), c_i64(D), _p(yf), c_i64(D), _p(mean), _p(rstd), _s()))
    return y, yf, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dres=None):
    _require_cuda(dy, x, mean, rstd, gamma, dres)
    rows, D = x.shape
    dx = torch.empty((rows, D), device=x.device, dtype=torch.float32)
    dxb = torch.empty((rows, D), device=x.device, dtype=torch.bfloat16)
    dgamma = torch.zeros(D, device=x.device, dtype=torch.float32)
    dbeta = torch.zeros(D, device=x.device, dtype=torch.float32)
    dxsum = torch.zeros(D, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().tic_layernorm_bwd(_p(dy), c_i64(dy.stride(0)), _p(x), c_i64(x.stride(0)), _p(mean), _p(rstd),
                                             _p(gamma), _p(dres), c_i64(0 if dres is None else dres.stride(0)),
                                             c_int(rows), c_int(D), _p(dx), c_i64(D), _p(dxb), c_i64(D), _p(dgamma),
                                             _p(dbeta), _p(dxsum), _s()))
    return dx, dxb, dgamma, dbeta, dxsum


def attention_fwd(qkv, B, N, H, scale=0.125, need_lse=True, num_queries=None):
    """qkv: bf16 [B*N, 3*H*64] (q | k | v column blocks). Returns ctx bf16 [B*N, H*64], lse fp32 [B,H,N].
    ``num_queries`` = Nq: only the first Nq tokens of every image are queries (other ctx rows are left zero, lse [B,H,Nq])."""
    _require_cuda(qkv)
    D = H * 64
    assert qkv.shape == (B * N, 3 * D) and qkv.dtype == torch.bfloat16 and qkv.is_contiguous()
    Nq = N if num_queries is None else num_queries
    ctx = (torch.empty if Nq == N else torch.zeros)((B * N, D), device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty((B, H, Nq), device=qkv.device, dtype=torch.float32) if need_lse else None
    base = qkv.data_ptr()
    if num_queries is None:
        _lib.check(_lib.load().tic_attention_fwd(c_void_p(base), c_void_p(base + 2 * D), c_void_p(base + 4 * D), c_i64(3 * D),
                                                 _p(ctx), c_i64(D), _p(lse), c_int(B), c_int(N), c_int(H), c_int(64),
                                                 c_float(scale), _s()))
    else:
        _lib.check(_lib.load().tic_attention_fwd_nq(c_void_p(base), c_void_p(base + 2 * D), c_void_p(base + 4 * D),
                                                    c_i64(3 * D), _p(ctx), c_i64(D), _p(lse), c_int(B), c_int(N), c_int(Nq),
                                                    c_int(H), c_int(64), c_float(scale), _s()))
    return ctx, lse


def attention_bwd(qkv, ctx, dctx, lse, B, N, H, scale=0.125, bias_grad=None, num_queries=None):
    """dqkv [B*N, 3D] bf16. ``bias_grad`` (fp32 [3D], optional) accumulates the column sums of dqkv (QKV bias gradient).
    ``num_queries`` = Nq: only the first Nq tokens of every image are queries (dq of the other rows is left zero)."""
    _require_cuda(qkv, ctx, dctx, lse)
    D = H * 64
    dqkv = torch.empty_like(qkv) if num_queries is None else torch.zeros_like(qkv)
    if num_queries is not None:
        delta = torch.empty((B, H, N), device=qkv.device, dtype=torch.float32)
        base, dbase = qkv.data_ptr(), dqkv.data_ptr()
        _lib.check(_lib.load().tic_attention_bwd_nq(c_void_p(base), c_void_p(base + 2 * D), c_void_p(base + 4 * D),
                                                    c_i64(3 * D), _p(ctx), c_i64(D), _p(dctx), c_i64(D), _p(lse), _p(delta),
                                                    c_void_p(dbase), c_void_p(dbase + 2 * D), c_void_p(dbase + 4 * D),
                                                    c_i64(3 * D), _p(bias_grad), c_int(B), c_int(N), c_int(num_queries),
                                                    c_int(H), c_int(64), c_float(scale), _s()))
        return dqkv
    delta = torch.empty((B, H, N), device=qkv.device, dtype=torch.float32)
    base, dbase = qkv.data_ptr(), dqkv.data_ptr()
    if bias_grad is None:
        _lib.check(_lib.load().tic_attention_bwd(c_void_p(base), c_void_p(base + 2 * D), c_void_p(base + 4 * D), c_i64(3 * D),
                                                 _p(ctx), c_i64(D), _p(dctx), c_i64(D), _p(lse), _p(delta), c_void_p(dbase),
                                                 c_void_p(dbase + 2 * D), c_void_p(dbase + 4 * D), c_i64(3 * D), c_int(B),
                                                 c_int(N), c_int(H), c_int(64), c_float(scale), _s()))
    else:
        assert bias_grad.dtype == torch.float32 and bias_grad.numel() == 3 * D and bias_grad.is_cuda
        _lib.check(_lib.load().tic_attention_bwd_bias(c_void_p(base), c_void_p(base + 2 * D), c_void_p(base + 4 * D),
                                                      c_i64(3 * D), _p(ctx), c_i64(D), _p(dctx), c_i64(D), _p(lse), _p(delta),
                                                      c_void_p(dbase), c_void_p(dbase + 2 * D), c_void_p(dbase + 4 * D),
                                                      c_i64(3 * D), _p(bias_grad), c_int(B), c_int(N), c_int(H), c_int(64),
                                                      c_float(scale), _s()))
    return dqkv


def softmax_xent(logits, target, grad_scale=None, round_grad=False, need_grad=True):
    """target: int64 [B] (hard) or float [B,C] (soft). Returns (loss[1], dlogits|None, correct[1] int32)."""
    _require_cuda(logits, target)
    B, C = logits.shape
    logits = logits.float().contiguous()
    hard = target.dtype in (torch.int64, torch.int32)
    tgt = target.long().contiguous() if hard else target.float().contiguous()
    loss = torch.empty(1, device=logits.device, dtype=torch.float32)
    correct = torch.zeros(1, device=logits.device, dtype=torch.int32)
    dl = torch.empty_like(logits) if need_grad else None
    gs = 1.0 / B if grad_scale is None else grad_scale
    _lib.check(_lib.load().tic_softmax_xent(_p(logits), _p(tgt if hard else None), _p(None if hard else tgt), c_int(B),
                                            c_int(C), c_float(gs), c_int(int(round_grad)), _p(loss), _p(dl), _p(correct),
                                            _s()))
    return loss, dl, correct


def adamw_step(p, g, m, v, shadow, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    _require_cuda(p, g, m, v, shadow)
    _lib.check(_lib.load().tic_adamw_step(_p(p), _p(g), _p(m), _p(v), _p(shadow), c_i64(p.numel()), c_float(lr),
                                          c_float(beta1), c_float(beta2), c_float(eps), c_float(weight_decay),
                                          c_int(step), c_float(grad_scale), _s()))


def patchify_f32(x):
    _require_cuda(x)
    B, C, S, _ = x.shape
    out = torch.empty((B * (S // 16) ** 2, 768), device=x.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().tic_patchify_f32(_p(x.contiguous()), _p(out), c_int(B), c_int(S), _s()))
    return out


def mix_batch(x, y, num_classes, mode, lam, box, lam_label, want_pixels=True, want_patches=False):
    """CutMix (mode 2) / MixUp (mode 1) of a device batch against its roll(1, 0), fused with the patchify.
    Returns (mixed fp32 pixels | None, soft labels fp32 [B, C], bf16 patch rows | None)."""
    _require_cuda(x, y)
    B, _, S, _ = x.shape
    xs = x.detach()
    if xs.dtype != torch.float32 or not xs.is_contiguous():
        xs = xs.float().contiguous()
    mixed = torch.empty_like(xs) if want_pixels else None
    patches = torch.empty((B * (S // 16) ** 2, 768), device=x.device, dtype=torch.bfloat16) if want_patches else None
    x1, y1, x2, y2 = box
    lib = _lib.load()
    _lib.check(lib.tic_mix_patchify_f32(_p(xs), _p(mixed), _p(patches), c_int(B), c_int(S), c_int(mode), c_float(lam),
                                        c_float(1.0 - lam), c_int(x1), c_int(y1), c_int(x2), c_int(y2), _s()))
    soft = torch.empty((B, num_classes), device=x.device, dtype=torch.float32)
    yl = y.long().contiguous()
    _lib.check(lib.tic_mix_targets(_p(yl), c_int(B), c_int(num_classes), c_float(lam_label), c_float(1.0 - lam_label),
                                   _p(soft), _s()))
    return mixed, soft, patches


def colsum_bf16(dy):
    _require_cuda(dy)
    rows, cols = dy.shape
    out = torch.zeros(cols, device=dy.device, dtype=torch.float32)
    _lib.check(_lib.load().tic_colsum_bf16(_p(dy), c_i64(dy.stride(0)), c_int(rows), c_int(cols), _p(out), _s()))
    return out
