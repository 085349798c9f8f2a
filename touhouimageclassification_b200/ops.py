"""Per-kernel Python wrappers over the C-ABI (used by the parity tests and by the host-side glue).

Every function takes CUDA torch tensors, passes raw device pointers + sizes + the current stream to
``libtic_b200.so`` and returns torch tensors it allocated itself. torch is plumbing here (device
memory and streams); no torch operator does arithmetic on this path.
"""
from __future__ import annotations


import ctypes

import torch

from . import _lib
from ._lib import c_float, c_i64, c_int, c_void_p

EPI_BF16, EPI_BF16_GELU, EPI_F32_RESID, EPI_BF16_DGELU, EPI_F32, EPI_F32_ATOMIC, EPI_F32_POSEMB = range(7)


def _p(t):
    return c_void_p(0 if t is None else t.data_ptr())


def _s():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _guarded(fn):
    """Run ``fn`` with the device of its first CUDA tensor argument current: kernels, tensor maps and the stream handed to
    the C-ABI all belong to the CURRENT device, which need not be the tensors' device (a replica on cuda:1, a worker
    thread whose current device defaults to 0)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index == torch.cuda.current_device():
                    break
                with torch.cuda.device(a.device):
                    return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapper


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise ValueError("tic_b200 ops take CUDA tensors only (no CPU fallback)")


@_guarded
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


@_guarded
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


@_guarded
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


@_guarded
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


def _attention_bwd_scratch(B, N, H, device):
    """delta (+ the bf16 dQ partials of the fused long-sequence kernel for N > 256): tic_attention_bwd_scratch_floats."""
    lib = _lib.load()
    lib.tic_attention_bwd_scratch_floats.restype = ctypes.c_int64
    n = lib.tic_attention_bwd_scratch_floats(c_int(B), c_int(N), c_int(H))
    return torch.empty(int(n), device=device, dtype=torch.float32)


@_guarded
def attention_bwd(qkv, ctx, dctx, lse, B, N, H, scale=0.125, bias_grad=None, num_queries=None):
    """dqkv [B*N, 3D] bf16. ``bias_grad`` (fp32 [3D], optional) accumulates the column sums of dqkv (QKV bias gradient).
    ``num_queries`` = Nq: only the first Nq tokens of every image are queries (dq of the other rows is left zero)."""
    _require_cuda(qkv, ctx, dctx, lse)
    D = H * 64
    dqkv = torch.empty_like(qkv) if num_queries is None else torch.zeros_like(qkv)
    if num_queries is not None:
        delta = _attention_bwd_scratch(B, N, H, qkv.device)
        base, dbase = qkv.data_ptr(), dqkv.data_ptr()
        _lib.check(_lib.load().tic_attention_bwd_nq(c_void_p(base), c_void_p(base + 2 * D), c_void_p(base + 4 * D),
                                                    c_i64(3 * D), _p(ctx), c_i64(D), _p(dctx), c_i64(D), _p(lse), _p(delta),
                                                    c_void_p(dbase), c_void_p(dbase + 2 * D), c_void_p(dbase + 4 * D),
                                                    c_i64(3 * D), _p(bias_grad), c_int(B), c_int(N), c_int(num_queries),
                                                    c_int(H), c_int(64), c_float(scale), _s()))
        return dqkv
    delta = _attention_bwd_scratch(B, N, H, qkv.device)
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


@_guarded
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


class _SoftmaxXent(torch.autograd.Function):
    """``F.cross_entropy`` (mean reduction, integer or soft targets) as one ``tic_softmax_xent`` launch: the forward also
    produces d loss / d logits, which the backward hands to autograd scaled by the upstream gradient."""

    @staticmethod
    def forward(ctx, logits, target):
        loss, dl, _ = softmax_xent(logits.detach(), target, need_grad=True)
        ctx.save_for_backward(dl)
        ctx.in_dtype = logits.dtype
        return loss.view(())

    @staticmethod
    def backward(ctx, grad_out):
        (dl,) = ctx.saved_tensors
        return (dl * grad_out).to(ctx.in_dtype), None


def cross_entropy(logits, target):
    """Drop-in for ``F.cross_entropy(logits, target)`` as the reference calls it (finetune.py:61 integer labels,
    ntrain.py:48 soft MixUp / CutMix targets) on the engine's fused softmax-CE kernel, differentiable w.r.t. ``logits``."""
    return _SoftmaxXent.apply(logits, target)


@_guarded
def softmax_top1(logits, want_probs=False):
    """``torch.softmax(logits, 1)`` + ``torch.max(probabilities, 1)`` of the serving path in one launch.
    Returns (confidence fp32 [B], index int32 [B], probabilities fp32 [B, C] | None)."""
    _require_cuda(logits)
    B, C = logits.shape
    lg = logits if (logits.dtype == torch.float32 and logits.is_contiguous()) else logits.float().contiguous()
    conf = torch.empty(B, device=lg.device, dtype=torch.float32)
    idx = torch.empty(B, device=lg.device, dtype=torch.int32)
    probs = torch.empty((B, C), device=lg.device, dtype=torch.float32) if want_probs else None
    _lib.check(_lib.load().tic_softmax_top1(_p(lg), c_int(B), c_int(C), _p(conf), _p(idx), _p(probs), _s()))
    return conf, idx, probs


@_guarded
def adamw_step(p, g, m, v, shadow, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    _require_cuda(p, g, m, v, shadow)
    _lib.check(_lib.load().tic_adamw_step(_p(p), _p(g), _p(m), _p(v), _p(shadow), c_i64(p.numel()), c_float(lr),
                                          c_float(beta1), c_float(beta2), c_float(eps), c_float(weight_decay),
                                          c_int(step), c_float(grad_scale), _s()))


@_guarded
def patchify_f32(x):
    _require_cuda(x)
    B, C, S, _ = x.shape
    out = torch.empty((B * (S // 16) ** 2, 768), device=x.device, dtype=torch.bfloat16)
    _lib.check(_lib.load().tic_patchify_f32(_p(x.contiguous()), _p(out), c_int(B), c_int(S), _s()))
    return out


@_guarded
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


@_guarded
def colsum_bf16(dy):
    _require_cuda(dy)
    rows, cols = dy.shape
    out = torch.zeros(cols, device=dy.device, dtype=torch.float32)
    _lib.check(_lib.load().tic_colsum_bf16(_p(dy), c_i64(dy.stride(0)), c_int(rows), c_int(cols), _p(out), _s()))
    return out
