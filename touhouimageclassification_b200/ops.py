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
