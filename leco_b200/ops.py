"""Thin torch-tensor front end over the C ABI (device pointers + sizes only cross it).

Every function enqueues hand-written sm_100a kernels on torch's current CUDA stream.
There is no PyTorch/CPU fallback here: a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import capi
from .capi import GemmArgs

BF16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise capi.LecoError("leco_b200 ops need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def _req_bf16(t: torch.Tensor, name: str):
    if t.dtype != BF16:
        raise capi.LecoError(f"{name} must be bfloat16, got {t.dtype}")


def gemm(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, *,
         bias: Optional[torch.Tensor] = None,
         rowbias: Optional[torch.Tensor] = None, rows_per_group: int = 1,
         residual: Optional[torch.Tensor] = None,
         lora_t: Optional[torch.Tensor] = None, lora_up: Optional[torch.Tensor] = None,
         geglu: bool = False, alpha: float = 1.0, out_fp32: bool = False,
         conv_nhw: Optional[tuple] = None, block_n: int = 0) -> torch.Tensor:
    """out[M,N] = epi(alpha * (a @ b.T + lora_t @ lora_up.T)).

    a: [M,K] bf16 (row stride free, inner stride 1), or with conv_nhw=(n,h,w) a contiguous NHWC
       image [n*h*w, C] convolved 3x3/s1/p1 with b = [N, 9*C] (k = tap*C + c).
    b: [N,K] bf16 K-major.  lora_t: [M,K2], lora_up: [N,K2] (K2 in 16..64, multiple of 16).
    """
    lib = capi.load()
    _req_bf16(a, "a"), _req_bf16(b, "b")
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, N = a.shape[0], b.shape[0]
    K = b.shape[1]
    n_out = N // 2 if geglu else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=torch.float32 if out_fp32 else BF16)
    assert out.stride(1) == 1 and out.shape == (M, n_out)
    g = GemmArgs()
    g.a, g.b, g.d = _ptr(a), _ptr(b), _ptr(out)
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldd = a.stride(0), b.stride(0), out.stride(0)
    g.batch0 = g.batch1 = 1
    if conv_nhw is not None:
        n, h, w = conv_nhw
        C = a.shape[1]
        assert a.is_contiguous() and M == n * h * w and K == 9 * C
        g.mode, g.cn, g.ch, g.cw, g.cc = 1, n, h, w, C
    else:
        assert a.shape[1] == K, (a.shape, b.shape)
    if lora_t is not None:
        _req_bf16(lora_t, "lora_t"), _req_bf16(lora_up, "lora_up")
        assert lora_t.shape[0] == M and lora_up.shape[0] == N and lora_t.shape[1] == lora_up.shape[1]
        g.a2, g.b2, g.K2 = _ptr(lora_t), _ptr(lora_up), lora_t.shape[1]
        g.lda2, g.ldb2 = lora_t.stride(0), lora_up.stride(0)
    if bias is not None:
        _req_bf16(bias, "bias")
        assert bias.numel() == N and bias.is_contiguous()
        g.bias = _ptr(bias)
    if rowbias is not None:
        _req_bf16(rowbias, "rowbias")
        assert rowbias.stride(1) == 1 and rowbias.shape[1] == N
        g.rowbias, g.rows_per_group, g.ld_rowbias = _ptr(rowbias), rows_per_group, rowbias.stride(0)
    if residual is not None:
        _req_bf16(residual, "residual")
        assert residual.shape == (M, n_out) and residual.stride(1) == 1
        g.residual, g.ldr = _ptr(residual), residual.stride(0)
    g.epilogue = 1 if geglu else 0
    g.alpha = alpha
    g.out_fp32 = 1 if out_fp32 else 0
    g.block_n = block_n
    capi.check(lib.leco_gemm_bf16(ctypes.byref(g), _stream()), "leco_gemm_bf16")
    return out


def gemm_batched(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, alpha: float = 1.0,
                 block_n: int = 0) -> torch.Tensor:
    """out[b1,b0,M,N] = alpha * a[b1,b0,M,K] @ b[b1,b0,N,K]^T over arbitrary-strided 4-D views
    (inner stride 1).  out may be bf16 or fp32."""
    lib = capi.load()
    _req_bf16(a, "a"), _req_bf16(b, "b")
    assert a.dim() == 4 and b.dim() == 4 and out.dim() == 4
    assert a.stride(3) == 1 and b.stride(3) == 1 and out.stride(3) == 1
    B1, B0, M, K = a.shape
    N = b.shape[2]
    assert b.shape == (B1, B0, N, K) and out.shape == (B1, B0, M, N)
    g = GemmArgs()
    g.a, g.b, g.d = _ptr(a), _ptr(b), _ptr(out)
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldd = a.stride(2), b.stride(2), out.stride(2)
    g.batch0, g.batch1 = B0, B1
    g.a_bs0, g.a_bs1 = a.stride(1), a.stride(0)
    g.b_bs0, g.b_bs1 = b.stride(1), b.stride(0)
    g.d_bs0, g.d_bs1 = out.stride(1), out.stride(0)
    g.alpha = alpha
    g.out_fp32 = 1 if out.dtype == torch.float32 else 0
    g.block_n = block_n
    capi.check(lib.leco_gemm_bf16(ctypes.byref(g), _stream()), "leco_gemm_bf16(batched)")
    return out
