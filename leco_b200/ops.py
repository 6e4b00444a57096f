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
GEMM_2CTA = int(__import__('os').environ.get('LECO_GEMM_2CTA', '2'))  # 0: never, 1: always, 2: library chooses per shape


_GEMM_DEBUG_MODE = int(__import__('os').environ.get('LECO_GEMM_DEBUG', '0'))
SPLIT_K = int(__import__('os').environ.get('LECO_SPLIT_K', '1'))
_SPLITK_WS = {}


def _splitk_workspace(device) -> torch.Tensor:
    """Zeroed fp32 scratch for split-K partial sums (the finalize kernel leaves it zeroed).  One per device:
    GEMMs on a stream are ordered, so it can be shared."""
    ws = _SPLITK_WS.get(device)
    if ws is None:
        ws = _SPLITK_WS[device] = torch.zeros(8 << 20, device=device, dtype=torch.float32)  # 32 MiB
    return ws


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
         conv_nhw: Optional[tuple] = None, block_n: int = 0, cta_pair: Optional[int] = None,
         fl_ad: Optional[torch.Tensor] = None, fl_bup: Optional[torch.Tensor] = None, fl_scale: float = 1.0,
         fl_rank: int = 0, fl_t_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] = epi(alpha * (a @ b.T + lora_t @ lora_up.T)).

    a: [M,K] bf16 (row stride free, inner stride 1), or with conv_nhw=(n,h,w) a contiguous NHWC
       image [n*h*w, C] convolved 3x3/s1/p1 with b = [N, 9*C] (k = tap*C + c).
    b: [N,K] bf16 K-major.  lora_t: [M,K2], lora_up: [N,K2] (K2 in 16..64, multiple of 16).
    In-kernel LoRA (instead of lora_t/lora_up): fl_ad [Kl,K] stacked lora_down, fl_bup [N,Kl] stacked lora_up:
    out += fl_scale * (a @ fl_ad.T) @ fl_bup.T inside the same kernel; fl_t_out [M,Kl] receives fl_scale*a@fl_ad.T.
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
    g.debug_mode = _GEMM_DEBUG_MODE
    g.cta_pair = GEMM_2CTA if cta_pair is None else int(cta_pair)
    if fl_ad is not None:
        _req_bf16(fl_ad, "fl_ad"), _req_bf16(fl_bup, "fl_bup")
        kl = fl_ad.shape[0]
        assert fl_ad.shape[1] == K and fl_bup.shape == (N, kl) and fl_ad.stride(1) == 1 and fl_bup.stride(1) == 1
        g.fl_ad, g.fl_bup, g.fl_kl, g.fl_rank = _ptr(fl_ad), _ptr(fl_bup), kl, (fl_rank or kl)
        g.fl_ld_ad, g.fl_ld_bup, g.fl_scale = fl_ad.stride(0), fl_bup.stride(0), fl_scale
        g.cta_pair = 0
        if fl_t_out is not None:
            _req_bf16(fl_t_out, "fl_t_out")
            assert fl_t_out.shape == (M, kl) and fl_t_out.stride(1) == 1
            g.fl_t_out, g.fl_ld_t = _ptr(fl_t_out), fl_t_out.stride(0)
    if SPLIT_K and g.cta_pair != 1:
        ws = _splitk_workspace(a.device)
        g.splitk_ws, g.splitk_ws_bytes = ws.data_ptr(), ws.numel() * 4
    capi.check(lib.leco_gemm_bf16(ctypes.byref(g), _stream()), "leco_gemm_bf16")
    return out


def gemm_batched(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, alpha: float = 1.0,
                 block_n: int = 0, n_pad: int = 0) -> torch.Tensor:
    """out[b1,b0,M,N] = alpha * a[b1,b0,M,K] @ b[b1,b0,N,K]^T over arbitrary-strided 4-D views
    (inner stride 1).  out may be bf16 or fp32."""
    lib = capi.load()
    _req_bf16(a, "a"), _req_bf16(b, "b")
    assert a.dim() == 4 and b.dim() == 4 and out.dim() == 4
    assert a.stride(3) == 1 and b.stride(3) == 1 and out.stride(3) == 1
    B1, B0, M, K = a.shape
    N = b.shape[2]
    assert b.shape == (B1, B0, N, K)
    g = GemmArgs()
    if n_pad:  # B has N real rows; output columns [N, n_pad) are produced as zeros
        g.b_rows, N = N, n_pad
    assert out.shape == (B1, B0, M, N)
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


# --------------------------------------------------------------------------------------
# everything below: one thin wrapper per C entry point (allocation + pointer passing only)
# --------------------------------------------------------------------------------------
def _lib():
    return capi.load()


def conv_in(x_nchw: torch.Tensor, w_oihw: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    n, c, h, w = x_nchw.shape
    assert c == 4 and x_nchw.is_contiguous() and x_nchw.dtype in (torch.float32, BF16)
    cout = w_oihw.shape[0]
    y = torch.empty((n * h * w, cout), device=x_nchw.device, dtype=BF16)
    capi.check(_lib().leco_conv_in(_ptr(x_nchw), int(x_nchw.dtype == torch.float32), _ptr(w_oihw), _ptr(bias),
                                   _ptr(y), n, h, w, cout, _stream()), "leco_conv_in")
    return y


_CONV_OUT_W8 = {}
CONV_OUT_GEMM = int(__import__('os').environ.get('LECO_CONV_OUT_GEMM', '1'))


def conv_out(x: torch.Tensor, w_ohwi: torch.Tensor, bias: torch.Tensor, n: int, h: int, w: int) -> torch.Tensor:
    cout, c = w_ohwi.shape[0], x.shape[1]
    y = torch.empty((n, cout, h, w), device=x.device, dtype=torch.float32)
    if CONV_OUT_GEMM and c % 64 == 0 and w <= 128 and cout <= 8:
        # tensor-core path: implicit-GEMM 3x3 conv with the 4 output channels padded to N = 8 (fp32 accumulate and
        # output, like the CUDA-core kernel), then a repack to NCHW (+ bias).  The padded weight is cached per tensor.
        key = (w_ohwi.data_ptr(), w_ohwi._version)
        w8 = _CONV_OUT_W8.get(key)
        if w8 is None:
            w8 = torch.zeros((8, 9 * c), device=x.device, dtype=BF16)
            w8[:cout] = w_ohwi.reshape(cout, 9 * c)
            if len(_CONV_OUT_W8) > 16:
                _CONV_OUT_W8.clear()
            _CONV_OUT_W8[key] = w8
        y8 = gemm(x, w8, conv_nhw=(n, h, w), out_fp32=True)
        capi.check(_lib().leco_cols_to_nchw(_ptr(y8), 8, _ptr(bias), _ptr(y), n, h * w, cout, _stream()),
                   "leco_cols_to_nchw")
        return y
    capi.check(_lib().leco_conv_out(_ptr(x), _ptr(w_ohwi), _ptr(bias), _ptr(y), n, h, w, c, cout, _stream()),
               "leco_conv_out")
    return y


def conv_out_bwd(dy: torch.Tensor, w_ohwi: torch.Tensor, c: int) -> torch.Tensor:
    n, cout, h, w = dy.shape
    assert dy.dtype == torch.float32 and dy.is_contiguous()
    dx = torch.empty((n * h * w, c), device=dy.device, dtype=BF16)
    capi.check(_lib().leco_conv_out_bwd(_ptr(dy), _ptr(w_ohwi), _ptr(dx), n, h, w, c, cout, _stream()),
               "leco_conv_out_bwd")
    return dx


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    assert t.dtype == torch.float32 and t.dim() == 1
    out = torch.empty((t.shape[0], dim), device=t.device, dtype=BF16)
    capi.check(_lib().leco_timestep_embedding(_ptr(t), _ptr(out), t.shape[0], dim, _stream()),
               "leco_timestep_embedding")
    return out


def silu(x: torch.Tensor) -> torch.Tensor:
    y = torch.empty_like(x)
    capi.check(_lib().leco_silu(_ptr(x), _ptr(y), x.numel(), _stream()), "leco_silu")
    return y


def activation(x: torch.Tensor, kind: int) -> torch.Tensor:
    """kind 0 SiLU, 1 quick_gelu, 2 erf GELU (the text encoder's MLP, text_encoder.py)."""
    _req_bf16(x, "x")
    assert x.is_contiguous()
    y = torch.empty_like(x)
    capi.check(_lib().leco_activation(_ptr(x), _ptr(y), x.numel(), kind, _stream()), "leco_activation")
    return y


def embed_tokens(ids: torch.Tensor, tok: torch.Tensor, pos: torch.Tensor, seq: int) -> torch.Tensor:
    """ids int32 [rows] -> bf16 [rows, D] = tok[ids] + pos[row % seq]."""
    _req_bf16(tok, "tok"), _req_bf16(pos, "pos")
    assert ids.dtype == torch.int32 and ids.is_contiguous() and tok.is_contiguous() and pos.is_contiguous()
    assert pos.shape[0] >= seq and pos.shape[1] == tok.shape[1]
    out = torch.empty((ids.numel(), tok.shape[1]), device=tok.device, dtype=BF16)
    capi.check(_lib().leco_embed_tokens(_ptr(ids), _ptr(tok), _ptr(pos), _ptr(out), ids.numel(), seq, tok.shape[1],
                                        tok.shape[0], _stream()), "leco_embed_tokens")
    return out


def add_(y: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    assert y.shape == x.shape and y.is_contiguous() and x.is_contiguous()
    capi.check(_lib().leco_add_inplace(_ptr(y), _ptr(x), y.numel(), _stream()), "leco_add_inplace")
    return y


def geglu_fwd(pre: torch.Tensor) -> torch.Tensor:
    M, H2 = pre.shape
    out = torch.empty((M, H2 // 2), device=pre.device, dtype=BF16)
    capi.check(_lib().leco_geglu_fwd(_ptr(pre), _ptr(out), M, H2 // 2, _stream()), "leco_geglu_fwd")
    return out


def geglu_bwd(pre: torch.Tensor, dout: torch.Tensor) -> torch.Tensor:
    M, H2 = pre.shape
    dpre = torch.empty_like(pre)
    capi.check(_lib().leco_geglu_bwd(_ptr(pre), _ptr(dout), _ptr(dpre), M, H2 // 2, _stream()), "leco_geglu_bwd")
    return dpre


def copy_cols(src: torch.Tensor, scol0: int, dst: torch.Tensor, dcol0: int, ncols: int):
    assert src.stride(1) == 1 and dst.stride(1) == 1 and src.shape[0] == dst.shape[0]
    capi.check(_lib().leco_copy_cols(_ptr(src), src.stride(0), scol0, _ptr(dst), dst.stride(0), dcol0,
                                     src.shape[0], ncols, _stream()), "leco_copy_cols")


def concat2(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    out = torch.empty((a.shape[0], a.shape[1] + b.shape[1]), device=a.device, dtype=BF16)
    copy_cols(a, 0, out, 0, a.shape[1])
    copy_cols(b, 0, out, a.shape[1], b.shape[1])
    return out


def split2(x: torch.Tensor, c1: int):
    a = torch.empty((x.shape[0], c1), device=x.device, dtype=BF16)
    b = torch.empty((x.shape[0], x.shape[1] - c1), device=x.device, dtype=BF16)
    copy_cols(x, 0, a, 0, c1)
    copy_cols(x, c1, b, 0, x.shape[1] - c1)
    return a, b


def upsample2x(x: torch.Tensor, n: int, h: int, w: int) -> torch.Tensor:
    c = x.shape[1]
    y = torch.empty((n * 4 * h * w, c), device=x.device, dtype=BF16)
    capi.check(_lib().leco_upsample2x(_ptr(x), _ptr(y), n, h, w, c, _stream()), "leco_upsample2x")
    return y


def upsample2x_bwd(dy: torch.Tensor, n: int, h: int, w: int) -> torch.Tensor:
    c = dy.shape[1]
    dx = torch.empty((n * h * w, c), device=dy.device, dtype=BF16)
    capi.check(_lib().leco_upsample2x_bwd(_ptr(dy), _ptr(dx), n, h, w, c, _stream()), "leco_upsample2x_bwd")
    return dx


def im2col_s2(x: torch.Tensor, n: int, h: int, w: int) -> torch.Tensor:
    c = x.shape[1]
    col = torch.empty((n * (h // 2) * (w // 2), 9 * c), device=x.device, dtype=BF16)
    capi.check(_lib().leco_im2col_s2(_ptr(x), _ptr(col), n, h, w, c, _stream()), "leco_im2col_s2")
    return col


def im2col_s1(x: torch.Tensor, n: int, h: int, w: int) -> torch.Tensor:
    c = x.shape[1]
    col = torch.empty((n * h * w, 9 * c), device=x.device, dtype=BF16)
    capi.check(_lib().leco_im2col_s1(_ptr(x), _ptr(col), n, h, w, c, _stream()), "leco_im2col_s1")
    return col


def rowgroup_sum(x: torch.Tensor, n: int, hw: int) -> torch.Tensor:
    c = x.shape[1]
    out = torch.empty((n, c), device=x.device, dtype=BF16)
    capi.check(_lib().leco_rowgroup_sum(_ptr(x), _ptr(out), n, hw, c, _stream()), "leco_rowgroup_sum")
    return out


def col2im_s2(dcol: torch.Tensor, n: int, h: int, w: int) -> torch.Tensor:
    c = dcol.shape[1] // 9
    dx = torch.empty((n * h * w, c), device=dcol.device, dtype=BF16)
    capi.check(_lib().leco_col2im_s2(_ptr(dcol), _ptr(dx), n, h, w, c, _stream()), "leco_col2im_s2")
    return dx


def transpose_batched(src: torch.Tensor, cols_pad: int = 0) -> torch.Tensor:
    """src: 4-D view [b1,b0,rows,cols] (inner stride 1) -> new [b1,b0,cols,rows_pad] (zero padded)."""
    _req_bf16(src, "src")
    B1, B0, rows, cols = src.shape
    rows_pad = max(cols_pad, rows)
    out = torch.empty((B1, B0, cols, rows_pad), device=src.device, dtype=BF16)
    capi.check(_lib().leco_transpose(_ptr(src), _ptr(out), rows, cols, rows_pad, src.stride(2), src.stride(1),
                                     src.stride(0), out.stride(2), out.stride(1), out.stride(0), B0, B1,
                                     _stream()), "leco_transpose")
    return out


def softmax_rows(s: torch.Tensor, n_valid: int, n_pad: int, causal_sq: int = 0) -> torch.Tensor:
    """s: fp32 [..., ld] contiguous -> p bf16 [..., n_pad].  causal_sq > 0: row r sees columns <= r % causal_sq."""
    assert s.dtype == torch.float32 and s.is_contiguous()
    rows = s.numel() // s.shape[-1]
    p = torch.empty(s.shape[:-1] + (n_pad,), device=s.device, dtype=BF16)
    if causal_sq:
        capi.check(_lib().leco_softmax_rows_causal(_ptr(s), _ptr(p), rows, n_valid, n_pad, s.shape[-1], n_pad,
                                                   causal_sq, _stream()), "leco_softmax_rows_causal")
    else:
        capi.check(_lib().leco_softmax_rows(_ptr(s), _ptr(p), rows, n_valid, n_pad, s.shape[-1], n_pad, _stream()),
                   "leco_softmax_rows")
    return p


def softmax_bwd_rows(p: torch.Tensor, dp: torch.Tensor, n_valid: int, scale: float) -> torch.Tensor:
    assert dp.dtype == torch.float32 and dp.is_contiguous() and p.is_contiguous()
    rows = p.numel() // p.shape[-1]
    ds = torch.empty_like(p)
    capi.check(_lib().leco_softmax_bwd_rows(_ptr(p), _ptr(dp), _ptr(ds), rows, n_valid, p.shape[-1], p.shape[-1],
                                            dp.shape[-1], scale, _stream()), "leco_softmax_bwd_rows")
    return ds


_gn_ws = {}


def _gn_workspace(dev, n, G):
    key = (dev, n, G)
    if key not in _gn_ws:
        _gn_ws[key] = torch.empty(int(_lib().leco_group_norm_workspace_bytes(n, G)), device=dev,
                                  dtype=torch.uint8)
    return _gn_ws[key]


GN_FUSED = int(__import__('os').environ.get('LECO_GN_FUSED', '0'))   # 1: single-launch GroupNorm forward (grid barrier); measured 245.4 vs 244.1 ms / iteration for the two-launch path, so off by default
# "v2": the statistics kernel finishes mean / rstd (last block per sample), the normalise kernel only reads them;
# "v1": the round-1 pair (every normalise block re-folds the partial sums)
# "v3" (default): one launch, a thread-block cluster per sample (falls back to v2 inside the library when a sample does
# not fit one cluster's shared memory; LECO_GN_CLUSTER=0/8/16 picks the cluster size)
GN_IMPL = __import__('os').environ.get('LECO_GN_IMPL', 'v3')
_GN_BARRIERS = {}
_GN_COUNTERS = {}


def _gn_counters(dev) -> torch.Tensor:
    c = _GN_COUNTERS.get(dev)
    if c is None:
        c = _GN_COUNTERS[dev] = torch.zeros(4096, device=dev, dtype=torch.int32)
    return c



def _gn_barriers(dev) -> torch.Tensor:
    """Persistent zero-initialised per-sample barrier state of the single-launch GroupNorm (one per device; created on
    first use, i.e. during the eager warm-up that precedes any graph capture)."""
    b = _GN_BARRIERS.get(dev)
    if b is None:
        b = _GN_BARRIERS[dev] = torch.zeros(4096 * 8, device=dev, dtype=torch.uint8)
    return b


def group_norm(x: torch.Tensor, n: int, hw: int, gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float,
               silu_act: bool):
    C = x.shape[1]
    assert x.is_contiguous() and x.shape[0] == n * hw
    y = torch.empty_like(x)
    stats = torch.empty((n, groups, 2), device=x.device, dtype=torch.float32)
    ws = torch.empty(int(_lib().leco_group_norm_workspace_bytes(n, groups)), device=x.device, dtype=torch.uint8)
    if GN_IMPL == "v3" and not GN_FUSED and n <= 4096:
        capi.check(_lib().leco_group_norm_v3(_ptr(x), _ptr(y), _ptr(stats), _ptr(gamma), _ptr(beta), n, hw, C, groups,
                                             eps, int(silu_act), _ptr(ws), _ptr(_gn_counters(x.device)), _stream()),
                   "leco_group_norm_v3")
        return y, stats
    if GN_IMPL == "v2" and not GN_FUSED and n <= 4096:
        capi.check(_lib().leco_group_norm_v2(_ptr(x), _ptr(y), _ptr(stats), _ptr(gamma), _ptr(beta), n, hw, C, groups,
                                             eps, int(silu_act), _ptr(ws), _ptr(_gn_counters(x.device)), _stream()),
                   "leco_group_norm_v2")
        return y, stats
    if GN_FUSED and n <= 148 and C <= 3072 and n <= 4096:
        capi.check(_lib().leco_group_norm_fused(_ptr(x), _ptr(y), _ptr(stats), _ptr(gamma), _ptr(beta), n, hw, C, groups,
                                                eps, int(silu_act), _ptr(ws), _ptr(_gn_barriers(x.device)), _stream()),
                   "leco_group_norm_fused")
        return y, stats
    capi.check(_lib().leco_group_norm(_ptr(x), _ptr(y), _ptr(stats), _ptr(gamma), _ptr(beta), n, hw, C, groups,
                                      eps, int(silu_act), _ptr(ws), _stream()), "leco_group_norm")
    return y, stats


def group_norm_bwd(x, dz, stats, gamma, beta, n, hw, groups, silu_act):
    C = x.shape[1]
    dx = torch.empty_like(x)
    ws = torch.empty(int(_lib().leco_group_norm_workspace_bytes(n, groups)), device=x.device, dtype=torch.uint8)
    capi.check(_lib().leco_group_norm_bwd(_ptr(x), _ptr(dz), _ptr(dx), _ptr(stats), _ptr(gamma), _ptr(beta), n, hw,
                                          C, groups, int(silu_act), _ptr(ws), _stream()), "leco_group_norm_bwd")
    return dx


def layer_norm(x: torch.Tensor, gamma, beta, eps: float, want_stats: bool = False):
    M, C = x.shape
    assert x.is_contiguous()
    y = torch.empty_like(x)
    stats = torch.empty((M, 2), device=x.device, dtype=torch.float32) if want_stats else None
    capi.check(_lib().leco_layer_norm(_ptr(x), _ptr(y), _ptr(stats), _ptr(gamma), _ptr(beta), M, C, eps,
                                      _stream()), "leco_layer_norm")
    return y, stats


def layer_norm_bwd(x, dy, stats, gamma):
    M, C = x.shape
    dx = torch.empty_like(x)
    capi.check(_lib().leco_layer_norm_bwd(_ptr(x), _ptr(dy), _ptr(dx), _ptr(stats), _ptr(gamma), M, C, _stream()),
               "leco_layer_norm_bwd")
    return dx


def tn_reduce(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, scale: float = 1.0,
              transpose_out: bool = False):
    """out[N1,N2] += scale * a[M,N1]^T @ b[M,N2]   (out fp32; with transpose_out it is [N2,N1])."""
    assert out.dtype == torch.float32 and out.stride(1) == 1 and a.stride(1) == 1 and b.stride(1) == 1
    assert out.shape == ((b.shape[1], a.shape[1]) if transpose_out else (a.shape[1], b.shape[1]))
    capi.check(_lib().leco_tn_reduce(_ptr(a), a.stride(0), _ptr(b), b.stride(0), _ptr(out), out.stride(0),
                                     a.shape[0], a.shape[1], b.shape[1], scale, int(transpose_out), _stream()),
               "leco_tn_reduce")


def adamw_flat(params, grads, exp_avg, exp_avg_sq, mask, hyper, zero_grad=True):
    capi.check(_lib().leco_adamw_flat(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq),
                                      int(exp_avg.dtype == torch.float32), _ptr(mask), _ptr(hyper),
                                      params.numel(), int(zero_grad), _stream()), "leco_adamw_flat")


def optim_flat(params, grads, exp_avg, exp_avg_sq, mask, hyper16, zero_grad=True):
    """leco_optim_flat: AdamW / Adam / Lion over the flat buffer (hyper16 = fp32[16], see include/leco_b200.h)."""
    assert hyper16.numel() >= 16 and hyper16.dtype == torch.float32
    capi.check(_lib().leco_optim_flat(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq),
                                      int(exp_avg.dtype == torch.float32), _ptr(mask), _ptr(hyper16),
                                      params.numel(), int(zero_grad), _stream()), "leco_optim_flat")


def optim_flat_master(master, shadow, grads, exp_avg, exp_avg_sq, mask, hyper16, zero_grad=True):
    """leco_optim_flat_master: fp32 master parameters + fp32 moments, bf16 operand copy written in the same pass."""
    assert hyper16.numel() >= 16 and hyper16.dtype == torch.float32
    assert master.dtype == exp_avg.dtype == exp_avg_sq.dtype == torch.float32 and shadow.dtype == BF16
    assert master.numel() == shadow.numel() == exp_avg.numel()
    capi.check(_lib().leco_optim_flat_master(_ptr(master), _ptr(shadow), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq),
                                             _ptr(mask), _ptr(hyper16), master.numel(), int(zero_grad), _stream()),
               "leco_optim_flat_master")


def transpose_tiles(src_flat, dst_flat, tiles, n_tiles):
    capi.check(_lib().leco_transpose_tiles(_ptr(src_flat), _ptr(dst_flat), _ptr(tiles), int(n_tiles), _stream()),
               "leco_transpose_tiles")


def set_deterministic(on: bool):
    """Fixed-order reductions (no fp32 atomics across CTAs, no split-K, materialised attention backward): bit-repeatable
    steps, slower."""
    capi.check(_lib().leco_set_deterministic(int(bool(on))), "leco_set_deterministic")
    _DETERMINISTIC[0] = bool(on)


def guided_step(eps_pair, x, coef, want_x=True, want_guided=False):
    half = eps_pair.numel() // 2
    shape = (eps_pair.shape[0] // 2,) + tuple(eps_pair.shape[1:])
    x_out = torch.empty(shape, device=eps_pair.device, dtype=torch.float32) if want_x else None
    g_out = torch.empty(shape, device=eps_pair.device, dtype=torch.float32) if want_guided else None
    capi.check(_lib().leco_guided_step(_ptr(eps_pair), _ptr(x), _ptr(x_out), _ptr(g_out), _ptr(coef), half,
                                       _stream()), "leco_guided_step")
    return x_out, g_out


def sched_step(eps_pair, x, coef, noise=None, hist=None, out=None):
    """General scheduler update fused with the CFG combine (see include/leco_b200.h: leco_sched_step)."""
    half = eps_pair.numel() // 2
    assert x.dtype == torch.float32 and eps_pair.dtype == torch.float32 and x.numel() == half and coef.numel() >= 12
    out = torch.empty_like(x) if out is None else out
    if hist is not None:
        assert hist.dtype == torch.float32 and hist.numel() == 4 * half and hist.is_contiguous()
    capi.check(_lib().leco_sched_step(_ptr(eps_pair), _ptr(x), _ptr(noise), _ptr(hist), _ptr(out), _ptr(coef), half,
                                      _stream()), "leco_sched_step")
    return out


def scale_by_dev(x, coef, idx: int):
    assert x.dtype == torch.float32 and x.is_contiguous()
    y = torch.empty_like(x)
    capi.check(_lib().leco_scale_by_dev(_ptr(x), _ptr(y), _ptr(coef), idx, x.numel(), _stream()), "leco_scale_by_dev")
    return y


def leco_loss(target, positive, neutral, uncond, sign_times_guidance: float, want_grad=True):
    loss = torch.empty((1,), device=target.device, dtype=torch.float32)
    dt = torch.empty_like(target) if want_grad else None
    capi.check(_lib().leco_loss(_ptr(target), _ptr(positive), _ptr(neutral), _ptr(uncond), sign_times_guidance,
                                _ptr(loss), _ptr(dt), target.numel(), _stream()), "leco_loss")
    return loss, dt


def axpby(x, y, a: float, b: float):
    assert x.dtype == y.dtype and x.dtype in (torch.float32, BF16) and x.is_contiguous() and y.is_contiguous()
    out = torch.empty_like(x)
    capi.check(_lib().leco_axpby(_ptr(x), _ptr(y), _ptr(out), a, b, x.numel(), int(x.dtype == torch.float32),
                                 _stream()), "leco_axpby")
    return out


def cast_f32_to_bf16(x, out=None):
    y = torch.empty(x.shape, device=x.device, dtype=BF16) if out is None else out
    assert x.dtype == torch.float32 and y.dtype == BF16 and y.numel() == x.numel() and x.is_contiguous() and y.is_contiguous()
    capi.check(_lib().leco_cast_f32_to_bf16(_ptr(x), _ptr(y), x.numel(), _stream()), "leco_cast")
    return y


def cast_bf16_to_f32(x):
    y = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    capi.check(_lib().leco_cast_bf16_to_f32(_ptr(x), _ptr(y), x.numel(), _stream()), "leco_cast")
    return y


# --------------------------------------------------------------------------------------
# plumbing helpers (allocation / memset / tiny copies: torch's allocator, no math)
# --------------------------------------------------------------------------------------
def clone(x):
    return x.clone()


def zeros_like(x):
    return torch.zeros_like(x)


def empty_like(x):
    return torch.empty_like(x)


def zeros(shape, like, dtype=None):
    return torch.zeros(shape, device=like.device, dtype=dtype or like.dtype)


def empty(shape, like, dtype=None):
    return torch.empty(shape, device=like.device, dtype=dtype or like.dtype)


def cat_cols(a, b):
    return torch.cat([a, b], dim=1).contiguous()


def transpose2d(x: torch.Tensor) -> torch.Tensor:
    """[R,C] -> contiguous [C,R]."""
    return transpose_batched(x.unsqueeze(0).unsqueeze(0)).squeeze(0).squeeze(0)


# --------------------------------------------------------------------------------------
# attention v0: S = scale QK^T (tcgen05 batched GEMM, fp32) -> row softmax -> P V (tcgen05)
# P is materialised; the fused flash kernel replaces this on the no-grad path.
# --------------------------------------------------------------------------------------
def _heads_view(t2d: torch.Tensor, nb: int, s: int, heads: int, d: int) -> torch.Tensor:
    """[nb*s, heads*d] (row stride free) -> strided view [nb, heads, s, d]."""
    return t2d.unflatten(0, (nb, s)).unflatten(2, (heads, d)).permute(0, 2, 1, 3)


import os as _os

# 0: V used in place (MN-major B operand); 1: transposed copy of V first
FLASH_V_MODE = int(_os.environ.get("LECO_FLASH_V_MODE", "0"))
# "flash" (fused, no-grad path) or "v0" (materialised P; always used on the grad path)
ATTENTION_IMPL = _os.environ.get("LECO_ATTENTION", "flash")
# widest head the fused forward takes (LECO_FLASH_WIDE=0: heads above 64 use the materialised path, for A/B runs)
FLASH_WIDE_MAX_D = 192 if _os.environ.get("LECO_FLASH_WIDE", "1") != "0" else 64


def flash_attention(qt, kt, vt, nb, sq, skv, heads, d, scale, v_mode=None):
    """Fused tcgen05 attention forward (head dim <= 192).  Views [rows, heads*d] with free row stride."""
    v_mode = FLASH_V_MODE if v_mode is None else v_mode
    o = torch.empty((nb * sq, heads * d), device=qt.device, dtype=BF16)
    v_t, skv_pad = None, 0
    if v_mode == 1:
        skv_pad = (skv + 7) // 8 * 8
        v_t = transpose_batched(_heads_view(vt, nb, skv, heads, d), cols_pad=skv_pad)   # [nb, heads, d, skv_pad]
    capi.check(_lib().leco_flash_attn_fwd(_ptr(qt), qt.stride(0), _ptr(kt), kt.stride(0), _ptr(vt), vt.stride(0),
                                          _ptr(v_t), skv_pad, _ptr(o), o.stride(0), nb, heads, sq, skv, d, scale,
                                          _stream()), "leco_flash_attn_fwd")
    return o


# grad pass: "flash" = fused forward (+lse) and fused backward kernel; "v0" = materialised P (always used for head dims
# > 64 and under leco_set_deterministic: the fused backward reduces dQ over key tiles with fp32 atomics)
ATTENTION_BWD_IMPL = _os.environ.get("LECO_ATTENTION_BWD", "flash")
_DETERMINISTIC = [_os.environ.get("LECO_DETERMINISTIC", "0") == "1"]


def flash_attention_lse(qt, kt, vt, nb, sq, skv, heads, d, scale):
    """Fused forward that also returns lse [nb, heads, sq] (fp32, log2 domain) for flash_attention_bwd."""
    o = torch.empty((nb * sq, heads * d), device=qt.device, dtype=BF16)
    lse = torch.empty((nb, heads, sq), device=qt.device, dtype=torch.float32)
    capi.check(_lib().leco_flash_attn_fwd_lse(_ptr(qt), qt.stride(0), _ptr(kt), kt.stride(0), _ptr(vt), vt.stride(0),
                                              _ptr(o), o.stride(0), _ptr(lse), nb, heads, sq, skv, d, scale, _stream()),
               "leco_flash_attn_fwd_lse")
    return o, lse


def flash_attention_bwd(go, qt, kt, vt, o, lse, nb, sq, skv, heads, d, scale, dq_out, dk_out, dv_out):
    """Fused attention backward: writes d(q), d(k), d(v) into the given [rows, heads*d] views (any may be None)."""
    _req_bf16(go, "go")
    assert go.stride(1) == 1 and o.stride(1) == 1
    dvec = torch.empty((nb, heads, sq), device=go.device, dtype=torch.float32)
    capi.check(_lib().leco_attn_bwd_prep(_ptr(o), o.stride(0), _ptr(go), go.stride(0), _ptr(dvec), nb, heads, sq, d,
                                         _stream()), "leco_attn_bwd_prep")
    dq_acc = torch.zeros((nb, heads, sq, 64), device=go.device, dtype=torch.float32)
    capi.check(_lib().leco_flash_attn_bwd(_ptr(qt), qt.stride(0), _ptr(kt), kt.stride(0), _ptr(vt), vt.stride(0),
                                          _ptr(go), go.stride(0), _ptr(lse), _ptr(dvec), _ptr(dq_acc),
                                          _ptr(dk_out), 0 if dk_out is None else dk_out.stride(0),
                                          _ptr(dv_out), 0 if dv_out is None else dv_out.stride(0),
                                          nb, heads, sq, skv, d, scale, _stream()), "leco_flash_attn_bwd")
    if dq_out is not None:
        capi.check(_lib().leco_attn_dq_cast(_ptr(dq_acc), _ptr(dq_out), dq_out.stride(0), nb, heads, sq, d, _stream()),
                   "leco_attn_dq_cast")


def attention(qt, kt, vt, nb, sq, skv, heads, d, scale, save_for_bwd=False):
    # forward: d <= 64 pipelined kernels, 64 < d <= 192 flash_attn_fwd_wide_kernel (SD1.5's d = 80 / 160 levels);
    # the fused backward covers d <= 64, so wider heads keep the materialised path on the grad pass only
    fwd_ok = ATTENTION_IMPL == "flash" and d % 8 == 0 and (d <= 64 or (d <= FLASH_WIDE_MAX_D and FLASH_V_MODE == 0))
    flash_ok = fwd_ok and d <= 64
    if not save_for_bwd and fwd_ok:
        return flash_attention(qt, kt, vt, nb, sq, skv, heads, d, scale), None
    if save_for_bwd and flash_ok and ATTENTION_BWD_IMPL == "flash" and not _DETERMINISTIC[0]:
        o, lse = flash_attention_lse(qt, kt, vt, nb, sq, skv, heads, d, scale)
        return o, ("flash", o, lse)
    return attention_v0(qt, kt, vt, nb, sq, skv, heads, d, scale, save_for_bwd)


def attention_v0(qt, kt, vt, nb, sq, skv, heads, d, scale, save_for_bwd=False, causal=False):
    skv_pad = (skv + 15) // 16 * 16
    q4, k4, v4 = _heads_view(qt, nb, sq, heads, d), _heads_view(kt, nb, skv, heads, d), _heads_view(vt, nb, skv, heads, d)
    S = torch.empty((nb, heads, sq, skv_pad), device=qt.device, dtype=torch.float32)
    gemm_batched(q4, k4, S, alpha=scale, n_pad=skv_pad if skv_pad != skv else 0)
    P = softmax_rows(S, skv, skv_pad, causal_sq=sq if causal else 0)
    del S
    Vt = transpose_batched(v4, cols_pad=skv_pad)               # [nb, heads, d, skv_pad]
    o = torch.empty((nb * sq, heads * d), device=qt.device, dtype=BF16)
    gemm_batched(P, Vt, _heads_view(o, nb, sq, heads, d))
    return o, ((P,) if save_for_bwd else None)


def attention_bwd(go, qt, kt, vt, saved, nb, sq, skv, heads, d, scale, dq_out, dk_out, dv_out):
    """Writes d(q), d(k), d(v) into the given [rows, heads*d] views (any may be None)."""
    if saved[0] == "flash":
        _, o, lse = saved
        return flash_attention_bwd(go, qt, kt, vt, o, lse, nb, sq, skv, heads, d, scale, dq_out, dk_out, dv_out)
    (P,) = saved
    skv_pad = P.shape[-1]
    pad = skv_pad if skv_pad != skv else 0
    q4, k4, v4 = _heads_view(qt, nb, sq, heads, d), _heads_view(kt, nb, skv, heads, d), _heads_view(vt, nb, skv, heads, d)
    go4 = _heads_view(go, nb, sq, heads, d)
    dP = torch.empty((nb, heads, sq, skv_pad), device=go.device, dtype=torch.float32)
    gemm_batched(go4, v4, dP, n_pad=pad)                         # dP = dO V^T
    dS = softmax_bwd_rows(P, dP, skv, scale)                     # scale * P o (dP - rowsum(P o dP))
    del dP
    sq_pad = (sq + 15) // 16 * 16        # K of the two "reduce over queries" GEMMs (zero padded)
    if dv_out is not None:
        Pt = transpose_batched(P, cols_pad=sq_pad)               # [.., skv_pad, sq_pad]
        goT = transpose_batched(go4, cols_pad=sq_pad)            # [.., d, sq_pad]
        gemm_batched(Pt[:, :, :skv], goT, _heads_view(dv_out, nb, skv, heads, d))
        del Pt, goT
    if dq_out is not None:
        Kt = transpose_batched(k4, cols_pad=skv_pad)             # [.., d, skv_pad]
        gemm_batched(dS, Kt, _heads_view(dq_out, nb, sq, heads, d))
    if dk_out is not None:
        dSt = transpose_batched(dS, cols_pad=sq_pad)             # [.., skv_pad, sq_pad]
        Qt = transpose_batched(q4, cols_pad=sq_pad)              # [.., d, sq_pad]
        gemm_batched(dSt[:, :, :skv], Qt, _heads_view(dk_out, nb, skv, heads, d))
