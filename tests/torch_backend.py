"""TEST DOUBLE (tests/ only): plain-PyTorch implementation of the backend interface that
`leco_b200.unet.EngineUNet` drives.  Two uses:

  * CPU tests inject it (`EngineUNet(spec, backend=torch_backend)`) to check the ENGINE's
    wiring — forward program, skip connections, LoRA fusion algebra, tape/chain rule —
    against the oracle in fp32, without a GPU;
  * GPU tests use each function as the plain-PyTorch fp32 reference of the CUDA kernel
    with the same name in `leco_b200.ops`.

It is never imported by the product package.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _f(x):
    return x.float()


# ------------------------------------------------------------------ GEMM family
def gemm(a, b, out=None, *, bias=None, rowbias=None, rows_per_group=1, residual=None, lora_t=None, lora_up=None,
         geglu=False, alpha=1.0, out_fp32=False, conv_nhw=None, block_n=0, cta_pair=None,
         fl_ad=None, fl_bup=None, fl_scale=1.0, fl_rank=0, fl_t_out=None):
    dt = torch.float32 if out_fp32 else a.dtype
    if conv_nhw is not None:
        n, h, w = conv_nhw
        c = a.shape[1]
        x = _f(a).reshape(n, h, w, c).permute(0, 3, 1, 2)
        wt = _f(b).reshape(b.shape[0], 3, 3, c).permute(0, 3, 1, 2)
        y = F.conv2d(x, wt, padding=1).permute(0, 2, 3, 1).reshape(n * h * w, -1)
    else:
        y = _f(a) @ _f(b).t()
    if lora_t is not None:
        y = y + _f(lora_t) @ _f(lora_up).t()
    y = y * alpha
    if fl_ad is not None:
        if conv_nhw is not None:
            wa = _f(fl_ad).reshape(fl_ad.shape[0], 3, 3, c).permute(0, 3, 1, 2)
            t = F.conv2d(x, wa, padding=1).permute(0, 2, 3, 1).reshape(n * h * w, -1) * fl_scale
        else:
            t = fl_scale * (_f(a) @ _f(fl_ad).t())
        y = y + t @ _f(fl_bup).t()
        if fl_t_out is not None:
            fl_t_out.copy_(t.to(fl_t_out.dtype))
    if bias is not None:
        y = y + _f(bias)[None]
    if rowbias is not None:
        y = y + _f(rowbias).repeat_interleave(rows_per_group, 0)[: y.shape[0]]
    if geglu:
        hh, gg = y.chunk(2, dim=1)
        y = hh * F.gelu(gg)
    if residual is not None:
        y = y + _f(residual)
    y = y.to(dt)
    if out is not None:
        out.copy_(y)
        return out
    return y


def gemm_batched(a, b, out, *, alpha=1.0, block_n=0, n_pad=0):
    y = alpha * torch.einsum("xymk,xynk->xymn", _f(a), _f(b))
    if n_pad:
        y = F.pad(y, (0, n_pad - y.shape[-1]))
    out.copy_(y.to(out.dtype))
    return out


# ------------------------------------------------------------------ boundary convs
def conv_in(x_nchw, w_oihw, bias):
    y = F.conv2d(_f(x_nchw), _f(w_oihw), _f(bias), padding=1)
    n, c, h, w = y.shape
    return y.permute(0, 2, 3, 1).reshape(n * h * w, c).to(w_oihw.dtype)


def _ohwi_to_oihw(w_ohwi):
    o, _, c = w_ohwi.shape
    return _f(w_ohwi).reshape(o, 3, 3, c).permute(0, 3, 1, 2)


def conv_out(x, w_ohwi, bias, n, h, w):
    c = x.shape[1]
    xi = _f(x).reshape(n, h, w, c).permute(0, 3, 1, 2)
    return F.conv2d(xi, _ohwi_to_oihw(w_ohwi), _f(bias), padding=1)


def conv_out_bwd(dy, w_ohwi, c):
    n, co, h, w = dy.shape
    dx = F.conv_transpose2d(_f(dy), _ohwi_to_oihw(w_ohwi), padding=1)
    return dx.permute(0, 2, 3, 1).reshape(n * h * w, c).to(w_ohwi.dtype)


def timestep_embedding(t, dim):
    half = dim // 2
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    ang = t.float()[:, None] * freq[None]
    return torch.cat([torch.cos(ang), torch.sin(ang)], dim=1).to(ACT_DTYPE)


ACT_DTYPE = torch.float32  # tests flip this to bfloat16 when they build GPU references


# ------------------------------------------------------------------ elementwise
def silu(x):
    return F.silu(_f(x)).to(x.dtype)


def add_(y, x):
    y.copy_((_f(y) + _f(x)).to(y.dtype))
    return y


def geglu_fwd(pre):
    h, g = _f(pre).chunk(2, dim=1)
    return (h * F.gelu(g)).to(pre.dtype)


def geglu_bwd(pre, dout):
    with torch.enable_grad():
        p = _f(pre).detach().requires_grad_(True)
        h, g = p.chunk(2, dim=1)
        (h * F.gelu(g)).backward(_f(dout))
    return p.grad.to(pre.dtype)


def concat2(a, b):
    return torch.cat([a, b], dim=1).contiguous()


def split2(x, c1):
    return x[:, :c1].contiguous(), x[:, c1:].contiguous()


def upsample2x(x, n, h, w):
    c = x.shape[1]
    xi = x.reshape(n, h, w, c)
    return xi.repeat_interleave(2, 1).repeat_interleave(2, 2).reshape(n * 4 * h * w, c).contiguous()


def upsample2x_bwd(dy, n, h, w):
    c = dy.shape[1]
    return _f(dy).reshape(n, h, 2, w, 2, c).sum(dim=(2, 4)).reshape(n * h * w, c).to(dy.dtype)


def im2col_s2(x, n, h, w):
    c = x.shape[1]
    xi = _f(x).reshape(n, h, w, c).permute(0, 3, 1, 2)
    cols = F.unfold(xi, 3, padding=1, stride=2)                 # [n, c*9, L]  (c-major, then tap)
    L = cols.shape[-1]
    cols = cols.reshape(n, c, 9, L).permute(0, 3, 2, 1).reshape(n * L, 9 * c)  # k = tap*c + ch
    return cols.to(x.dtype)


def im2col_s1(x, n, h, w):
    c = x.shape[1]
    xi = _f(x).reshape(n, h, w, c).permute(0, 3, 1, 2)
    cols = F.unfold(xi, 3, padding=1, stride=1)
    L = cols.shape[-1]
    return cols.reshape(n, c, 9, L).permute(0, 3, 2, 1).reshape(n * L, 9 * c).to(x.dtype)


def rowgroup_sum(x, n, hw):
    return _f(x).reshape(n, hw, -1).sum(1).to(x.dtype)


def copy_cols(src, scol0, dst, dcol0, ncols):
    dst[:, dcol0:dcol0 + ncols] = src[:, scol0:scol0 + ncols]


def col2im_s2(dcol, n, h, w):
    c = dcol.shape[1] // 9
    L = (h // 2) * (w // 2)
    cols = _f(dcol).reshape(n, L, 9, c).permute(0, 3, 2, 1).reshape(n, c * 9, L)
    dx = F.fold(cols, (h, w), 3, padding=1, stride=2)
    return dx.permute(0, 2, 3, 1).reshape(n * h * w, c).to(dcol.dtype)


def transpose2d(x):
    return x.t().contiguous()


def transpose_batched(src, cols_pad=0):
    out = src.transpose(-1, -2).contiguous()
    if cols_pad and cols_pad > out.shape[-1]:
        out = F.pad(out, (0, cols_pad - out.shape[-1]))
    return out


# ------------------------------------------------------------------ norms
def group_norm(x, n, hw, gamma, beta, groups, eps, silu_act):
    c = x.shape[1]
    xi = _f(x).reshape(n, hw, c).permute(0, 2, 1)
    y = F.group_norm(xi, groups, _f(gamma), _f(beta), eps)
    if silu_act:
        y = F.silu(y)
    xg = xi.reshape(n, groups, -1)
    mean = xg.mean(-1)
    rstd = (xg.var(-1, unbiased=False) + eps).rsqrt()
    return y.permute(0, 2, 1).reshape(n * hw, c).to(x.dtype), torch.stack([mean, rstd], -1)


def group_norm_bwd(x, dz, stats, gamma, beta, n, hw, groups, silu_act):
    c = x.shape[1]
    cpg = c // groups
    xf = _f(x).reshape(n, hw, groups, cpg)
    mean = stats[..., 0].reshape(n, 1, groups, 1)
    rstd = stats[..., 1].reshape(n, 1, groups, 1)
    gm, bt = _f(gamma).reshape(1, 1, groups, cpg), _f(beta).reshape(1, 1, groups, cpg)
    xh = (xf - mean) * rstd
    dy = _f(dz).reshape(n, hw, groups, cpg)
    if silu_act:
        y = xh * gm + bt
        sg = torch.sigmoid(y)
        dy = dy * sg * (1 + y * (1 - sg))
    g = dy * gm
    m1 = g.mean(dim=(1, 3), keepdim=True)
    m2 = (g * xh).mean(dim=(1, 3), keepdim=True)
    return (rstd * (g - m1 - xh * m2)).reshape(n * hw, c).to(x.dtype)


def layer_norm(x, gamma, beta, eps, want_stats=False):
    xf = _f(x)
    y = F.layer_norm(xf, (x.shape[1],), _f(gamma), _f(beta), eps)
    stats = None
    if want_stats:
        mean = xf.mean(-1)
        rstd = (xf.var(-1, unbiased=False) + eps).rsqrt()
        stats = torch.stack([mean, rstd], -1)
    return y.to(x.dtype), stats


def layer_norm_bwd(x, dy, stats, gamma):
    xf = _f(x)
    mean, rstd = stats[:, 0:1], stats[:, 1:2]
    xh = (xf - mean) * rstd
    g = _f(dy) * _f(gamma)[None]
    dx = rstd * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    return dx.to(x.dtype)


# ------------------------------------------------------------------ training side
def tn_reduce(a, b, out, scale=1.0, transpose_out=False):
    r = scale * (_f(a).t() @ _f(b))
    out += (r.t() if transpose_out else r)


def clone(x):
    return x.clone()


def zeros_like(x):
    return torch.zeros_like(x)


def empty_like(x):
    return torch.zeros_like(x)


def empty(shape, like, dtype=None):
    return torch.zeros(shape, device=like.device, dtype=dtype or like.dtype)


def zeros(shape, like, dtype=None):
    return torch.zeros(shape, device=like.device, dtype=dtype or like.dtype)


def cat_cols(a, b):
    return torch.cat([a, b], dim=1).contiguous()


# ------------------------------------------------------------------ attention
def _heads(t2d, nb, s, heads, d):
    return _f(t2d).reshape(nb, s, heads, d).permute(0, 2, 1, 3)


def attention(qt, kt, vt, nb, sq, skv, heads, d, scale, save_for_bwd=False):
    q, k, v = _heads(qt, nb, sq, heads, d), _heads(kt, nb, skv, heads, d), _heads(vt, nb, skv, heads, d)
    p = torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(nb * sq, heads * d).to(qt.dtype)
    return o, ((p,) if save_for_bwd else None)


def attention_bwd(go, qt, kt, vt, saved, nb, sq, skv, heads, d, scale, dq_out, dk_out, dv_out):
    with torch.enable_grad():
        q = _heads(qt, nb, sq, heads, d).detach().requires_grad_(True)
        k = _heads(kt, nb, skv, heads, d).detach().requires_grad_(True)
        v = _heads(vt, nb, skv, heads, d).detach().requires_grad_(True)
        p = torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1)
        o = (p @ v).permute(0, 2, 1, 3).reshape(nb * sq, heads * d)
        o.backward(_f(go))
    for g, dst, s in ((q.grad, dq_out, sq), (k.grad, dk_out, skv), (v.grad, dv_out, skv)):
        if dst is not None:
            dst.copy_(g.permute(0, 2, 1, 3).reshape(nb * s, heads * d).to(dst.dtype))


# ------------------------------------------------------------------ text-encoder prologue (leco_b200/text_encoder.py)
def attention_v0(qt, kt, vt, nb, sq, skv, heads, d, scale, save_for_bwd=False, causal=False):
    q, k, v = _heads(qt, nb, sq, heads, d), _heads(kt, nb, skv, heads, d), _heads(vt, nb, skv, heads, d)
    s = q @ k.transpose(-1, -2) * scale
    if causal:
        s = s.masked_fill(torch.ones(sq, skv, dtype=torch.bool, device=s.device).triu(1), float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(nb * sq, heads * d).to(qt.dtype)
    return o, ((p,) if save_for_bwd else None)


def activation(x, kind):
    xf = _f(x)
    y = xf * torch.sigmoid(1.702 * xf) if kind == 1 else (F.gelu(xf) if kind == 2 else F.silu(xf))
    return y.to(x.dtype)


def embed_tokens(ids, tok, pos, seq):
    rows = torch.arange(ids.numel(), device=ids.device) % seq
    return (_f(tok)[ids.long()] + _f(pos)[rows]).to(tok.dtype)


def softmax_rows(s, n_valid, n_pad, causal_sq=0):
    rows = s.reshape(-1, s.shape[-1])[:, :n_valid].float().clone()
    if causal_sq:
        r = torch.arange(rows.shape[0], device=s.device) % causal_sq
        rows = rows.masked_fill(torch.arange(n_valid, device=s.device)[None] > r[:, None], float("-inf"))
    p = torch.zeros(rows.shape[0], n_pad, device=s.device)
    p[:, :n_valid] = torch.softmax(rows, dim=-1)
    return p.reshape(s.shape[:-1] + (n_pad,))
